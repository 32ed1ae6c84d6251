// kernels.cu — the wavefront trace loop of libxnart.so (hand-written CUDA, sm_100a).
//
// One frame = one launch sequence on one stream, no host round trip in between:
//   k_level / k_level_coop, level 0      primary rays generated in-kernel (RayTracer.Render,
//                         RayTracer.cs:410-421), closest hit, shadow rays, shading; reflection and
//                         refraction rays appended to the next level's queue with warp-aggregated atomics
//   k_level / k_level_coop, level 1..D   same, rays read from the queue (the RayTracer.CastRay
//                         recursion, :506-737, turned into one wave per recursion depth)
//   k_resolve, level D-1..0   bottom-up colour resolve; every level packs to an 8-bit Color exactly
//                         where the reference's recursion does (RayTracer.cs:584,699,705), level 0
//                         writes renderTargetData (RayTracer.cs:425)
// Queue sizes live in device memory (levelCount/levelBase); every kernel is a grid-stride loop
// over "however many rays the previous level produced".
//
// Two schedules of the same work: k_level walks one ray per thread (best when the octrees are a
// leaf or two deep, e.g. the 14-triangle crate scene); k_level_coop lets the 8 lanes of an octet
// walk each ray together (trace.cuh: traceMeshOctet), which keeps deep-tree traversal convergent.
#include "kernels.h"
#include "trace.cuh"

#ifndef XN_LEVEL_MIN_BLOCKS
#define XN_LEVEL_MIN_BLOCKS 2
#endif
#ifndef XN_COOP_MIN_BLOCKS
#define XN_COOP_MIN_BLOCKS 3
#endif
#ifndef XN_DEEP_MIN_BLOCKS
#define XN_DEEP_MIN_BLOCKS 4
#endif

namespace xn {

__device__ __forceinline__ unsigned warpSum(unsigned v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// shard-local pixel index -> (x, y) of the full frame
__device__ __forceinline__ void pixelOf(const DevFrame& F, long long q, int& x, int& y) {
    int lr = (int)(q / F.vpW);
    x = (int)(q - (long long)lr * F.vpW);
    if (F.shardCount > 1) {
        int lb = lr / F.shardBandRows, within = lr - lb * F.shardBandRows;
        y = (lb * F.shardCount + F.shardIndex) * F.shardBandRows + within;
    } else {
        y = lr;
    }
}

// Viewport.Unproject with the frame-constant inverse.
__device__ __forceinline__ V3 unproject(const DevFrame& F, float sx, float sy, float sz) {
    V3 s;
    s.x = (((sx - (float)F.vpX) / (float)F.vpW) * 2.0f) - 1.0f;
    s.y = -((((sy - (float)F.vpY) / (float)F.vpH) * 2.0f) - 1.0f);
    s.z = (sz - F.minDepth) / (F.maxDepth - F.minDepth);
    const float* m = F.inv;
    V3 v = mk((((s.x * m[0]) + (s.y * m[4])) + (s.z * m[8])) + m[12],
              (((s.x * m[1]) + (s.y * m[5])) + (s.z * m[9])) + m[13],
              (((s.x * m[2]) + (s.y * m[6])) + (s.z * m[10])) + m[14]);
    float a = (((s.x * m[3]) + (s.y * m[7])) + (s.z * m[11])) + m[15];
    float num = a - 1.0f;
    bool within = (-1.401298E-45f <= num) && (num <= 1.401298E-45f);
    if (!within) v = divs(v, a);
    return v;
}

struct RayIn {
    V3 o, d;
    int ignM, ignT;
    float refIndex;
};

// Level 0: RayTracer.Render's per-pixel Unproject pair (:412-421).  Level > 0: the queued ray.
__device__ __forceinline__ RayIn loadRay(const DevFrame& F, const DevArena& A, int level, unsigned idx) {
    RayIn r;
    if (level == 0) {
        int x, y;
        pixelOf(F, F.pixelBase + idx, x, y);
        r.o = unproject(F, (float)x, (float)y, 0.0f);
        V3 farP = unproject(F, (float)x, (float)y, 1.0f);
        r.d = normalize(sub(farP, r.o));
        r.ignM = -1; r.ignT = -1; r.refIndex = 1.0f;
    } else {
        const int in = level & 1;
        float4 o4 = A.rayO[in][idx], d4 = A.rayD[in][idx];
        int2 ig = A.rayIgn[in][idx];
        r.o = mk(o4.x, o4.y, o4.z); r.refIndex = o4.w;
        r.d = mk(d4.x, d4.y, d4.z);
        r.ignM = ig.x; r.ignT = ig.y;
    }
    return r;
}

// fragment normal (:520-531)
__device__ __forceinline__ V3 fragmentNormal(const DevScene& S, const DevMesh* __restrict__ M, int triG, const Hit& h) {
    if (M->flags & MESH_INTERP_NORMALS) {
        const float4* __restrict__ T = S.triS + 5 * (size_t)triG;
        float4 s1 = __ldg(&T[1]), s2 = __ldg(&T[2]), s3 = __ldg(&T[3]);
        V3 n1 = mk(s1.z, s1.w, s2.x), n2 = mk(s2.y, s2.z, s2.w), n3 = mk(s3.x, s3.y, s3.z);
        return normalize(add(add(n1, mul(sub(n2, n1), h.u)), mul(sub(n3, n1), h.v)));
    }
    float4 n4 = __ldg(&S.leafTri[4 * (size_t)h.key]);
    return mk(n4.x, n4.y, n4.z);
}

// shadow ray towards a light (:467-482): no epsilon offset, ignoring only the hit triangle (Q4)
__device__ __forceinline__ V3 shadowDir(const DevLight& L, V3 pos, float& distToLight) {
    if (L.kind == XNART_LIGHT_SPOT) {
        V3 dir = sub(mk(L.px, L.py, L.pz), pos);
        distToLight = length(dir);
        return normalize(dir);
    }
    distToLight = FLT_MAX;
    return neg(mk(L.dx, L.dy, L.dz));
}

// IsLightPathObstructed's verdict (:485-501) folded into the light sum (:537-541)
__device__ __forceinline__ V3 addLight(const DevScene& S, const DevLight& L, bool shadowHit, const Hit& sh, float distToLight,
                                       V3 pos, V3 N, V3 lightResult) {
    float amount = 0.0f;
    if (shadowHit && sh.d < distToLight) {
        const DevMesh* __restrict__ SM = &S.meshes[sh.mesh];
        amount = (SM->flags & MESH_TRANSPARENT) ? __ldg(&S.triS[5 * (size_t)(SM->triBase + sh.tri) + 4]).z : 1.0f;
    }
    if (amount != 1.0f) lightResult = add(lightResult, mul(lightForFragment(L, pos, N), 1.0f - amount));
    return lightResult;
}

__device__ __forceinline__ unsigned appendRay(const DevArena& A, int level, int outq, unsigned nextBase, V3 o, V3 d, float refIndex,
                                              int mesh, int tri) {
    cg::coalesced_group g = cg::coalesced_threads();
    unsigned b = 0;
    if (g.thread_rank() == 0) b = atomicAdd(&A.levelCount[level + 1], g.size());
    const unsigned slot = g.shfl(b, 0) + g.thread_rank();
    const unsigned child = nextBase + slot;
    if (slot < A.rayCapacity && child < A.nodeCapacity) {
        A.rayO[outq][slot] = make_float4(o.x, o.y, o.z, refIndex);
        A.rayD[outq][slot] = make_float4(d.x, d.y, d.z, 0.0f);
        A.rayIgn[outq][slot] = make_int2(mesh, tri);
    } else {
        *A.overflow = 1u;
    }
    return child;
}

// Everything CastRay does after the light loop (:545-727): surface colour, the path node, and the
// reflection / refraction rays of the next level.
__device__ __forceinline__ void finishHit(const DevScene& S, const DevFrame& F, const DevArena& A, int level, bool last, unsigned node,
                                          unsigned nextBase, const Hit& h0, const DevMesh* __restrict__ M, int triG, V3 N,
                                          V3 lightResult, V3 rd0, float refIndex, unsigned& nRefl, unsigned& nRefr) {
    const float4* __restrict__ T = S.triS + 5 * (size_t)triG;
    V3 surf;
    float4 s3 = __ldg(&T[3]), s4 = __ldg(&T[4]);
    if (M->flags & MESH_USE_TEXTURE) {  // :568-575
        float4 s0 = __ldg(&T[0]), s1 = __ldg(&T[1]);
        float ax = s0.z - s0.x, ay = s0.w - s0.y;  // uv2 - uv1
        float bx = s1.x - s0.x, by = s1.y - s0.y;  // uv3 - uv1
        float tu = (s0.x + (ax * h0.u)) + (bx * h0.v);
        float tv = (s0.y + (ay * h0.u)) + (by * h0.v);
        surf = lookupUV(S.texels + M->texOffset, M->texW, M->texH, tu, tv, F.addressMode, F.filtering);
    } else {
        surf = mk(s3.w, s4.x, s4.y);
    }
    if (last) {
        A.color[node] = packColor(mulv(lightResult, surf), F.semantics);  // :726
        A.flags[node] = 0u;
        return;
    }
    const int outq = (level & 1) ^ 1;
    // reflection ray (:547-559) — always cast, whatever the reflectiveness (Q9)
    V3 rdir = normalize(reflect(rd0, N));
    unsigned reflChild = appendRay(A, level, outq, nextBase, h0.pos, rdir, refIndex, h0.mesh, h0.tri);
    unsigned refrChild = 0xFFFFFFFFu;
    nRefl++;
    const float alphaW = s4.z;  // triangle.color.W
    if (M->flags & MESH_TRANSPARENT) {
        // refraction (:656-698), double-precision island for the Snell term
        float n1, n2;
        if (refIndex == M->refractionIndex) { n1 = 1.0f; n2 = refIndex; }
        else { n1 = M->refractionIndex; n2 = 1.0f; }
        float cos1 = dot(N, neg(rd0));
        float k = n1 / n2;
        double ratio = (double)k, c1 = (double)cos1;
        float cos2 = (float)sqrt(1.0 - (ratio * ratio) * (1.0 - (c1 * c1)));
        float w = (k * cos1) - cos2;
        V3 refr = (cos1 >= 0.0f) ? add(mul(rd0, k), mul(N, w)) : sub(mul(rd0, k), mul(N, w));
        refrChild = appendRay(A, level, outq, nextBase, h0.pos, normalize(refr), n2, h0.mesh, h0.tri);
        nRefr++;
    }
    A.nodeA[node] = make_float4(lightResult.x, lightResult.y, lightResult.z, M->oneMinusReflectiveness);
    A.nodeB[node] = make_float4(surf.x, surf.y, surf.z, alphaW);
    A.nodeC[node] = make_uint2(reflChild, refrChild);
    A.flags[node] = 1u;
}

template <bool COUNT>
__device__ __forceinline__ void flushCounters(const DevArena& A, const Ctr& ctr, unsigned nHits, unsigned nRefl, unsigned nRefr) {
    // ray statistics: one atomic per warp and counter
    const unsigned lane = threadIdx.x & 31u;
    unsigned sH = warpSum(nHits), sR = warpSum(nRefl), sF = warpSum(nRefr);
    if (lane == 0) {
        if (sH) atomicAdd(&A.counters[CTR_HITS], (unsigned long long)sH);
        if (sR) atomicAdd(&A.counters[CTR_RAYS_REFLECTION], (unsigned long long)sR);
        if (sF) atomicAdd(&A.counters[CTR_RAYS_REFRACTION], (unsigned long long)sF);
    }
    if (COUNT) {
        unsigned a = warpSum(ctr.nBox), b = warpSum(ctr.nTriRef), c = warpSum(ctr.nTriCulled), d = warpSum(ctr.nTriFull);
        if (lane == 0) {
            atomicAdd(&A.counters[CTR_NBOX], (unsigned long long)a);
            atomicAdd(&A.counters[CTR_NTRIREF], (unsigned long long)b);
            atomicAdd(&A.counters[CTR_NTRICULLED], (unsigned long long)c);
            atomicAdd(&A.counters[CTR_NTRIFULL], (unsigned long long)d);
        }
    }
}

__device__ __forceinline__ void levelHeader(const DevFrame& F, const DevArena& A, int level, unsigned& count, unsigned& base) {
    count = (level == 0) ? (unsigned)F.chunkPixels : A.levelCount[level];
    base = (level == 0) ? 0u : A.levelBase[level];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (level == 0) { A.levelCount[0] = count; A.levelBase[0] = 0u; }
        A.levelBase[level + 1] = base + count;
    }
}

// ---------------------------------------------------------------------------------------------
// one thread, one ray
// ---------------------------------------------------------------------------------------------
// MINB = resident blocks per SM the register allocation is sized for: 2 (128 registers) suits
// leaf-sized meshes where the kernel is issue bound; 4 (64 registers, some spilling) suits deep
// octrees, where the walk is bound by dependent-load latency and more resident warps hide it.
template <bool COUNT>
__device__ __forceinline__ void levelBody(const DevScene& S, const DevFrame& F, const DevArena& A, const DevLight* __restrict__ lights,
                                          int level, Ctr& ctr, unsigned& nHits, unsigned& nRefl, unsigned& nRefr) {
    unsigned count, base;
    levelHeader(F, A, level, count, base);
    const unsigned nextBase = base + count;
    const bool last = !(level < F.maxReflections);  // RayTracer.cs:545

    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x) {
        const unsigned node = base + idx;
        const RayIn ray = loadRay(F, A, level, idx);
        // q = 0: the closest-hit query of CastRay (:512); q = 1..nLights: IsLightPathObstructed (:465-502).
        // One call site keeps a single copy of the traversal code in the kernel.
        Hit h0;
        V3 N = mk(0.0f, 0.0f, 0.0f), lightResult = mk(0.0f, 0.0f, 0.0f);
        V3 qo = ray.o, qd = ray.d;
        int qm = ray.ignM, qt = ray.ignT;
        float distToLight = 0.0f;
        bool hit = false;
        const DevMesh* __restrict__ M = nullptr;
        int triG = 0;
        for (int q = 0;; q++) {
            Hit h;
            bool ok = traceScene<COUNT>(S, qo, qd, qm, qt, h, ctr);
            if (q == 0) {
                if (!ok) break;
                hit = true;
                h0 = h;
                M = &S.meshes[h0.mesh];
                triG = M->triBase + h0.tri;
                N = fragmentNormal(S, M, triG, h0);
            } else {
                lightResult = addLight(S, lights[q - 1], ok, h, distToLight, h0.pos, N, lightResult);
            }
            if (q == F.nLights) break;
            qd = shadowDir(lights[q], h0.pos, distToLight);
            qo = h0.pos; qm = h0.mesh; qt = h0.tri;
        }
        if (!hit) {
            A.color[node] = packColor(mk(0.0f, 0.0f, 0.0f), F.semantics);  // :732
            A.flags[node] = 0u;
            continue;
        }
        nHits++;
        finishHit(S, F, A, level, last, node, nextBase, h0, M, triG, N, lightResult, ray.d, ray.refIndex, nRefl, nRefr);
    }
}

template <bool COUNT, int MINB>
__global__ void __launch_bounds__(256, MINB) k_level(DevScene S, DevFrame F, DevArena A, const DevLight* __restrict__ lights, int level) {
    Ctr ctr;
    ctr.nBox = ctr.nTriRef = ctr.nTriCulled = ctr.nTriFull = 0;
    unsigned nHits = 0, nRefl = 0, nRefr = 0;
    levelBody<COUNT>(S, F, A, lights, level, ctr, nHits, nRefl, nRefr);
    flushCounters<COUNT>(A, ctr, nHits, nRefl, nRefr);
}

__device__ __forceinline__ void resolveBody(const DevFrame& F, const DevArena& A, int level, uint32_t* __restrict__ out);

__device__ __forceinline__ unsigned long long globalTimerNs() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// The whole frame in ONE cooperative launch: level 0..D, then the resolve D-1..0, separated by
// grid-wide barriers instead of kernel boundaries.  On small scenes the deep levels hold a few
// thousand rays each and a launch per level costs more than the level itself (ncu launch list,
// profiles/r01a: 6 tail levels x 10 us + 8 resolve launches on a 0.65 ms frame).  Grid = resident
// blocks only (cooperative launch), every phase is a grid-stride loop.  stamps[0..2] receive the
// global timer at start / after the last level / after the resolve, for the per-stage timings.
template <bool COUNT>
__global__ void __launch_bounds__(256, XN_LEVEL_MIN_BLOCKS) k_frame(DevScene S, DevFrame F, DevArena A, const DevLight* __restrict__ lights,
                                                                     uint32_t* __restrict__ out, unsigned long long* __restrict__ stamps) {
    cg::grid_group grid = cg::this_grid();
    Ctr ctr;
    ctr.nBox = ctr.nTriRef = ctr.nTriCulled = ctr.nTriFull = 0;
    unsigned nHits = 0, nRefl = 0, nRefr = 0;
    const bool stamper = blockIdx.x == 0 && threadIdx.x == 0;
    if (stamper) stamps[0] = globalTimerNs();
    const int D = F.maxReflections;
    for (int level = 0; level <= D; level++) {
        levelBody<COUNT>(S, F, A, lights, level, ctr, nHits, nRefl, nRefr);
        grid.sync();
    }
    if (stamper) stamps[1] = globalTimerNs();
    for (int level = (D > 0 ? D - 1 : 0); level >= 0; level--) {
        resolveBody(F, A, level, out);
        if (level > 0) grid.sync();
    }
    flushCounters<COUNT>(A, ctr, nHits, nRefl, nRefr);
    if (stamper) stamps[2] = globalTimerNs();
}

// ---------------------------------------------------------------------------------------------
// one octet (8 lanes), one ray at a time: every lane still OWNS one ray of the queue (loads it,
// shades it, emits its children), but all traversal is done through traceSceneCoop, where the 8
// lanes of an octet serve each other's rays one after the other.  Loops are octet-uniform.
// ---------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(256, XN_COOP_MIN_BLOCKS) k_level_coop(DevScene S, DevFrame F, DevArena A, const DevLight* __restrict__ lights, int level) {
    const Octet g = cg::tiled_partition<8>(cg::this_thread_block());
    unsigned count, base;
    levelHeader(F, A, level, count, base);
    const unsigned nextBase = base + count;
    const bool last = !(level < F.maxReflections);
    Ctr ctr;
    ctr.nBox = ctr.nTriRef = ctr.nTriCulled = ctr.nTriFull = 0;
    unsigned nHits = 0, nRefl = 0, nRefr = 0;
    const unsigned lane8 = threadIdx.x & 7u;

    for (unsigned idx0 = blockIdx.x * blockDim.x + (threadIdx.x & ~7u); idx0 < count; idx0 += gridDim.x * blockDim.x) {
        const unsigned idx = idx0 + lane8;
        const bool valid = idx < count;
        const unsigned node = base + idx;
        RayIn ray;
        ray.o = ray.d = mk(0.0f, 0.0f, 0.0f); ray.ignM = ray.ignT = -1; ray.refIndex = 1.0f;
        if (valid) ray = loadRay(F, A, level, idx);

        Hit h0;
        h0.obj = h0.mesh = h0.tri = -1; h0.key = 0; h0.u = h0.v = h0.d = 0.0f; h0.pos = mk(0.0f, 0.0f, 0.0f);
        bool hit = traceSceneCoop<COUNT>(g, S, valid, ray.o, ray.d, ray.ignM, ray.ignT, h0, ctr);
        V3 N = mk(0.0f, 0.0f, 0.0f), lightResult = mk(0.0f, 0.0f, 0.0f);
        const DevMesh* __restrict__ M = nullptr;
        int triG = 0;
        if (hit) {
            M = &S.meshes[h0.mesh];
            triG = M->triBase + h0.tri;
            N = fragmentNormal(S, M, triG, h0);
        }
        for (int q = 0; q < F.nLights; q++) {  // uniform trip count: lanes without a hit just help
            float distToLight = 0.0f;
            V3 sd = mk(0.0f, 0.0f, 0.0f);
            if (hit) sd = shadowDir(lights[q], h0.pos, distToLight);
            Hit sh;
            sh.obj = sh.mesh = sh.tri = -1; sh.key = 0; sh.u = sh.v = sh.d = 0.0f; sh.pos = mk(0.0f, 0.0f, 0.0f);
            bool ok = traceSceneCoop<COUNT>(g, S, hit, h0.pos, sd, h0.mesh, h0.tri, sh, ctr);
            if (hit) lightResult = addLight(S, lights[q], ok, sh, distToLight, h0.pos, N, lightResult);
        }
        if (valid && !hit) {
            A.color[node] = packColor(mk(0.0f, 0.0f, 0.0f), F.semantics);
            A.flags[node] = 0u;
        }
        if (hit) {
            nHits++;
            finishHit(S, F, A, level, last, node, nextBase, h0, M, triG, N, lightResult, ray.d, ray.refIndex, nRefl, nRefr);
        }
    }
    flushCounters<COUNT>(A, ctr, nHits, nRefl, nRefr);
}

// Bottom-up resolve of one level.  Nodes whose colour is final (miss, or deepest level) are left
// alone; the others combine their children as CastRay does after its recursive calls return.
__device__ __forceinline__ void resolveBody(const DevFrame& F, const DevArena& A, int level, uint32_t* __restrict__ out) {
    const unsigned count = A.levelCount[level];
    const unsigned base = A.levelBase[level];
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x) {
        const unsigned node = base + idx;
        uint32_t col;
        if (A.flags[node]) {
            float4 a = A.nodeA[node], b = A.nodeB[node];
            uint2 ch = A.nodeC[node];
            uint32_t reflColor = (ch.x < A.nodeCapacity) ? A.color[ch.x] : 0xFF000000u;
            // Lerp(reflectionColor.ToVector3(), surfaceColor, 1 - Reflectiveness) * lightResult (:584)
            V3 cv = mulv(lerp(unpackColor(reflColor), mk(b.x, b.y, b.z), a.w), mk(a.x, a.y, a.z));
            if (ch.y != 0xFFFFFFFFu) {
                uint32_t refrColor = (ch.y < A.nodeCapacity) ? A.color[ch.y] : 0xFF000000u;
                cv = lerp(unpackColor(refrColor), cv, b.w);  // :699
            }
            col = packColor(cv, F.semantics);  // :705
            A.color[node] = col;
        } else {
            col = A.color[node];
        }
        if (level == 0) {
            int x, y;
            pixelOf(F, F.pixelBase + idx, x, y);
            out[(size_t)y * (size_t)F.vpW + (size_t)x] = col;  // renderTargetData[y*W + x] (:425)
        }
    }
}

__global__ void __launch_bounds__(256) k_resolve(DevFrame F, DevArena A, int level, uint32_t* __restrict__ out) {
    resolveBody(F, A, level, out);
}

// Batch form of ISpatialManager.GetRayIntersection for parity tests (both schedules).
template <bool COUNT, bool COOP>
__global__ void __launch_bounds__(256, XN_LEVEL_MIN_BLOCKS) k_trace_rays(DevScene S, const float* __restrict__ rays6, long long n,
                                                                         const int* __restrict__ ignore2, int* __restrict__ objId,
                                                                         int* __restrict__ meshId, int* __restrict__ triId,
                                                                         float* __restrict__ u, float* __restrict__ v, float* __restrict__ d,
                                                                         float* __restrict__ pos3, unsigned long long* __restrict__ counters) {
    const Octet g = cg::tiled_partition<8>(cg::this_thread_block());
    Ctr ctr;
    ctr.nBox = ctr.nTriRef = ctr.nTriCulled = ctr.nTriFull = 0;
    unsigned nHits = 0;
    const long long lane8 = threadIdx.x & 7u;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~7u); i0 < n; i0 += (long long)gridDim.x * blockDim.x) {
        const long long i = i0 + lane8;
        const bool valid = i < n;
        V3 ro = mk(0.0f, 0.0f, 0.0f), rd = mk(0.0f, 0.0f, 0.0f);
        int im = -1, it = -1;
        if (valid) {
            ro = mk(rays6[6 * i], rays6[6 * i + 1], rays6[6 * i + 2]);
            rd = mk(rays6[6 * i + 3], rays6[6 * i + 4], rays6[6 * i + 5]);
            if (ignore2) { im = ignore2[2 * i]; it = ignore2[2 * i + 1]; }
        }
        Hit h;
        h.obj = h.mesh = h.tri = -1; h.key = 0; h.u = h.v = h.d = 0.0f; h.pos = mk(0.0f, 0.0f, 0.0f);
        bool ok;
        if (COOP) ok = traceSceneCoop<COUNT>(g, S, valid, ro, rd, im, it, h, ctr);
        else ok = valid && traceScene<COUNT>(S, ro, rd, im, it, h, ctr);
        if (!valid) continue;
        if (ok) nHits++;
        objId[i] = ok ? h.obj : -1; meshId[i] = ok ? h.mesh : -1; triId[i] = ok ? h.tri : -1;
        u[i] = ok ? h.u : 0.0f; v[i] = ok ? h.v : 0.0f; d[i] = ok ? h.d : 0.0f;
        if (pos3) { pos3[3 * i] = ok ? h.pos.x : 0.0f; pos3[3 * i + 1] = ok ? h.pos.y : 0.0f; pos3[3 * i + 2] = ok ? h.pos.z : 0.0f; }
    }
    const unsigned lane = threadIdx.x & 31u;
    unsigned sH = warpSum(nHits);
    if (lane == 0 && sH) atomicAdd(&counters[CTR_HITS], (unsigned long long)sH);
    if (COUNT) {
        unsigned a = warpSum(ctr.nBox), b = warpSum(ctr.nTriRef), c = warpSum(ctr.nTriCulled), e = warpSum(ctr.nTriFull);
        if (lane == 0) {
            atomicAdd(&counters[CTR_NBOX], (unsigned long long)a);
            atomicAdd(&counters[CTR_NTRIREF], (unsigned long long)b);
            atomicAdd(&counters[CTR_NTRICULLED], (unsigned long long)c);
            atomicAdd(&counters[CTR_NTRIFULL], (unsigned long long)e);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// launch wrappers
// ---------------------------------------------------------------------------------------------
static inline int gridFor(long long n, int maxBlocks) {
    long long b = (n + 255) / 256;
    if (b < 1) b = 1;
    if (b > maxBlocks) b = maxBlocks;
    return (int)b;
}

cudaError_t launchLevel(const DevScene& S, const DevFrame& F, const DevArena& A, const DevLight* lights, int level,
                        long long upperBound, int maxBlocks, bool count, bool coop, bool deep, cudaStream_t st) {
    int grid = gridFor(upperBound, maxBlocks);
    if (coop) {
        if (count) k_level_coop<true><<<grid, 256, 0, st>>>(S, F, A, lights, level);
        else k_level_coop<false><<<grid, 256, 0, st>>>(S, F, A, lights, level);
    } else if (deep) {
        if (count) k_level<true, XN_DEEP_MIN_BLOCKS><<<grid, 256, 0, st>>>(S, F, A, lights, level);
        else k_level<false, XN_DEEP_MIN_BLOCKS><<<grid, 256, 0, st>>>(S, F, A, lights, level);
    } else {
        if (count) k_level<true, XN_LEVEL_MIN_BLOCKS><<<grid, 256, 0, st>>>(S, F, A, lights, level);
        else k_level<false, XN_LEVEL_MIN_BLOCKS><<<grid, 256, 0, st>>>(S, F, A, lights, level);
    }
    return cudaGetLastError();
}

// Resident-grid size of k_frame on this device (cooperative launch needs every block resident).
int frameKernelMaxBlocks(bool count) {
    int perSm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = count ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_frame<true>, 256, 0)
                          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_frame<false>, 256, 0);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return perSm * sms;
}

cudaError_t launchFrame(const DevScene& S, const DevFrame& F, const DevArena& A, const DevLight* lights, uint32_t* out,
                        unsigned long long* stamps, int blocks, bool count, cudaStream_t st) {
    void* args[] = {(void*)&S, (void*)&F, (void*)&A, (void*)&lights, (void*)&out, (void*)&stamps};
    const void* fn = count ? (const void*)k_frame<true> : (const void*)k_frame<false>;
    return cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(256), args, 0, st);
}

cudaError_t launchResolve(const DevFrame& F, const DevArena& A, int level, long long upperBound, int maxBlocks,
                          uint32_t* out, cudaStream_t st) {
    int grid = gridFor(upperBound, maxBlocks);
    k_resolve<<<grid, 256, 0, st>>>(F, A, level, out);
    return cudaGetLastError();
}

cudaError_t launchTraceRays(const DevScene& S, const float* rays6, long long n, const int* ignore2, int* objId, int* meshId,
                            int* triId, float* u, float* v, float* d, float* pos3, unsigned long long* counters,
                            int maxBlocks, bool count, bool coop, cudaStream_t st) {
    int grid = gridFor(n, maxBlocks);
    if (coop) {
        if (count) k_trace_rays<true, true><<<grid, 256, 0, st>>>(S, rays6, n, ignore2, objId, meshId, triId, u, v, d, pos3, counters);
        else k_trace_rays<false, true><<<grid, 256, 0, st>>>(S, rays6, n, ignore2, objId, meshId, triId, u, v, d, pos3, counters);
    } else {
        if (count) k_trace_rays<true, false><<<grid, 256, 0, st>>>(S, rays6, n, ignore2, objId, meshId, triId, u, v, d, pos3, counters);
        else k_trace_rays<false, false><<<grid, 256, 0, st>>>(S, rays6, n, ignore2, objId, meshId, triId, u, v, d, pos3, counters);
    }
    return cudaGetLastError();
}

}  // namespace xn
