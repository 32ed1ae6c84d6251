// kernels.cu — the wavefront trace loop of libxnart.so (hand-written CUDA, sm_100a).
//
// One frame = one launch sequence on one stream, no host round trip in between:
//   k_level<L = 0>        primary rays generated in-kernel (RayTracer.Render, RayTracer.cs:410-421),
//                         closest hit, shadow rays, shading; reflection/refraction rays appended to
//                         the next level's queue with warp-aggregated atomics
//   k_level<L = 1..D>     same, rays read from the queue (RayTracer.CastRay recursion, :506-737,
//                         turned into one wave per recursion depth)
//   k_resolve<L = D-1..0> bottom-up colour resolve; every level packs to an 8-bit Color exactly where
//                         the reference's recursion does (RayTracer.cs:584,699,705), level 0 writes
//                         renderTargetData (RayTracer.cs:425)
// Queue sizes live in device memory (levelCount/levelBase); every kernel is a grid-stride loop
// over "however many rays the previous level produced".
#include "kernels.h"
#include "trace.cuh"

#ifndef XN_LEVEL_MIN_BLOCKS
#define XN_LEVEL_MIN_BLOCKS 2
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

template <bool COUNT>
__global__ void __launch_bounds__(256, XN_LEVEL_MIN_BLOCKS) k_level(DevScene S, DevFrame F, DevArena A, const DevLight* __restrict__ lights, int level) {
    const unsigned count = (level == 0) ? (unsigned)F.chunkPixels : A.levelCount[level];
    const unsigned base = (level == 0) ? 0u : A.levelBase[level];
    const unsigned nextBase = base + count;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (level == 0) { A.levelCount[0] = count; A.levelBase[0] = 0u; }
        A.levelBase[level + 1] = nextBase;
    }
    const int in = level & 1, outq = in ^ 1;
    const bool last = !(level < F.maxReflections);  // RayTracer.cs:545
    Ctr ctr;
    ctr.nBox = ctr.nTriRef = ctr.nTriCulled = ctr.nTriFull = 0;
    unsigned nHits = 0, nRefl = 0, nRefr = 0;

    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x) {
        const unsigned node = base + idx;
        V3 ro, rd0;
        int ignM = -1, ignT = -1;
        float refIndex = 1.0f;
        if (level == 0) {
            int x, y;
            pixelOf(F, F.pixelBase + idx, x, y);
            ro = unproject(F, (float)x, (float)y, 0.0f);
            V3 farP = unproject(F, (float)x, (float)y, 1.0f);
            rd0 = normalize(sub(farP, ro));
        } else {
            float4 o4 = A.rayO[in][idx], d4 = A.rayD[in][idx];
            int2 ig = A.rayIgn[in][idx];
            ro = mk(o4.x, o4.y, o4.z); refIndex = o4.w;
            rd0 = mk(d4.x, d4.y, d4.z);
            ignM = ig.x; ignT = ig.y;
        }

        // q = 0: the closest-hit query of CastRay (:512); q = 1..nLights: IsLightPathObstructed (:465-502).
        // One call site keeps a single copy of the traversal code in the kernel.
        Hit h0;
        V3 N = mk(0.0f, 0.0f, 0.0f), lightResult = mk(0.0f, 0.0f, 0.0f);
        V3 qo = ro, qd = rd0;
        int qm = ignM, qt = ignT;
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
                if (M->flags & MESH_INTERP_NORMALS) {  // :520-527
                    const float4* __restrict__ T = S.triS + 5 * (size_t)triG;
                    float4 s1 = __ldg(&T[1]), s2 = __ldg(&T[2]), s3 = __ldg(&T[3]);
                    V3 n1 = mk(s1.z, s1.w, s2.x), n2 = mk(s2.y, s2.z, s2.w), n3 = mk(s3.x, s3.y, s3.z);
                    N = normalize(add(add(n1, mul(sub(n2, n1), h0.u)), mul(sub(n3, n1), h0.v)));
                } else {
                    float4 n4 = __ldg(&S.triI[4 * (size_t)triG]);
                    N = mk(n4.x, n4.y, n4.z);
                }
            } else {
                const DevLight L = lights[q - 1];
                float amount = 0.0f;
                if (ok && h.d < distToLight) {
                    const DevMesh* __restrict__ SM = &S.meshes[h.mesh];
                    amount = (SM->flags & MESH_TRANSPARENT) ? __ldg(&S.triS[5 * (size_t)(SM->triBase + h.tri) + 4]).z : 1.0f;
                }
                if (amount != 1.0f)  // :538-541
                    lightResult = add(lightResult, mul(lightForFragment(L, h0.pos, N), 1.0f - amount));
            }
            if (q == F.nLights) break;
            // shadow ray towards light q (no epsilon offset, ignoring only the hit triangle: Q4)
            const DevLight L = lights[q];
            V3 dir;
            if (L.kind == XNART_LIGHT_SPOT) {
                dir = sub(mk(L.px, L.py, L.pz), h0.pos);
                distToLight = length(dir);
                dir = normalize(dir);
            } else {
                dir = neg(mk(L.dx, L.dy, L.dz));
                distToLight = FLT_MAX;
            }
            qo = h0.pos; qd = dir; qm = h0.mesh; qt = h0.tri;
        }

        if (!hit) {
            A.color[node] = packColor(mk(0.0f, 0.0f, 0.0f), F.semantics);  // :732
            A.flags[node] = 0u;
            continue;
        }
        nHits++;

        // surface colour (:568-581 / :711-724)
        const float4* __restrict__ T = S.triS + 5 * (size_t)triG;
        V3 surf;
        float4 s3 = __ldg(&T[3]), s4 = __ldg(&T[4]);
        if (M->flags & MESH_USE_TEXTURE) {
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
            continue;
        }

        // reflection ray (:547-559) — always cast, whatever the reflectiveness (Q9)
        V3 rdir = normalize(reflect(rd0, N));
        unsigned reflChild, refrChild = 0xFFFFFFFFu;
        {
            cg::coalesced_group g = cg::coalesced_threads();
            unsigned b = 0;
            if (g.thread_rank() == 0) b = atomicAdd(&A.levelCount[level + 1], g.size());
            unsigned slot = g.shfl(b, 0) + g.thread_rank();
            reflChild = nextBase + slot;
            if (slot < A.rayCapacity && reflChild < A.nodeCapacity) {
                A.rayO[outq][slot] = make_float4(h0.pos.x, h0.pos.y, h0.pos.z, refIndex);
                A.rayD[outq][slot] = make_float4(rdir.x, rdir.y, rdir.z, 0.0f);
                A.rayIgn[outq][slot] = make_int2(h0.mesh, h0.tri);
            } else {
                *A.overflow = 1u;
            }
            nRefl++;
        }
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
            V3 refr;
            float w = (k * cos1) - cos2;
            if (cos1 >= 0.0f) refr = add(mul(rd0, k), mul(N, w));
            else refr = sub(mul(rd0, k), mul(N, w));
            V3 fdir = normalize(refr);
            cg::coalesced_group g = cg::coalesced_threads();
            unsigned b = 0;
            if (g.thread_rank() == 0) b = atomicAdd(&A.levelCount[level + 1], g.size());
            unsigned slot = g.shfl(b, 0) + g.thread_rank();
            refrChild = nextBase + slot;
            if (slot < A.rayCapacity && refrChild < A.nodeCapacity) {
                A.rayO[outq][slot] = make_float4(h0.pos.x, h0.pos.y, h0.pos.z, n2);
                A.rayD[outq][slot] = make_float4(fdir.x, fdir.y, fdir.z, 0.0f);
                A.rayIgn[outq][slot] = make_int2(h0.mesh, h0.tri);
            } else {
                *A.overflow = 1u;
            }
            nRefr++;
        }
        A.nodeA[node] = make_float4(lightResult.x, lightResult.y, lightResult.z, M->oneMinusReflectiveness);
        A.nodeB[node] = make_float4(surf.x, surf.y, surf.z, alphaW);
        A.nodeC[node] = make_uint2(reflChild, refrChild);
        A.flags[node] = 1u;
    }

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

// Bottom-up resolve of one level.  Nodes whose colour is final (miss, or deepest level) are left
// alone; the others combine their children as CastRay does after its recursive calls return.
__global__ void __launch_bounds__(256) k_resolve(DevFrame F, DevArena A, int level, uint32_t* __restrict__ out) {
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

// Batch form of ISpatialManager.GetRayIntersection for parity tests.
template <bool COUNT>
__global__ void __launch_bounds__(256, XN_LEVEL_MIN_BLOCKS) k_trace_rays(DevScene S, const float* __restrict__ rays6, long long n,
                                                    const int* __restrict__ ignore2, int* __restrict__ objId,
                                                    int* __restrict__ meshId, int* __restrict__ triId, float* __restrict__ u,
                                                    float* __restrict__ v, float* __restrict__ d, float* __restrict__ pos3,
                                                    unsigned long long* __restrict__ counters) {
    Ctr ctr;
    ctr.nBox = ctr.nTriRef = ctr.nTriCulled = ctr.nTriFull = 0;
    unsigned nHits = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        V3 ro = mk(rays6[6 * i], rays6[6 * i + 1], rays6[6 * i + 2]);
        V3 rd = mk(rays6[6 * i + 3], rays6[6 * i + 4], rays6[6 * i + 5]);
        int im = ignore2 ? ignore2[2 * i] : -1, it = ignore2 ? ignore2[2 * i + 1] : -1;
        Hit h;
        bool ok = traceScene<COUNT>(S, ro, rd, im, it, h, ctr);
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
                        long long upperBound, int maxBlocks, bool count, cudaStream_t st) {
    int grid = gridFor(upperBound, maxBlocks);
    if (count) k_level<true><<<grid, 256, 0, st>>>(S, F, A, lights, level);
    else k_level<false><<<grid, 256, 0, st>>>(S, F, A, lights, level);
    return cudaGetLastError();
}

cudaError_t launchResolve(const DevFrame& F, const DevArena& A, int level, long long upperBound, int maxBlocks,
                          uint32_t* out, cudaStream_t st) {
    int grid = gridFor(upperBound, maxBlocks);
    k_resolve<<<grid, 256, 0, st>>>(F, A, level, out);
    return cudaGetLastError();
}

cudaError_t launchTraceRays(const DevScene& S, const float* rays6, long long n, const int* ignore2, int* objId, int* meshId,
                            int* triId, float* u, float* v, float* d, float* pos3, unsigned long long* counters,
                            int maxBlocks, bool count, cudaStream_t st) {
    int grid = gridFor(n, maxBlocks);
    if (count) k_trace_rays<true><<<grid, 256, 0, st>>>(S, rays6, n, ignore2, objId, meshId, triId, u, v, d, pos3, counters);
    else k_trace_rays<false><<<grid, 256, 0, st>>>(S, rays6, n, ignore2, objId, meshId, triId, u, v, d, pos3, counters);
    return cudaGetLastError();
}

}  // namespace xn
