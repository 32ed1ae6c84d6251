// kernels.cu — the trace loop of libxnart.so for small MeshOctrees (hand-written CUDA, sm_100a): one thread
// per ray, the whole CastRay per thread.  (Deep octrees: wavefront.cu.)
//
// One frame = one launch sequence on one stream, no host round trip in between:
//   k_level, level 0      primary rays generated in-kernel (RayTracer.Render, RayTracer.cs:410-421), closest
//                         hit, shadow rays, shading; reflection and refraction rays appended to the next
//                         level's queue with warp-aggregated atomics
//   k_level, level 1..D   same, rays read from the queue (the RayTracer.CastRay recursion, :506-737, turned
//                         into one wave per recursion depth)
//   k_resolve_chain       opaque scenes: one thread folds a pixel's reflection chain bottom-up, or
//   k_resolve, level D-1..0   level-wise bottom-up colour resolve (binary path trees of transparent scenes);
//                         every level packs to an 8-bit Color exactly where the reference's recursion does
//                         (RayTracer.cs:584,699,705), level 0 writes renderTargetData (RayTracer.cs:425)
//   k_frame               optional: all of the above in one cooperative launch (XNART_OPT_FUSE_FRAME)
// Queue sizes live in device memory (levelCount/levelBase); every kernel is a grid-stride loop over "however
// many rays the previous level produced".
#include <algorithm>

#include "kernels.h"
#include "shade.cuh"

#ifndef XN_LEVEL_MIN_BLOCKS
#define XN_LEVEL_MIN_BLOCKS 2
#endif
#ifndef XN_LEAF_MIN_BLOCKS
#define XN_LEAF_MIN_BLOCKS 6
#endif
#ifndef XN_LEAF_THREADS
#define XN_LEAF_THREADS 128
#endif
#ifndef XN_DEEP_MIN_BLOCKS
#define XN_DEEP_MIN_BLOCKS 4
#endif

namespace xn {

// ---------------------------------------------------------------------------------------------
// one thread, one ray
// ---------------------------------------------------------------------------------------------
// MINB = resident blocks per SM the register allocation is sized for: 2 (128 registers) suits
// leaf-sized meshes where the kernel is issue bound; 4 (64 registers, some spilling) suits deep
// octrees, where the walk is bound by dependent-load latency and more resident warps hide it.
template <bool COUNT, bool LEAFONLY, bool MONO = false>
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
            bool ok = traceScene<COUNT, LEAFONLY, MONO>(S, qo, qd, qm, qt, h, ctr);
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

// THREADS x MINB fixes the register budget (128 x 6 or 256 x 3 -> 80 registers, 128 x 5 -> 96, 256 x 2 -> 128, 256 x 4 -> 64).
// Measured on the 1080p crate frame (levels, ms): 128x6 0.404, 256x3 0.417, 128x5 0.415, 128x4 0.418, 192x3 0.438.
// MONO: the MonoGame form of BoundingBox.Intersects (trace.cuh: slabMonoGame) everywhere a box is tested; one generic
// build serves every scene (XnartFrame.semantics == XNART_SEMANTICS_MONOGAME is a compatibility mode, not a fast path).
template <bool COUNT, int MINB, bool LEAFONLY, int THREADS = 256, bool MONO = false>
__global__ void __launch_bounds__(THREADS, MINB) k_level(DevScene S, DevFrame F, DevArena A, const DevLight* __restrict__ lights, int level) {
    Ctr ctr;
    ctr.nBox = ctr.nTriRef = ctr.nTriCulled = ctr.nTriFull = 0;
    unsigned nHits = 0, nRefl = 0, nRefr = 0;
    levelBody<COUNT, LEAFONLY, MONO>(S, F, A, lights, level, ctr, nHits, nRefl, nRefr);
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
        levelBody<COUNT, false>(S, F, A, lights, level, ctr, nHits, nRefl, nRefr);
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

// Opaque scenes: the path of a pixel is a chain (one reflection ray per hit, RayTracer.cs:545-559), so one
// thread can fold it without the per-level launches: walk down to the node whose colour is final (a miss, or
// the deepest level), then apply CastRay's combination (:584, :705) level by level on the way back up, packing
// to the 8-bit Color at every level exactly as the recursion does.  Same arithmetic as resolveBody.
__global__ void __launch_bounds__(256) k_resolve_chain(DevFrame F, DevArena A, uint32_t* __restrict__ out) {
    const unsigned count = (unsigned)F.chunkPixels;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x) {
        unsigned path[32];   // kMaxLevels + 2 in xnart_capi.cu
        int depth = 0;
        unsigned node = idx;
        while (A.flags[node] && depth < 31) {
            path[depth++] = node;
            const unsigned child = A.nodeC[node].x;
            if (!(child < A.nodeCapacity)) { node = 0xFFFFFFFFu; break; }
            node = child;
        }
        uint32_t col = (node != 0xFFFFFFFFu) ? A.color[node] : 0xFF000000u;
        for (int k = depth - 1; k >= 0; k--) {
            const float4 a = A.nodeA[path[k]], b = A.nodeB[path[k]];
            // Lerp(reflectionColor.ToVector3(), surfaceColor, 1 - Reflectiveness) * lightResult (:584)
            const V3 cv = mulv(lerp(unpackColor(col), mk(b.x, b.y, b.z), a.w), mk(a.x, a.y, a.z));
            col = packColor(cv, F.semantics);  // :705
        }
        int x, y;
        pixelOf(F, F.pixelBase + idx, x, y);
        out[(size_t)y * (size_t)F.vpW + (size_t)x] = col;  // renderTargetData[y*W + x] (:425)
    }
}

// Batch form of ISpatialManager.GetRayIntersection for parity tests (one thread per ray).
template <bool COUNT, bool MONO = false>
__global__ void __launch_bounds__(256, XN_LEVEL_MIN_BLOCKS) k_trace_rays(DevScene S, const float* __restrict__ rays6, long long n,
                                                                         const int* __restrict__ ignore2, int* __restrict__ objId,
                                                                         int* __restrict__ meshId, int* __restrict__ triId,
                                                                         float* __restrict__ u, float* __restrict__ v, float* __restrict__ d,
                                                                         float* __restrict__ pos3, unsigned long long* __restrict__ counters) {
    Ctr ctr;
    ctr.nBox = ctr.nTriRef = ctr.nTriCulled = ctr.nTriFull = 0;
    unsigned nHits = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        V3 ro = mk(rays6[6 * i], rays6[6 * i + 1], rays6[6 * i + 2]);
        V3 rd = mk(rays6[6 * i + 3], rays6[6 * i + 4], rays6[6 * i + 5]);
        int im = -1, it = -1;
        if (ignore2) { im = ignore2[2 * i]; it = ignore2[2 * i + 1]; }
        Hit h;
        h.obj = h.mesh = h.tri = -1; h.key = 0; h.u = h.v = h.d = 0.0f; h.pos = mk(0.0f, 0.0f, 0.0f);
        const bool ok = traceScene<COUNT, false, MONO>(S, ro, rd, im, it, h, ctr);
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
                        long long upperBound, int maxBlocks, bool count, int kind, cudaStream_t st) {
    int grid = gridFor(upperBound, maxBlocks);
    if (kind == 3) {          // MonoGame slab form: one generic build (the instrumented counters are those of the XNA 4.0 walk)
        k_level<false, XN_LEVEL_MIN_BLOCKS, false, 256, true><<<grid, 256, 0, st>>>(S, F, A, lights, level);
    } else if (kind == 2) {          // deep octrees: high-occupancy build
        if (count) k_level<true, XN_DEEP_MIN_BLOCKS, false><<<grid, 256, 0, st>>>(S, F, A, lights, level);
        else k_level<false, XN_DEEP_MIN_BLOCKS, false><<<grid, 256, 0, st>>>(S, F, A, lights, level);
    } else if (kind == 1) {   // octrees with one split
        if (count) k_level<true, XN_LEVEL_MIN_BLOCKS, false><<<grid, 256, 0, st>>>(S, F, A, lights, level);
        else k_level<false, XN_LEVEL_MIN_BLOCKS, false><<<grid, 256, 0, st>>>(S, F, A, lights, level);
    } else {                  // every MeshOctree a single leaf: no expansion code, no stack
        const int g2 = (int)std::min<long long>((long long)grid * 256 / XN_LEAF_THREADS, 0x7fffffff);
        if (count) k_level<true, XN_LEAF_MIN_BLOCKS, true, XN_LEAF_THREADS><<<g2, XN_LEAF_THREADS, 0, st>>>(S, F, A, lights, level);
        else k_level<false, XN_LEAF_MIN_BLOCKS, true, XN_LEAF_THREADS><<<g2, XN_LEAF_THREADS, 0, st>>>(S, F, A, lights, level);
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

cudaError_t launchResolveChain(const DevFrame& F, const DevArena& A, int maxBlocks, uint32_t* out, cudaStream_t st) {
    k_resolve_chain<<<gridFor(F.chunkPixels, maxBlocks), 256, 0, st>>>(F, A, out);
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
                            int maxBlocks, bool count, bool mono, cudaStream_t st) {
    int grid = gridFor(n, maxBlocks);
    if (mono) k_trace_rays<false, true><<<grid, 256, 0, st>>>(S, rays6, n, ignore2, objId, meshId, triId, u, v, d, pos3, counters);
    else if (count) k_trace_rays<true><<<grid, 256, 0, st>>>(S, rays6, n, ignore2, objId, meshId, triId, u, v, d, pos3, counters);
    else k_trace_rays<false><<<grid, 256, 0, st>>>(S, rays6, n, ignore2, objId, meshId, triId, u, v, d, pos3, counters);
    return cudaGetLastError();
}

}  // namespace xn
