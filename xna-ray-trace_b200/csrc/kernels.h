// kernels.h — launch wrappers of the trace kernels (kernels.cu: one thread per ray; wavefront.cu:
// persistent trace kernel + streaming shade kernels for deep octrees).
#pragma once
#include <cuda_runtime.h>

#include "../../include/xnart.h"
#include "device_types.h"

namespace xn {

// One wave of the trace loop: closest hit + shadow rays + shading for every ray of `level`.
// upperBound only sizes the grid (the real count is read on the device).  kind: 0 = every MeshOctree is a
// single leaf (build without expansion code), 1 = shallow octrees, 2 = deep octrees (64-register build), 3 = any scene with the
// MonoGame form of the slab test (XNART_SEMANTICS_MONOGAME).
cudaError_t launchLevel(const DevScene& S, const DevFrame& F, const DevArena& A, const DevLight* lights, int level,
                        long long upperBound, int maxBlocks, bool count, int kind, cudaStream_t st);

// The whole frame (all levels + resolve) in one cooperative launch; `blocks` must not exceed
// frameKernelMaxBlocks().  stamps = 3 x uint64 device words (global timer at the phase boundaries).
int frameKernelMaxBlocks(bool count);
cudaError_t launchFrame(const DevScene& S, const DevFrame& F, const DevArena& A, const DevLight* lights, uint32_t* out,
                        unsigned long long* stamps, int blocks, bool count, cudaStream_t st);

// Bottom-up colour resolve of one level; level 0 writes the frame.
cudaError_t launchResolve(const DevFrame& F, const DevArena& A, int level, long long upperBound, int maxBlocks,
                          uint32_t* out, cudaStream_t st);

// Opaque scenes (every path a chain): the whole bottom-up resolve in one launch, one thread per pixel.
cudaError_t launchResolveChain(const DevFrame& F, const DevArena& A, int maxBlocks, uint32_t* out, cudaStream_t st);

// Batch closest-hit queries (xnart_trace_rays).
cudaError_t launchTraceRays(const DevScene& S, const float* rays6, long long n, const int* ignore2, int* objId, int* meshId,
                            int* triId, float* u, float* v, float* d, float* pos3, unsigned long long* counters,
                            int maxBlocks, bool count, bool mono, cudaStream_t st);

// ---- small opaque scenes (path.cu): the whole frame in one persistent launch -----------------
int pathKernelBlocksPerSm(bool count);
// fetch: one zeroed device word (work counter).
cudaError_t launchPath(const SmallScene& SS, const DevScene& S, const DevFrame& F, const DevLight* lights, uint32_t* out,
                       unsigned* fetch, unsigned long long* counters, int blocks, bool count, cudaStream_t st);

// ---- wavefront path (wavefront.cu) ----------------------------------------------------------
int traceKernelBlocksPerSm(bool count);   // resident blocks of the persistent trace kernel
cudaError_t launchGenPrimary(const DevFrame& F, const DevArena& A, int maxBlocks, cudaStream_t st);
// One batch of GetRayIntersection queries through the persistent kernel (`blocks` resident blocks).
cudaError_t launchTrace(const DevScene& S, const TraceJob& J, unsigned long long* counters, int blocks, bool count, cudaStream_t st);
cudaError_t launchShadeA(const DevScene& S, const DevFrame& F, const DevArena& A, const DevLight* lights, int level,
                         long long upperBound, int maxBlocks, cudaStream_t st);
cudaError_t launchShadeB(const DevScene& S, const DevFrame& F, const DevArena& A, const DevLight* lights, int level,
                         long long upperBound, int maxBlocks, cudaStream_t st);
cudaError_t launchPackRays(const float* rays6, const int* ignore2, long long n, float4* rayO, float4* rayD, int2* rayIgn,
                           int maxBlocks, cudaStream_t st);
cudaError_t launchUnpackHits(const DevScene& S, const int4* hit, long long n, int* objId, int* meshId, int* triId, float* u, float* v,
                             float* d, float* pos3, unsigned long long* counters, int maxBlocks, cudaStream_t st);

}  // namespace xn
