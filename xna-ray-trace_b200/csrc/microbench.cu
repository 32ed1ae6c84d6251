// microbench.cu — measured ceilings for the roofline that bench.py reports (hand-written CUDA, sm_100a).
//
// The trace kernels never get near HBM bandwidth (ncu: DRAM 0.04-5 % busy, profiles/): what they run out of is
// issue slots for un-fused FP32 / compare / select instructions and, on deep octrees, L1 / L2 bandwidth for
// divergent 32-byte loads.  MEASURED_PEAKS.json only holds the HBM copy bandwidth and the bf16 tensor peak, so
// the other ceilings are measured here, on the same GPU and in the same process as the bench:
//   XNART_PEAK_FP32_NOFMA   FMUL + FADD streams without contraction (the library is built with -fmad=false:
//                           the reference's arithmetic is one rounding per operator), in operations/s.
//                           Nominal: SMs x 128 lanes x clock.
//   XNART_PEAK_L1_READ      every block re-reads its own 16 KB window through L1 (ld.global.ca, 16-byte loads), bytes/s
//   XNART_PEAK_L2_READ      the grid re-reads a 32 MB buffer with L1 bypassed (ld.global.cg), bytes/s
//   XNART_PEAK_HBM_READ     the grid streams a 4 GB buffer once per repetition (larger than the 126 MB L2), bytes/s
// Each figure is the best of `reps` timed launches (CUDA events on the launching stream) after a warm-up launch.
#include <algorithm>
#include <cstdint>

#include "../../include/xnart.h"

namespace {

constexpr int kChains = 8;

__global__ void __launch_bounds__(256) k_peak_fp32(float* __restrict__ out, int iters, float a, float b) {
    float x[kChains];
#pragma unroll
    for (int c = 0; c < kChains; c++) x[c] = (float)(threadIdx.x + c) * 1e-3f;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int c = 0; c < kChains; c++) x[c] = __fadd_rn(__fmul_rn(x[c], a), b);   // never contracted
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < kChains; c++) s += x[c];
    if (s == 123.456f) out[0] = s;   // keeps the chains alive, never true in practice
}

__device__ __forceinline__ float4 ldca(const float4* p) {
    float4 v;
    asm volatile("ld.global.ca.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ldcg(const float4* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// window = float4 elements a block cycles through (power of two); L1: per-block window, L2/HBM: the whole buffer
template <bool L1>
__global__ void __launch_bounds__(256) k_peak_read(const float4* __restrict__ buf, size_t elems, int iters, float* __restrict__ out) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (L1) {
        const size_t window = 1024;   // 16 KB per block
        const float4* base = buf + ((size_t)blockIdx.x * window) % elems;
#pragma unroll 1
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                // the window is re-read from a different rotation every iteration (a loop-invariant address would let the
                // compiler hoist the load)
                const float4 v = ldca(base + ((threadIdx.x + 256 * u + 32 * i) & (window - 1)));
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
    } else {
        const size_t stride = (size_t)gridDim.x * blockDim.x;
        size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll 1
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float4 v = ldcg(buf + idx);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                idx += stride;
                if (idx >= elems) idx -= elems;
            }
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) out[0] = acc.x;
}

}  // namespace

extern "C" int32_t xnart_measure_peak(int32_t device, int32_t which, int32_t reps, double* unitsPerSecond) {
    if (!unitsPerSecond || reps < 1 || which < XNART_PEAK_FP32_NOFMA || which > XNART_PEAK_HBM_READ) return XNART_ERR_INVALID_ARGUMENT;
    *unitsPerSecond = 0.0;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return XNART_ERR_NO_DEVICE; }
    if (device < 0 || device >= n || cudaSetDevice(device) != cudaSuccess) return XNART_ERR_INVALID_ARGUMENT;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaStream_t st;
    cudaEvent_t e0, e1;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return XNART_ERR_CUDA;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const size_t bytes = which == XNART_PEAK_HBM_READ ? ((size_t)4 << 30) : ((size_t)32 << 20);
    const size_t elems = bytes / 16;
    float4* buf = nullptr;
    float* out = nullptr;
    int rc = XNART_OK;
    if (cudaMalloc(&out, 64) != cudaSuccess) rc = XNART_ERR_OUT_OF_MEMORY;
    if (rc == XNART_OK && which != XNART_PEAK_FP32_NOFMA) {
        if (cudaMalloc(&buf, bytes) != cudaSuccess) rc = XNART_ERR_OUT_OF_MEMORY;
        else cudaMemsetAsync(buf, 0, bytes, st);
    }
    double best = 0.0;
    const int blocks = sms * 8;
    for (int r = 0; rc == XNART_OK && r <= reps; r++) {   // launch 0 is the warm-up
        double units = 0.0;
        cudaEventRecord(e0, st);
        if (which == XNART_PEAK_FP32_NOFMA) {
            const int iters = 4096;
            k_peak_fp32<<<blocks, 256, 0, st>>>(out, iters, 1.0000001f, 1e-7f);
            units = (double)blocks * 256.0 * iters * 4.0 * kChains * 2.0;
        } else if (which == XNART_PEAK_L1_READ) {
            const int iters = 2048;
            k_peak_read<true><<<blocks, 256, 0, st>>>(buf, elems, iters, out);
            units = (double)blocks * 256.0 * iters * 4.0 * 16.0;
        } else {
            // L2: 32 MB read 16 times over; HBM: 4 GB read once
            const double passes = which == XNART_PEAK_L2_READ ? 16.0 : 1.0;
            const int iters = (int)std::max(1.0, passes * (double)elems / ((double)blocks * 256.0 * 4.0));
            k_peak_read<false><<<blocks, 256, 0, st>>>(buf, elems, iters, out);
            units = (double)blocks * 256.0 * iters * 4.0 * 16.0;
        }
        cudaEventRecord(e1, st);
        if (cudaEventSynchronize(e1) != cudaSuccess || cudaGetLastError() != cudaSuccess) { rc = XNART_ERR_CUDA; break; }
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms > 0.0f) best = std::max(best, units / ((double)ms * 1e-3));
    }
    cudaFree(buf);
    cudaFree(out);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    *unitsPerSecond = best;
    return rc;
}
