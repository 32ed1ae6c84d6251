// frame_ops.cu — finished frames on their way to the AviFile-style frame sink (hand-written CUDA, sm_100a).
//
// VideoStream.AddFrame (aviFileWrapper_src/VideoStream.cs:344-365) flips the Bitmap (RotateNoneFlipY, :345),
// locks its bits as 24bpp BGR rows padded to a multiple of 4 bytes and hands them to AVIStreamWrite.  A
// rendered frame is a Color[] (R | G<<8 | B<<16 | A<<24, top-down, RayTracer.cs:425); converting it on the GPU
// before the device->host copy moves 3 bytes per pixel over PCIe instead of 4 and leaves the host nothing to do
// but append the bytes to the file.
#include <cstdint>

#include "../../include/xnart.h"

namespace {

// One thread = 4 pixels of one row = 12 output bytes: one 16-byte load and three 32-bit stores when the rows
// are aligned for it (vec: W % 4 == 0, stride % 4 == 0, aligned base pointers), bytewise otherwise.
__global__ void __launch_bounds__(256) k_dib24(const uint32_t* __restrict__ src, uint8_t* __restrict__ dst, int W, int H, int stride,
                                              long long nFrames, bool vec) {
    const int quads = (W + 3) / 4;
    const long long total = nFrames * (long long)H * quads;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int qx = (int)(i % quads);
        const long long fy = i / quads;
        const int y = (int)(fy % H);
        const long long f = fy / H;
        const uint32_t* __restrict__ s = src + (f * H + y) * (long long)W + 4 * qx;
        uint8_t* __restrict__ d = dst + (f * H + (H - 1 - y)) * (long long)stride + 12 * qx;   // bottom-up rows
        if (vec) {
            const uint4 p = *reinterpret_cast<const uint4*>(s);   // rows start 16-byte aligned when W % 4 == 0
            // Color byte order in memory is R,G,B,A; the DIB wants B,G,R
            auto bgr = [](uint32_t c) { return ((c >> 16) & 255u) | (c & 0xFF00u) | ((c & 255u) << 16); };
            const uint32_t a = bgr(p.x), b = bgr(p.y), c = bgr(p.z), e = bgr(p.w);
            uint32_t* o = reinterpret_cast<uint32_t*>(d);
            o[0] = a | (b << 24);
            o[1] = (b >> 8) | (c << 16);
            o[2] = (c >> 16) | (e << 8);
        } else {
            for (int k = 0; k < 4 && 4 * qx + k < W; k++) {
                const uint32_t c = s[k];
                d[3 * k + 0] = (uint8_t)((c >> 16) & 255u);
                d[3 * k + 1] = (uint8_t)((c >> 8) & 255u);
                d[3 * k + 2] = (uint8_t)(c & 255u);
            }
            if (4 * qx + 4 >= W) for (int b = 3 * W; b < stride; b++) d[b - 12 * qx] = 0;   // row padding
        }
    }
}

}  // namespace

extern "C" int32_t xnart_frames_to_dib24(int32_t device, const void* devRgba8, int64_t nFrames, int32_t width, int32_t height,
                                         int32_t stride, void* devDib, void* cudaStream) {
    if (!devRgba8 || !devDib || nFrames < 0 || width <= 0 || height <= 0 || stride < 3 * width) return XNART_ERR_INVALID_ARGUMENT;
    if (nFrames == 0) return XNART_OK;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return XNART_ERR_CUDA; }
    const long long total = nFrames * (long long)height * ((width + 3) / 4);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    const bool vec = width % 4 == 0 && stride % 4 == 0 && ((uintptr_t)devRgba8 & 15) == 0 && ((uintptr_t)devDib & 3) == 0;
    k_dib24<<<(int)blocks, 256, 0, (cudaStream_t)cudaStream>>>((const uint32_t*)devRgba8, (uint8_t*)devDib, width, height, stride, nFrames, vec);
    return cudaGetLastError() == cudaSuccess ? XNART_OK : XNART_ERR_CUDA;
}
