// xna_math.cuh — XNA 4.0 value-type arithmetic for the trace path, host + device.
//
// Every function evaluates in IEEE binary32, one rounding per C# operator, in the order the
// framework parenthesises it.  The translation unit must be compiled
// with -fmad=false (device) and -ffp-contract=off (host) so that no multiply-add is fused;
// divisions and square roots are the IEEE-rounded forms (nvcc defaults -prec-div/-prec-sqrt).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

#define XD __host__ __device__ __forceinline__

namespace xn {

struct V3 { float x, y, z; };

XD V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
XD V3 add(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
XD V3 sub(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
XD V3 mul(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
XD V3 mulv(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
XD V3 neg(V3 a) { return mk(-a.x, -a.y, -a.z); }
XD float dot(V3 a, V3 b) { return ((a.x * b.x) + (a.y * b.y)) + (a.z * b.z); }
XD V3 cross(V3 a, V3 b) {
    return mk((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
XD float lengthSq(V3 v) { return ((v.x * v.x) + (v.y * v.y)) + (v.z * v.z); }
// (float)Math.Sqrt((double)s) == correctly rounded binary32 square root
XD float length(V3 v) { return sqrtf(lengthSq(v)); }
XD V3 normalize(V3 v) {
    float f = 1.0f / sqrtf(lengthSq(v));
    return mk(v.x * f, v.y * f, v.z * f);
}
XD V3 reflect(V3 v, V3 n) {
    float d = dot(v, n);
    return mk(v.x - ((2.0f * d) * n.x), v.y - ((2.0f * d) * n.y), v.z - ((2.0f * d) * n.z));
}
XD V3 lerp(V3 a, V3 b, float t) {
    return mk(a.x + ((b.x - a.x) * t), a.y + ((b.y - a.y) * t), a.z + ((b.z - a.z) * t));
}
// Vector3 / float: multiply by the reciprocal
XD V3 divs(V3 a, float s) { float f = 1.0f / s; return mk(a.x * f, a.y * f, a.z * f); }

// Vector3.Transform with a 4x3 matrix (the 4th column never participates): m[4*r + c], c < 3.
XD V3 transform43(V3 p, const float* m) {
    return mk((((p.x * m[0]) + (p.y * m[4])) + (p.z * m[8])) + m[12],
              (((p.x * m[1]) + (p.y * m[5])) + (p.z * m[9])) + m[13],
              (((p.x * m[2]) + (p.y * m[6])) + (p.z * m[10])) + m[14]);
}

// Matrix.Multiply, row-major 4x4.
inline void multiply44(const float* A, const float* B, float* C) {
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++)
            C[4 * r + c] = (((A[4 * r + 0] * B[0 + c]) + (A[4 * r + 1] * B[4 + c])) + (A[4 * r + 2] * B[8 + c])) +
                           (A[4 * r + 3] * B[12 + c]);
}

// Matrix.Invert: adjugate over determinant, each 3x3 cofactor built from 2x2 minors of the two
// lower rows in the framework's order: ((p*X) - (q*Y)) + (r*Z).
inline void invert44(const float* M, float* R) {
    auto minor2 = [&](int ra, int rb, int ca, int cb) { return (M[4 * ra + ca] * M[4 * rb + cb]) - (M[4 * ra + cb] * M[4 * rb + ca]); };
    auto cof = [](float p, float X, float q, float Y, float r, float Z) { return ((p * X) - (q * Y)) + (r * Z); };
    const float a11 = M[0], a12 = M[1], a13 = M[2], a14 = M[3];
    const float a21 = M[4], a22 = M[5], a23 = M[6], a24 = M[7];
    float s34 = minor2(2, 3, 2, 3), s24 = minor2(2, 3, 1, 3), s23 = minor2(2, 3, 1, 2);
    float s14 = minor2(2, 3, 0, 3), s13 = minor2(2, 3, 0, 2), s12 = minor2(2, 3, 0, 1);
    const float c1 = cof(a22, s34, a23, s24, a24, s23);
    const float c2 = -cof(a21, s34, a23, s14, a24, s13);
    const float c3 = cof(a21, s24, a22, s14, a24, s12);
    const float c4 = -cof(a21, s23, a22, s13, a23, s12);
    const float inv = 1.0f / ((((a11 * c1) + (a12 * c2)) + (a13 * c3)) + (a14 * c4));
    R[0] = c1 * inv; R[4] = c2 * inv; R[8] = c3 * inv; R[12] = c4 * inv;
    R[1] = -cof(a12, s34, a13, s24, a14, s23) * inv;
    R[5] = cof(a11, s34, a13, s14, a14, s13) * inv;
    R[9] = -cof(a11, s24, a12, s14, a14, s12) * inv;
    R[13] = cof(a11, s23, a12, s13, a13, s12) * inv;
    s34 = minor2(1, 3, 2, 3); s24 = minor2(1, 3, 1, 3); s23 = minor2(1, 3, 1, 2);
    s14 = minor2(1, 3, 0, 3); s13 = minor2(1, 3, 0, 2); s12 = minor2(1, 3, 0, 1);
    R[2] = cof(a12, s34, a13, s24, a14, s23) * inv;
    R[6] = -cof(a11, s34, a13, s14, a14, s13) * inv;
    R[10] = cof(a11, s24, a12, s14, a14, s12) * inv;
    R[14] = -cof(a11, s23, a12, s13, a13, s12) * inv;
    s34 = minor2(1, 2, 2, 3); s24 = minor2(1, 2, 1, 3); s23 = minor2(1, 2, 1, 2);
    s14 = minor2(1, 2, 0, 3); s13 = minor2(1, 2, 0, 2); s12 = minor2(1, 2, 0, 1);
    R[3] = -cof(a12, s34, a13, s24, a14, s23) * inv;
    R[7] = cof(a11, s34, a13, s14, a14, s13) * inv;
    R[11] = -cof(a11, s24, a12, s14, a14, s12) * inv;
    R[15] = cof(a11, s23, a12, s13, a13, s12) * inv;
}

// MathHelper.Max / Min: (a > b) ? a : b  /  (a < b) ? a : b
XD float maxx(float a, float b) { return (a > b) ? a : b; }
XD float minx(float a, float b) { return (a < b) ? a : b; }

// Color(Vector3) channel, XNA 4.0: ClampAndRound(v * 255) with Math.Round (half to even).
XD uint32_t packChannelXna40(float v) {
    float t = v * 255.0f;
    if (!(t == t)) return 0u;
    if (t < 0.0f) return 0u;
    if (t > 255.0f) return 255u;  // also +inf
#ifdef __CUDA_ARCH__
    return (uint32_t)__float2int_rn(t);
#else
    return (uint32_t)nearbyintf(t);
#endif
}
// MonoGame: (int)(v * 255), then the int constructor clamps to the byte range.
XD uint32_t packChannelMonoGame(float v) {
    float t = v * 255.0f;
    if (!(t == t)) return 0u;
    if (t <= 0.0f) return 0u;
    if (t >= 255.0f) return 255u;
    return (uint32_t)(int32_t)t;
}
XD uint32_t packColor(V3 c, int semantics) {
    uint32_t r, g, b;
    if (semantics == 1) { r = packChannelMonoGame(c.x); g = packChannelMonoGame(c.y); b = packChannelMonoGame(c.z); }
    else { r = packChannelXna40(c.x); g = packChannelXna40(c.y); b = packChannelXna40(c.z); }
    return r | (g << 8) | (b << 16) | 0xFF000000u;
}
// byte / 255f.  On the device a zero dividend takes the IEEE division's slow-path subroutine (ncu: 177 k warp-level calls
// per crate frame, black being the colour of every miss), so 0 is divided as 1 and selected back to 0.0f = 0 / 255.
XD float byteToUnit(uint32_t b) {
#ifdef __CUDA_ARCH__
    const float q = (float)(b | (uint32_t)(b == 0u)) / 255.0f;
    return b ? q : 0.0f;
#else
    return (float)b / 255.0f;
#endif
}
// Color.ToVector3(): byte / 255f
XD V3 unpackColor(uint32_t c) {
    return mk(byteToUnit(c & 255u), byteToUnit((c >> 8) & 255u), byteToUnit((c >> 16) & 255u));
}

}  // namespace xn
