// oracle/xna_math.h — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of the closed-source Microsoft.Xna.Framework 4.0 math used on the trace
// path.  The assembly is not under /root/reference
// (RayTraceProject.csproj:65 links Microsoft.Xna.Framework, Version=4.0.0.0, x86), so these
// are restated from its published behaviour; PARITY UNPINNED for everything in this file.
// Arithmetic rule: strict IEEE-754 binary32 per C# operator, evaluated exactly
// as parenthesised, no FMA contraction — compile with -ffp-contract=off, no -ffast-math.
#pragma once
#include <cmath>
#include <cfloat>
#include <cstdint>

namespace xo {

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
struct M4 { float m[16]; };  // row-major: m[4*r+c] = M(r+1)(c+1)
struct Box { V3 mn, mx; };
struct Ray { V3 p, d; };

// Vector3 operators: component-wise (A.1).
static inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 mul(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 mul(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
// Vector3 / float: reciprocal then multiply.
static inline V3 divs(V3 a, float s) { float f = 1.0f / s; return {a.x * f, a.y * f, a.z * f}; }
static inline float dot(V3 a, V3 b) { return ((a.x * b.x) + (a.y * b.y)) + (a.z * b.z); }
static inline V3 cross(V3 a, V3 b) {
    return {(a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)};
}
static inline float lengthSq(V3 v) { return ((v.x * v.x) + (v.y * v.y)) + (v.z * v.z); }
// (float)Math.Sqrt((double)s): sqrt of a float in double, rounded to float == correctly
// rounded float sqrt (double has > 2*24+2 bits).
static inline float length(V3 v) { return (float)std::sqrt((double)lengthSq(v)); }
static inline V3 normalize(V3 v) {
    float f = 1.0f / (float)std::sqrt((double)lengthSq(v));
    return {v.x * f, v.y * f, v.z * f};
}
static inline V3 reflect(V3 v, V3 n) {
    float d = dot(v, n);
    return {v.x - ((2.0f * d) * n.x), v.y - ((2.0f * d) * n.y), v.z - ((2.0f * d) * n.z)};
}
static inline V3 lerp(V3 a, V3 b, float t) {
    return {a.x + ((b.x - a.x) * t), a.y + ((b.y - a.y) * t), a.z + ((b.z - a.z) * t)};
}
// MathHelper.Min/Max and Vector3.Min/Max: (a < b) ? a : b  /  (a > b) ? a : b.
static inline float fmin_xna(float a, float b) { return (a < b) ? a : b; }
static inline float fmax_xna(float a, float b) { return (a > b) ? a : b; }
static inline V3 vmin(V3 a, V3 b) { return {fmin_xna(a.x, b.x), fmin_xna(a.y, b.y), fmin_xna(a.z, b.z)}; }
static inline V3 vmax(V3 a, V3 b) { return {fmax_xna(a.x, b.x), fmax_xna(a.y, b.y), fmax_xna(a.z, b.z)}; }

// Vector3.Transform(p, M): row vector times matrix, translation row added last.
static inline V3 transform(V3 p, const M4& M) {
    const float* m = M.m;
    return {(((p.x * m[0]) + (p.y * m[4])) + (p.z * m[8])) + m[12],
            (((p.x * m[1]) + (p.y * m[5])) + (p.z * m[9])) + m[13],
            (((p.x * m[2]) + (p.y * m[6])) + (p.z * m[10])) + m[14]};
}

static inline M4 identity() {
    M4 r{};
    r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f;
    return r;
}

// Matrix.Multiply: C(r,c) = (((A(r,1)*B(1,c)) + (A(r,2)*B(2,c))) + (A(r,3)*B(3,c))) + (A(r,4)*B(4,c)).
static inline M4 multiply(const M4& A, const M4& B) {
    M4 C;
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++)
            C.m[4 * r + c] = (((A.m[4 * r + 0] * B.m[0 + c]) + (A.m[4 * r + 1] * B.m[4 + c])) +
                              (A.m[4 * r + 2] * B.m[8 + c])) + (A.m[4 * r + 3] * B.m[12 + c]);
    return C;
}

// Matrix.Invert: cofactor expansion in the framework's evaluation order (A.1).
static inline M4 invert(const M4& M) {
    const float a11 = M.m[0], a12 = M.m[1], a13 = M.m[2], a14 = M.m[3];
    const float a21 = M.m[4], a22 = M.m[5], a23 = M.m[6], a24 = M.m[7];
    const float a31 = M.m[8], a32 = M.m[9], a33 = M.m[10], a34 = M.m[11];
    const float a41 = M.m[12], a42 = M.m[13], a43 = M.m[14], a44 = M.m[15];
    // 2x2 minors of rows 3,4
    const float s34_34 = (a33 * a44) - (a34 * a43);
    const float s34_24 = (a32 * a44) - (a34 * a42);
    const float s34_23 = (a32 * a43) - (a33 * a42);
    const float s34_14 = (a31 * a44) - (a34 * a41);
    const float s34_13 = (a31 * a43) - (a33 * a41);
    const float s34_12 = (a31 * a42) - (a32 * a41);
    const float c1 = ((a22 * s34_34) - (a23 * s34_24)) + (a24 * s34_23);
    const float c2 = -(((a21 * s34_34) - (a23 * s34_14)) + (a24 * s34_13));
    const float c3 = ((a21 * s34_24) - (a22 * s34_14)) + (a24 * s34_12);
    const float c4 = -(((a21 * s34_23) - (a22 * s34_13)) + (a23 * s34_12));
    const float inv = 1.0f / ((((a11 * c1) + (a12 * c2)) + (a13 * c3)) + (a14 * c4));
    M4 R;
    R.m[0] = c1 * inv;
    R.m[4] = c2 * inv;
    R.m[8] = c3 * inv;
    R.m[12] = c4 * inv;
    R.m[1] = -(((a12 * s34_34) - (a13 * s34_24)) + (a14 * s34_23)) * inv;
    R.m[5] = (((a11 * s34_34) - (a13 * s34_14)) + (a14 * s34_13)) * inv;
    R.m[9] = -(((a11 * s34_24) - (a12 * s34_14)) + (a14 * s34_12)) * inv;
    R.m[13] = (((a11 * s34_23) - (a12 * s34_13)) + (a13 * s34_12)) * inv;
    // 2x2 minors of rows 2,4
    const float s24_34 = (a23 * a44) - (a24 * a43);
    const float s24_24 = (a22 * a44) - (a24 * a42);
    const float s24_23 = (a22 * a43) - (a23 * a42);
    const float s24_14 = (a21 * a44) - (a24 * a41);
    const float s24_13 = (a21 * a43) - (a23 * a41);
    const float s24_12 = (a21 * a42) - (a22 * a41);
    R.m[2] = (((a12 * s24_34) - (a13 * s24_24)) + (a14 * s24_23)) * inv;
    R.m[6] = -(((a11 * s24_34) - (a13 * s24_14)) + (a14 * s24_13)) * inv;
    R.m[10] = (((a11 * s24_24) - (a12 * s24_14)) + (a14 * s24_12)) * inv;
    R.m[14] = -(((a11 * s24_23) - (a12 * s24_13)) + (a13 * s24_12)) * inv;
    // 2x2 minors of rows 2,3
    const float s23_34 = (a23 * a34) - (a24 * a33);
    const float s23_24 = (a22 * a34) - (a24 * a32);
    const float s23_23 = (a22 * a33) - (a23 * a32);
    const float s23_14 = (a21 * a34) - (a24 * a31);
    const float s23_13 = (a21 * a33) - (a23 * a31);
    const float s23_12 = (a21 * a32) - (a22 * a31);
    R.m[3] = -(((a12 * s23_34) - (a13 * s23_24)) + (a14 * s23_23)) * inv;
    R.m[7] = (((a11 * s23_34) - (a13 * s23_14)) + (a14 * s23_13)) * inv;
    R.m[11] = -(((a11 * s23_24) - (a12 * s23_14)) + (a14 * s23_12)) * inv;
    R.m[15] = (((a11 * s23_23) - (a12 * s23_13)) + (a13 * s23_12)) * inv;
    return R;
}

// Matrix.CreateLookAt (A.1).
static inline M4 createLookAt(V3 eye, V3 target, V3 up) {
    V3 z = normalize(sub(eye, target));
    V3 x = normalize(cross(up, z));
    V3 y = cross(z, x);
    M4 r{};
    r.m[0] = x.x; r.m[1] = y.x; r.m[2] = z.x; r.m[3] = 0.0f;
    r.m[4] = x.y; r.m[5] = y.y; r.m[6] = z.y; r.m[7] = 0.0f;
    r.m[8] = x.z; r.m[9] = y.z; r.m[10] = z.z; r.m[11] = 0.0f;
    r.m[12] = -dot(x, eye); r.m[13] = -dot(y, eye); r.m[14] = -dot(z, eye); r.m[15] = 1.0f;
    return r;
}

// Matrix.CreatePerspectiveFieldOfView (A.1).
static inline M4 createPerspectiveFov(float fov, float aspect, float zn, float zf) {
    M4 r{};
    float yScale = 1.0f / (float)std::tan((double)(fov * 0.5f));
    float xScale = yScale / aspect;
    r.m[0] = xScale;
    r.m[5] = yScale;
    r.m[10] = zf / (zn - zf);
    r.m[11] = -1.0f;
    r.m[14] = (zn * zf) / (zn - zf);
    return r;
}

struct Viewport { int x, y, w, h; float minDepth, maxDepth; };

// Viewport.Unproject with the matrix already inverted (the framework re-inverts
// world*view*proj per call; the product and inverse only depend on the frame).
static inline V3 unprojectWithInverse(const Viewport& vp, V3 src, const M4& inv) {
    V3 s;
    s.x = (((src.x - (float)vp.x) / (float)vp.w) * 2.0f) - 1.0f;
    s.y = -((((src.y - (float)vp.y) / (float)vp.h) * 2.0f) - 1.0f);
    s.z = (src.z - vp.minDepth) / (vp.maxDepth - vp.minDepth);
    V3 v = transform(s, inv);
    float a = (((s.x * inv.m[3]) + (s.y * inv.m[7])) + (s.z * inv.m[11])) + inv.m[15];
    float num = a - 1.0f;
    bool within = (-1.401298E-45f <= num) && (num <= 1.401298E-45f);
    if (!within) v = divs(v, a);
    return v;
}

// BoundingBox.Contains(Vector3) != Disjoint : inclusive on all faces.
static inline bool boxContains(const Box& b, V3 p) {
    return !(b.mn.x > p.x || p.x > b.mx.x || b.mn.y > p.y || p.y > b.mx.y || b.mn.z > p.z || p.z > b.mx.z);
}
// BoundingBox.Intersects(BoundingBox): inclusive overlap.
static inline bool boxIntersects(const Box& a, const Box& b) {
    if (a.mx.x < b.mn.x || a.mn.x > b.mx.x) return false;
    if (a.mx.y < b.mn.y || a.mn.y > b.mx.y) return false;
    if (a.mx.z < b.mn.z || a.mn.z > b.mx.z) return false;
    return true;
}
static inline Box boxMerge(const Box& a, const Box& b) { return {vmin(a.mn, b.mn), vmax(a.mx, b.mx)}; }

// BoundingBox.Intersects(ref Ray, out float?) — XNA 4.0 form (A.1): reciprocal-multiply,
// tmin starts at 0, returns tmin.  Returns false for null.
static inline bool slabXna40(const Box& b, const Ray& r, float& tOut) {
    float tmin = 0.0f, tmax = FLT_MAX;
    const float mn[3] = {b.mn.x, b.mn.y, b.mn.z}, mx[3] = {b.mx.x, b.mx.y, b.mx.z};
    const float p[3] = {r.p.x, r.p.y, r.p.z}, d[3] = {r.d.x, r.d.y, r.d.z};
    for (int a = 0; a < 3; a++) {
        if (std::fabs(d[a]) < 1e-6f) {
            if (p[a] < mn[a] || p[a] > mx[a]) return false;
        } else {
            float inv = 1.0f / d[a];
            float t1 = (mn[a] - p[a]) * inv;
            float t2 = (mx[a] - p[a]) * inv;
            if (t1 > t2) { float t = t1; t1 = t2; t2 = t; }
            tmin = fmax_xna(t1, tmin);
            tmax = fmin_xna(t2, tmax);
            if (tmin > tmax) return false;
        }
    }
    tOut = tmin;
    return true;
}

// MonoGame form (A.1): divides, nullable tMin/tMax, 0 when the origin is inside.
static inline bool slabMonoGame(const Box& b, const Ray& r, float& tOut) {
    bool hasMin = false, hasMax = false;
    float tMin = 0.0f, tMax = 0.0f;
    const float mn[3] = {b.mn.x, b.mn.y, b.mn.z}, mx[3] = {b.mx.x, b.mx.y, b.mx.z};
    const float p[3] = {r.p.x, r.p.y, r.p.z}, d[3] = {r.d.x, r.d.y, r.d.z};
    for (int a = 0; a < 3; a++) {
        if (std::fabs(d[a]) < 1e-6f) {
            if (p[a] < mn[a] || p[a] > mx[a]) return false;
        } else {
            float t1 = (mn[a] - p[a]) / d[a];
            float t2 = (mx[a] - p[a]) / d[a];
            if (t1 > t2) { float t = t1; t1 = t2; t2 = t; }
            if (a == 0) {
                tMin = t1; tMax = t2; hasMin = hasMax = true;
            } else {
                if ((hasMin && tMin > t2) || (hasMax && t1 > tMax)) return false;
                if (!hasMin || t1 > tMin) { tMin = t1; hasMin = true; }
                if (!hasMax || t2 < tMax) { tMax = t2; hasMax = true; }
            }
        }
    }
    // lifted comparisons on null are false
    if ((hasMin && tMin < 0.0f) && (hasMax && tMax > 0.0f)) { tOut = 0.0f; return true; }
    if (hasMin && tMin < 0.0f) return false;
    if (!hasMin) return false;  // "return tMin" with tMin == null
    tOut = tMin;
    return true;
}

// Color(Vector3) channel pack (A.2).  XNA 4.0: PackUNorm = ClampAndRound(v*255) with
// Math.Round (half to even).  MonoGame: (int)(v*255) truncation, then clamp to byte range.
static inline uint32_t packChannelXna40(float v) {
    float t = v * 255.0f;
    if (std::isnan(t)) return 0u;
    if (std::isinf(t)) return t < 0.0f ? 0u : 255u;
    if (t < 0.0f) return 0u;
    if (t > 255.0f) return 255u;
    return (uint32_t)std::nearbyint((double)t);  // default rounding mode = half to even
}
static inline uint32_t packChannelMonoGame(float v) {
    float t = v * 255.0f;
    if (std::isnan(t)) return 0u;
    if (t <= 0.0f) return 0u;
    if (t >= 255.0f) return 255u;
    return (uint32_t)(int32_t)t;
}
static inline uint32_t packColor(V3 c, int semantics) {
    uint32_t r, g, b;
    if (semantics == 1) { r = packChannelMonoGame(c.x); g = packChannelMonoGame(c.y); b = packChannelMonoGame(c.z); }
    else { r = packChannelXna40(c.x); g = packChannelXna40(c.y); b = packChannelXna40(c.z); }
    return r | (g << 8) | (b << 16) | (255u << 24);
}
// Color.ToVector3(): byte / 255f (division).
static inline V3 unpackColor(uint32_t c) {
    return {(float)(c & 255u) / 255.0f, (float)((c >> 8) & 255u) / 255.0f, (float)((c >> 16) & 255u) / 255.0f};
}

}  // namespace xo
