// trace.cuh — device-side restatement of the reference's intersection and shading path.
//
//   slab test           XNA BoundingBox.Intersects(ref Ray, out float?)   (call sites
//                       MeshOctree.cs:331, OctreeSpatialManager.cs:460, Mesh.cs:37)
//   triangle test       RayExtensions.IntersectsTriangleBackfaceCulling    RayExtensions.cs:42-75
//   mesh closest hit    MeshOctree.GetRayIntersection                      MeshOctree.cs:259-353
//   scene closest hit   OctreeSpatialManager.GetRayIntersection            OctreeSpatialManager.cs:312-482
//   texture lookup      Material.LookupUV                                  Material.cs:71-232
//   lights              SpotLight / DirectionalLight.GetLightForFragment   SpotLight.cs:37-62, DirectionalLight.cs:23-30
//
// Traversal semantics.  The reference collects every leaf the ray's slab test accepts, sorts the
// leaves by the slab entry distance t, and walks groups of bit-equal t in ascending order until the
// first group that yields an accepted triangle (NOT a closest-hit search: quirk Q1).  That is
// order-independent:  K = min{ t(leaf) : leaf has an accepted triangle },  answer = the accepted
// triangle of minimum distance over the leaves with t == K, ties to the earliest position in the
// reference's visiting order (= smallest position in refs[], laid out in depth-first leaf order).
// So each thread runs a depth-first walk with a short private stack, children visited front to
// back by the ray's octant, and discards a subtree when its entry distance exceeds the current K.
// The entry distance of an internal node is taken on its conservative box (consMax) because a
// child's Max can overshoot the parent's by rounding; the decision to descend uses the node's own
// box exactly as the reference does.  Before the first accepted triangle K = +inf and nothing is
// pruned, so the visiting order never changes the answer.
//
// All of this file must be compiled with -fmad=false: the reference's arithmetic is binary32 with
// one rounding per operator.
#pragma once
#include <cooperative_groups.h>

#include "device_types.h"
#include "xna_math.cuh"

namespace xn {
namespace cg = cooperative_groups;

struct Ctr {
    unsigned nBox, nTriRef, nTriCulled, nTriFull;
};

struct RayPre {
    V3 o, d, inv;
    int par;    // bit a set: |d[a]| < 1e-6 (the slab test's parallel branch)
    int oct;    // bit 2/1/0 set: d.x / d.y / d.z negative -> children visited as (i ^ oct)
};

// 256-bit read-only load (LDG.E.256 on sm_100a): a divergent access costs the L1 one wavefront per
// lane and per INSTRUCTION, so the 128-byte node record and the 64-byte triangle record are fetched
// with 32-byte loads (ncu: l1tex__data_pipe_lsu_wavefronts was the binding unit with 16-byte loads).
struct __align__(32) F8 { float4 a, b; };
__device__ __forceinline__ F8 ldg256(const void* p) {
    F8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ bool finite3(V3 v) {
    return (fabsf(v.x) <= FLT_MAX) && (fabsf(v.y) <= FLT_MAX) && (fabsf(v.z) <= FLT_MAX);
}

__device__ __forceinline__ RayPre prepRay(V3 o, V3 d) {
    RayPre r;
    r.o = o; r.d = d;
    r.par = (fabsf(d.x) < 1e-6f ? 1 : 0) | (fabsf(d.y) < 1e-6f ? 2 : 0) | (fabsf(d.z) < 1e-6f ? 4 : 0);
    // float inv = 1f / ray.Direction.X, recomputed per box by the framework: ray-invariant
    r.inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    r.oct = (d.x < 0.0f ? 4 : 0) | (d.y < 0.0f ? 2 : 0) | (d.z < 0.0f ? 1 : 0);
    return r;
}

#ifndef XN_SLAB_FAST
#define XN_SLAB_FAST 1
#endif
// One axis of the slab test.  Returns false when the parallel branch rejects.
__device__ __forceinline__ bool slabAxis(float mn, float mx, float p, float inv, bool par, float& tmin, float& tmax) {
    if (par) {
        if (p < mn || p > mx) return false;
    } else {
        // The framework swaps with `if (t1 > t2)` and accumulates with MathHelper.Max / Min, i.e. compare + select.
        // For operands that are not NaN that is fminf / fmaxf (one FMNMX each instead of FSETP + FSEL) up to the
        // sign of a zero, which no later comparison distinguishes.  t1 / t2 cannot be NaN here: the ray is finite
        // (callers test finite3 first), |d| >= 1e-6 on this branch so inv is finite and non-zero, and the API
        // rejects NaN coordinates (xnart_add_mesh / xnart_add_object).
        const float t1 = (mn - p) * inv;
        const float t2 = (mx - p) * inv;
        tmin = fmaxf(fminf(t1, t2), tmin);
        tmax = fminf(fmaxf(t1, t2), tmax);
    }
    return true;
}

// BoundingBox.Intersects(ref Ray, out float?) — XNA 4.0 form.  tmin only grows and tmax only
// shrinks, so the framework's per-axis `if (tmin > tmax) return null` equals one test at the end.
__device__ __forceinline__ bool slabGeneral(float3 mn, float3 mx, const RayPre& r, float& t, float& tExit) {
    float tmin = 0.0f, tmax = FLT_MAX;
    bool ok = slabAxis(mn.x, mx.x, r.o.x, r.inv.x, r.par & 1, tmin, tmax);
    ok = slabAxis(mn.y, mx.y, r.o.y, r.inv.y, r.par & 2, tmin, tmax) && ok;
    ok = slabAxis(mn.z, mx.z, r.o.z, r.inv.z, r.par & 4, tmin, tmax) && ok;
    t = tmin;
    tExit = tmax;
    return ok && !(tmin > tmax);
}
__device__ __forceinline__ bool slab2(float3 mn, float3 mx, const RayPre& r, float& t, float& tExit) {
#if XN_SLAB_FAST
    if (r.par == 0) {
        float tmin = 0.0f, tmax = FLT_MAX;
        // no component below the 1e-6 threshold (almost every ray): the three non-parallel axes of slabAxis alone —
        // the predicated containment code of the general form was half of the test's instructions (ncu, profiles/r02k)
        const float x1 = (mn.x - r.o.x) * r.inv.x, x2 = (mx.x - r.o.x) * r.inv.x;
        tmin = fmaxf(fminf(x1, x2), tmin); tmax = fminf(fmaxf(x1, x2), tmax);
        const float y1 = (mn.y - r.o.y) * r.inv.y, y2 = (mx.y - r.o.y) * r.inv.y;
        tmin = fmaxf(fminf(y1, y2), tmin); tmax = fminf(fmaxf(y1, y2), tmax);
        const float z1 = (mn.z - r.o.z) * r.inv.z, z2 = (mx.z - r.o.z) * r.inv.z;
        tmin = fmaxf(fminf(z1, z2), tmin); tmax = fminf(fmaxf(z1, z2), tmax);
        t = tmin;
        tExit = tmax;
        return !(tmin > tmax);
    }
#endif
    return slabGeneral(mn, mx, r, t, tExit);
}
__device__ __forceinline__ bool slab(float3 mn, float3 mx, const RayPre& r, float& t) {
    float tExit;
    return slab2(mn, mx, r, t, tExit);
}

// BoundingBox.Intersects(ref Ray, out float?) — MonoGame form: divides by the direction instead of
// multiplying by its reciprocal (different last-bit rounding, hence different groups of bit-equal entry distances),
// nullable tMin / tMax set by the first non-parallel axis, per-axis rejection, and at the end: 0 when the origin lies
// between the planes (tMin < 0 < tMax), null when the box lies behind (tMin < 0), tMin otherwise.
__device__ __forceinline__ bool slabMonoGame(float3 mn, float3 mx, const RayPre& r, float& t) {
    bool has = false;
    float tMin = 0.0f, tMax = 0.0f;
    const float lo[3] = {mn.x, mn.y, mn.z}, hi[3] = {mx.x, mx.y, mx.z};
    const float p[3] = {r.o.x, r.o.y, r.o.z}, d[3] = {r.d.x, r.d.y, r.d.z};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        if (fabsf(d[a]) < 1e-6f) {
            if (p[a] < lo[a] || p[a] > hi[a]) return false;
        } else {
            float t1 = (lo[a] - p[a]) / d[a];
            float t2 = (hi[a] - p[a]) / d[a];
            if (t1 > t2) { const float s = t1; t1 = t2; t2 = s; }
            if (!has) {
                tMin = t1; tMax = t2; has = true;
            } else {
                if (tMin > t2 || t1 > tMax) return false;
                if (t1 > tMin) tMin = t1;
                if (t2 < tMax) tMax = t2;
            }
        }
    }
    if (!has) return false;                               // "return tMin" with tMin == null
    if (tMin < 0.0f && tMax > 0.0f) { t = 0.0f; return true; }
    if (tMin < 0.0f) return false;
    t = tMin;
    return true;
}
// The slab test of the framework the caller asked for (XnartFrame.semantics / XNART_OPT_SEMANTICS).
template <bool MONO>
__device__ __forceinline__ bool slabF(float3 mn, float3 mx, const RayPre& r, float& t) {
    if (MONO) return slabMonoGame(mn, mx, r, t);
    return slab(mn, mx, r, t);
}

// Vector3.Transform with the matrix fetched as four 16-byte rows (SceneObject.World / InverseWorld
// live in global memory: 4 loads instead of 12).
struct M43 { float4 r0, r1, r2, r3; };
__device__ __forceinline__ M43 loadM43(const float* __restrict__ m) {
    const float4* __restrict__ q = reinterpret_cast<const float4*>(m);
    M43 r;
    r.r0 = __ldg(&q[0]); r.r1 = __ldg(&q[1]); r.r2 = __ldg(&q[2]); r.r3 = __ldg(&q[3]);
    return r;
}
__device__ __forceinline__ V3 transform43(V3 p, const M43& m) {
    return mk((((p.x * m.r0.x) + (p.y * m.r1.x)) + (p.z * m.r2.x)) + m.r3.x,
              (((p.x * m.r0.y) + (p.y * m.r1.y)) + (p.z * m.r2.y)) + m.r3.y,
              (((p.x * m.r0.z) + (p.y * m.r1.z)) + (p.z * m.r2.z)) + m.r3.z);
}

// Mesh.RayIntersects (Mesh.cs:34-39): slab test on Mesh.MeshBoundingBox (bytes 0..23 of DevMesh).
template <bool MONO = false>
__device__ __forceinline__ bool meshBoxSlab(const DevMesh* __restrict__ M, const RayPre& r) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(M->box));
    const float2 b1 = __ldg(reinterpret_cast<const float2*>(M->box + 4));
    float tb;
    return slabF<MONO>(make_float3(b0.x, b0.y, b0.z), make_float3(b0.w, b1.x, b1.y), r, tb);
}

struct MeshHit {
    int tri;       // mesh-relative triangle index, -1 = none
    int key;       // position of the winning leaf reference in refs[] / leafTri[]
    float u, v, d;
};

struct StackEntry { int node; float t; };

// Mesh-walk stack entry: a child descriptor (see device_types.h, inner[8k+4..7]) and its entry distance.
// 16 bytes so that a push is ONE local store and a pop ONE local load (STL.128 / LDL.128).
typedef int4 WalkEntry;  // (a, b, bits(t), 0)

#ifndef XN_LEAF_CHUNK
#define XN_LEAF_CHUNK 16   // triangles tested per visit of the leaf phase (<= 32; longer leaves are resumed)
#endif
#ifndef XN_LEAF_SURVIVORS
#define XN_LEAF_SURVIVORS 10  // > 0: the leaf phase scans until this many triangles await the full test (see walkLeaf); 0: XN_LEAF_CHUNK triangles
                              // (1M-triangle grid, 1080p: 10.24 ms with a fixed 16-triangle chunk, 9.87 ms with 10 survivors)
#endif
#ifndef XN_LEAF_PAIR
#define XN_LEAF_PAIR 1   // the full test runs on two surviving triangles per trip (records requested together, tests interleaved): 9.88 -> 9.52 ms on the 1M grid
#endif
#ifndef XN_DESCEND_MAX
#define XN_DESCEND_MAX 4   // node expansions per visit of the descend phase
#endif

// State of one MeshOctree.GetRayIntersection walk (all scalars: lives in registers).
struct Walk {
    float K;            // entry distance of the nearest leaf that produced an accepted triangle
    float bd, bu, bv;   // best candidate of the current K group
    int btri, bkey;     // btri < 0: nothing accepted yet
    int ca, cb;         // held node: leaf (first ref, remaining count) or inner (record, -1)
    float ct;           // its entry distance
    bool have;
    int sp;
};

// The full part of IntersectsTriangleBackfaceCulling (RayExtensions.cs:52-75) for a triangle that
// passed the back-face test, and the closest-hit bookkeeping of MeshOctree.GetRayIntersection
// (MeshOctree.cs:286-304) restated per triangle.  The reference keeps, per group of leaves with a
// bit-equal entry distance, the accepted triangle of minimum distance, `distance < minDistance`
// with minDistance starting at float.MaxValue (:294), first of equal distances.  Here: an accepted
// triangle of a leaf entered at ct < K opens a new group (older candidates are void: reset to the
// MaxValue sentinel); inside a group candidates merge by (distance, position in refs[]) — positions
// increase in the reference's visiting order.
__device__ __forceinline__ void fullTriangleTest(const float4 nt, const float4 v1, const F8 B, int pos, const RayPre& r, Walk& w) {
    V3 E1 = mk(B.a.x, B.a.y, B.a.z), E2 = mk(B.b.x, B.b.y, B.b.z);
    V3 T = sub(r.o, mk(v1.x, v1.y, v1.z));
    V3 P = cross(r.d, E2);
    V3 Q = cross(T, E1);
    float row1 = dot(Q, E2), row2 = dot(P, T), row3 = dot(Q, r.d);
    float inv = 1.0f / dot(P, E1);
    float dist = inv * row1, uu = inv * row2, vv = inv * row3;
    if (uu >= 0.0f && vv >= 0.0f && dist >= 0.0f && (uu + vv) <= 1.0f && dist < FLT_MAX) {
        if (w.btri < 0 || w.ct < w.K) { w.K = w.ct; w.bd = FLT_MAX; w.bkey = INT_MAX; }
        if (dist < w.bd || (dist == w.bd && pos < w.bkey)) { w.bd = dist; w.bu = uu; w.bv = vv; w.btri = __float_as_int(nt.w); w.bkey = pos; }
    }
}

// ---------------------------------------------------------------------------------------------
// MeshOctree.GetRayIntersection as two alternating steps.
//
// walkDescend pops / expands internal nodes until the lane holds a leaf (at most XN_DESCEND_MAX
// expansions per call); walkLeaf tests at most XN_LEAF_CHUNK triangles of the held leaf.  Callers
// alternate the two, so the lanes of a warp run the same kind of work at the same time instead of
// one lane grinding through a 50-triangle leaf while its neighbours expand nodes.
// Expansion uses the node's 128-byte record: per axis the two child halves are two consecutive
// slabs, so the 8 entry/exit distances come from 12 plane distances (computed once) and min/max.
// The nearest accepted child stays in registers; the others go to the stack farthest first.
// Visiting order, K rule, prune rule and tie-break are those of the oracle's pruned walk, so the
// work counters agree with it exactly.
// ---------------------------------------------------------------------------------------------
// Ray segment of a split query (wavefront.cu, k_trace<.., SPLIT>): only leaves whose entry distance t
// satisfies a <= t < b belong to the segment.  The reference answers a query from the group of leaves
// with the SMALLEST entry distance that holds an accepted triangle (MeshOctree.cs:281-306), and leaves
// with bit-equal t fall into the same segment, so the answer of the first non-empty segment IS the
// reference's answer; the segments of a ray can therefore be walked by different lanes at the same
// time.  Internal nodes are skipped when no descendant leaf can lie in the segment: descendant entry
// distances lie in [entry, exit] of the node's conservative box, which differs from the node's own
// box by at most `overshoot` (DevMesh) in plane position, i.e. by overshoot * max|1/d| in t; the
// slack below is several times that plus a relative term for the rounding of t itself.
struct Seg {
    float a, b;       // leaf filter: a <= t < b
    float aLo, bHi;   // node filter: skip when exit < aLo or entry >= bHi
};

template <bool COUNT, bool MONO = false>
__device__ __forceinline__ bool walkStart(const DevMesh* __restrict__ M, const RayPre& r, Walk& w, Ctr& ctr, float* tExitOut = nullptr) {
    float t0, t1 = 0.0f;
    if (COUNT) ctr.nBox++;
    // rootBox[0..5], rootA, rootB = bytes 24..55 of DevMesh
    const float2 rb0 = __ldg(reinterpret_cast<const float2*>(M->rootBox));
    const float4 rb1 = __ldg(reinterpret_cast<const float4*>(M->rootBox + 2));
    if (MONO) {
        if (!slabMonoGame(make_float3(rb0.x, rb0.y, rb1.x), make_float3(rb1.y, rb1.z, rb1.w), r, t0)) return false;
    } else {
        if (!slab2(make_float3(rb0.x, rb0.y, rb1.x), make_float3(rb1.y, rb1.z, rb1.w), r, t0, t1)) return false;
    }
    if (tExitOut) *tExitOut = t1;
    const int2 rd = __ldg(reinterpret_cast<const int2*>(&M->rootA));
    w.K = INFINITY; w.bd = w.bu = w.bv = 0.0f; w.btri = -1; w.bkey = 0;
    w.ca = rd.x; w.cb = rd.y; w.ct = t0; w.have = true; w.sp = 0;
    return true;
}

// Returns true when the walk is over (nothing held, stack empty).
template <bool COUNT, int STEPS = XN_DESCEND_MAX, bool SEG = false, bool MONO = false>
__device__ __forceinline__ bool walkDescend(const DevScene& S, const RayPre& r, Walk& w, WalkEntry* __restrict__ stack, Ctr& ctr,
                                            const Seg& seg = Seg()) {
    const bool ox = (r.oct & 4) != 0, oy = (r.oct & 2) != 0, oz = (r.oct & 1) != 0;
#pragma unroll 1
    for (int step = 0; step < STEPS; step++) {
        if (!w.have) {
            // pop the next entry that survives the cheap prune test
            while (w.sp > 0) {
                w.sp--;
                const WalkEntry e = stack[w.sp];
                const float et = __int_as_float(e.z);
                if (e.y >= 0 && et > w.K) continue;  // leaf entered after K
                w.ca = e.x; w.cb = e.y; w.ct = et; w.have = true;
                break;
            }
            if (!w.have) return true;
        }
        if (w.cb >= 0) return false;  // a leaf
        // ---- internal node: slab-test all 8 children (MeshOctree.cs:336-339)
        const float4* __restrict__ R = S.inner + 8 * (size_t)w.ca;
        // the 128-byte record in four 32-byte loads: planes x,y | planes z, consMax | descriptors
        const F8 Rxy = ldg256(R), Rzc = ldg256(R + 2);
        const float4 qx = Rxy.a, qy = Rxy.b, qz = Rzc.a;
        // descriptors in front-to-back position order: position p holds child p ^ octant.  The x bit
        // of the octant picks which half of the 64 descriptor bytes is loaded first, the y and z bits
        // are applied as register swaps.
        const F8 Dn = ldg256(R + (ox ? 6 : 4)), Df = ldg256(R + (ox ? 4 : 6));
        int2 ds[8];
        {
            const int2 e0 = make_int2(__float_as_int(Dn.a.x), __float_as_int(Dn.a.y)), e1 = make_int2(__float_as_int(Dn.a.z), __float_as_int(Dn.a.w));
            const int2 e2 = make_int2(__float_as_int(Dn.b.x), __float_as_int(Dn.b.y)), e3 = make_int2(__float_as_int(Dn.b.z), __float_as_int(Dn.b.w));
            const int2 f0 = make_int2(__float_as_int(Df.a.x), __float_as_int(Df.a.y)), f1 = make_int2(__float_as_int(Df.a.z), __float_as_int(Df.a.w));
            const int2 f2 = make_int2(__float_as_int(Df.b.x), __float_as_int(Df.b.y)), f3 = make_int2(__float_as_int(Df.b.z), __float_as_int(Df.b.w));
            const int2 g0 = oy ? e2 : e0, g1 = oy ? e3 : e1, g2 = oy ? e0 : e2, g3 = oy ? e1 : e3;
            const int2 h0 = oy ? f2 : f0, h1 = oy ? f3 : f1, h2 = oy ? f0 : f2, h3 = oy ? f1 : f3;
            ds[0] = oz ? g1 : g0; ds[1] = oz ? g0 : g1; ds[2] = oz ? g3 : g2; ds[3] = oz ? g2 : g3;
            ds[4] = oz ? h1 : h0; ds[5] = oz ? h0 : h1; ds[6] = oz ? h3 : h2; ds[7] = oz ? h2 : h3;
        }
        if (w.ct > w.K) {
            // entered after K: prune on the conservative box (own Min, consMax)
            const float4 qc = Rzc.b;
            float tcns;
            bool h = slabF<MONO>(make_float3(qx.x, qy.x, qz.x), make_float3(qc.x, qc.y, qc.z), r, tcns);
            if (h && tcns > w.K) { w.have = false; continue; }
        }
        if (COUNT) ctr.nBox += 8;
        w.have = false;
        if (!MONO && r.par == 0) {
            // planes in the order the ray meets them: (lo0, hi0, lo1, hi1), reversed for a negative
            // direction; (plane - origin) * inv is monotone in the plane, so near/far need no compare
            const float x0 = ((ox ? qx.w : qx.x) - r.o.x) * r.inv.x, x1 = ((ox ? qx.z : qx.y) - r.o.x) * r.inv.x;
            const float x2 = ((ox ? qx.y : qx.z) - r.o.x) * r.inv.x, x3 = ((ox ? qx.x : qx.w) - r.o.x) * r.inv.x;
            const float y0 = ((oy ? qy.w : qy.x) - r.o.y) * r.inv.y, y1 = ((oy ? qy.z : qy.y) - r.o.y) * r.inv.y;
            const float y2 = ((oy ? qy.y : qy.z) - r.o.y) * r.inv.y, y3 = ((oy ? qy.x : qy.w) - r.o.y) * r.inv.y;
            const float z0 = ((oz ? qz.w : qz.x) - r.o.z) * r.inv.z, z1 = ((oz ? qz.z : qz.y) - r.o.z) * r.inv.z;
            const float z2 = ((oz ? qz.y : qz.z) - r.o.z) * r.inv.z, z3 = ((oz ? qz.x : qz.w) - r.o.z) * r.inv.z;
            // tmin = Max(tz, Max(ty, Max(tx, 0))), tmax = Min(tz', Min(ty', Min(tx', MaxValue))) in the
            // framework's axis order; position p = 4*px + 2*py + pz (first/second half met per axis)
            const float nx[2] = {fmaxf(x0, 0.0f), fmaxf(x2, 0.0f)}, fx[2] = {fminf(x1, FLT_MAX), fminf(x3, FLT_MAX)};
            const float ny[2] = {y0, y2}, fy[2] = {y1, y3}, nz[2] = {z0, z2}, fz[2] = {z1, z3};
#pragma unroll
            for (int p = 7; p >= 0; p--) {
                const int px = p >> 2, py = (p >> 1) & 1, pz = p & 1;
                const float tmin = fmaxf(nz[pz], fmaxf(ny[py], nx[px]));
                const float tmax = fminf(fz[pz], fminf(fy[py], fx[px]));
                const int2 d = ds[p];
                // empty leaves are dropped; a leaf entered after K can never matter; internal nodes are
                // re-examined when popped (conservative box)
                bool take = !(tmin > tmax) && d.y != 0 && (d.y < 0 || !(tmin > w.K));
                if (SEG) take = take && (d.y < 0 ? !(tmax < seg.aLo || tmin >= seg.bHi) : (tmin >= seg.a && tmin < seg.b));
                if (take) {
                    if (w.have) { stack[w.sp] = make_int4(w.ca, w.cb, __float_as_int(w.ct), 0); w.sp++; }
                    w.ca = d.x; w.cb = d.y; w.ct = tmin; w.have = true;
                }
            }
        } else {
            // a direction component below 1e-6 takes the slab test's containment branch: general form
            const float lx[2] = {qx.x, qx.z}, hx[2] = {qx.y, qx.w}, ly[2] = {qy.x, qy.z}, hy[2] = {qy.y, qy.w};
            const float lz[2] = {qz.x, qz.z}, hz[2] = {qz.y, qz.w};
#pragma unroll
            for (int p = 7; p >= 0; p--) {
                const int c = p ^ r.oct;
                const int i = c >> 2, j = (c >> 1) & 1, k = c & 1;
                float tc;
                const bool h = slabF<MONO>(make_float3(i ? lx[1] : lx[0], j ? ly[1] : ly[0], k ? lz[1] : lz[0]),
                                           make_float3(i ? hx[1] : hx[0], j ? hy[1] : hy[0], k ? hz[1] : hz[0]), r, tc);
                const int2 d = ds[p];
                // (segments: the containment branch has no exit distance at hand; internal nodes are kept)
                bool take = h && d.y != 0 && (d.y < 0 || !(tc > w.K));
                if (SEG) take = take && (d.y < 0 ? !(tc >= seg.bHi) : (tc >= seg.a && tc < seg.b));
                if (take) {
                    if (w.have) { stack[w.sp] = make_int4(w.ca, w.cb, __float_as_int(w.ct), 0); w.sp++; }
                    w.ca = d.x; w.cb = d.y; w.ct = tc; w.have = true;
                }
            }
        }
    }
    return false;
}

// Up to XN_LEAF_CHUNK (<= 32) triangles of the held leaf's containingObjects (MeshOctree.cs:286-304),
// in two passes so that the lanes of a warp do not all pay for the longest path of every triangle:
//   pass 1  the reference-identity and back-face tests (RayExtensions.cs:46-51) on the first 16 bytes
//           of each record (surfaceNormal, triangle index) -> bit mask of the triangles to test fully
//   pass 2  the full test for the survivors, in list order.
template <bool COUNT>
__device__ __forceinline__ void walkLeaf(const DevScene& S, const RayPre& r, int ignoreTri, Walk& w, Ctr& ctr) {
    if (!(w.have && w.cb >= 0)) return;
    const float4* __restrict__ T4 = S.leafTri + 4 * (size_t)w.ca;
    unsigned m = 0u;
#if XN_LEAF_SURVIVORS > 0
    // Pass 1 scans until XN_LEAF_SURVIVORS triangles are waiting for the full test (or 32 are scanned, or the leaf
    // ends) instead of a fixed number of triangles: the expensive pass then runs about the same number of times on
    // every lane of the warp, and the trip-count imbalance moves to the cheap pass (ncu, profiles/r01j: the full test
    // was half of all warp instructions at 17 of 32 lanes).
    const int nmax = min(w.cb, 32);
    int n = 0;
#pragma unroll 1
    while (n < nmax && __popc(m) < XN_LEAF_SURVIVORS) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = n + k;
            if (i < nmax) {
                const float4 nt = __ldg(&T4[4 * i]);
                const bool ign = __float_as_int(nt.w) == ignoreTri;          // ignoreTriangle != triangles[i] (reference identity)
                const bool back = dot(mk(nt.x, nt.y, nt.z), r.d) > 0.0f;     // dot(surfaceNormal, D) > 0 -> back face
                if (COUNT) { ctr.nTriRef++; if (!ign) { if (back) ctr.nTriCulled++; else ctr.nTriFull++; } }
                if (!ign && !back) m |= 1u << i;
            }
        }
        n = min(n + 4, nmax);
    }
#else
    const int n = min(w.cb, XN_LEAF_CHUNK);
#pragma unroll 4
    for (int i = 0; i < n; i++) {
        const float4 nt = __ldg(&T4[4 * i]);
        const bool ign = __float_as_int(nt.w) == ignoreTri;          // ignoreTriangle != triangles[i] (reference identity)
        const bool back = dot(mk(nt.x, nt.y, nt.z), r.d) > 0.0f;     // dot(surfaceNormal, D) > 0 -> back face
        if (COUNT) { ctr.nTriRef++; if (!ign) { if (back) ctr.nTriCulled++; else ctr.nTriFull++; } }
        if (!ign && !back) m |= 1u << i;
    }
#endif
#if XN_LEAF_PAIR
    // two survivors per trip: both records are requested before either test starts, and the two tests interleave
    while (m) {
        const int i0 = __ffs((int)m) - 1;
        m &= m - 1u;
        const bool two = m != 0u;
        const int i1 = two ? __ffs((int)m) - 1 : i0;
        if (two) m &= m - 1u;
        const float4* __restrict__ R0 = T4 + 4 * i0;
        const float4* __restrict__ R1 = T4 + 4 * i1;
        const F8 A0 = ldg256(R0), B0 = ldg256(R0 + 2), A1 = ldg256(R1), B1 = ldg256(R1 + 2);
        fullTriangleTest(A0.a, A0.b, B0, w.ca + i0, r, w);
        if (two) fullTriangleTest(A1.a, A1.b, B1, w.ca + i1, r, w);
    }
#else
    while (m) {
        const int i = __ffs((int)m) - 1;
        m &= m - 1u;
        const float4* __restrict__ R = T4 + 4 * i;
        const F8 A = ldg256(R), B = ldg256(R + 2);
        fullTriangleTest(A.a, A.b, B, w.ca + i, r, w);
    }
#endif
    w.ca += n; w.cb -= n;
    if (w.cb == 0) w.have = false;
}

// The whole warp on ONE lane's leaf: lane j tests triangles j, j+32, ... of the leaf held by lane
// `src`, the candidates are reduced by (distance, position) and the owner applies them to its
// walk.  Used when only a few lanes of a warp hold leaves (the tail of a batch, where a handful of
// grazing rays walk hundreds of leaves one after the other): a 50-triangle leaf costs two steps
// instead of fifty.  The minimum over a set does not depend on the order of its elements, so the
// owner ends in exactly the state the sequential loop would have produced.  Must be called by all
// 32 lanes; `r`, `ignoreTri`, `w` are the caller's own (only lane src's are used / updated).
template <bool COUNT>
__device__ __forceinline__ void walkLeafWarp(const DevScene& S, int src, const RayPre& r, int ignoreTri, Walk& w, Ctr& ctr) {
    const unsigned FULL = 0xffffffffu;
    const int lane = (int)(threadIdx.x & 31u);
    RayPre q;
    q.o = mk(__shfl_sync(FULL, r.o.x, src), __shfl_sync(FULL, r.o.y, src), __shfl_sync(FULL, r.o.z, src));
    q.d = mk(__shfl_sync(FULL, r.d.x, src), __shfl_sync(FULL, r.d.y, src), __shfl_sync(FULL, r.d.z, src));
    q.inv = q.o; q.par = 0; q.oct = 0;  // unused by the triangle test
    const int ign = __shfl_sync(FULL, ignoreTri, src);
    const int first = __shfl_sync(FULL, w.ca, src), n = __shfl_sync(FULL, w.cb, src);
    Walk c;  // this lane's candidate: reuse the bookkeeping with an "open group" that accepts anything
    c.K = 0.0f; c.ct = 0.0f; c.bd = FLT_MAX; c.bu = c.bv = 0.0f; c.btri = -1; c.bkey = INT_MAX;
    c.ca = c.cb = 0; c.have = false; c.sp = 0;
    const float4* __restrict__ T4 = S.leafTri + 4 * (size_t)first;
    for (int i = lane; i < n; i += 32) {
        const float4* __restrict__ R = T4 + 4 * i;
        const F8 A = ldg256(R);
        const bool ig = __float_as_int(A.a.w) == ign;
        const bool back = dot(mk(A.a.x, A.a.y, A.a.z), q.d) > 0.0f;
        if (COUNT) { ctr.nTriRef++; if (!ig) { if (back) ctr.nTriCulled++; else ctr.nTriFull++; } }
        if (!ig && !back) {
            const F8 B = ldg256(R + 2);
            // c.btri >= 0 after the first accepted triangle and c.ct == c.K: plain (distance, position) merge
            fullTriangleTest(A.a, A.b, B, first + i, q, c);
        }
    }
    // minimum (distance, position) over the lanes
    float d = c.bd;
    int key = c.bkey, from = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float d2 = __shfl_xor_sync(FULL, d, o);
        const int k2 = __shfl_xor_sync(FULL, key, o), f2 = __shfl_xor_sync(FULL, from, o);
        if (d2 < d || (d2 == d && k2 < key)) { d = d2; key = k2; from = f2; }
    }
    const float bu = __shfl_sync(FULL, c.bu, from), bv = __shfl_sync(FULL, c.bv, from);
    const int bt = __shfl_sync(FULL, c.btri, from);
    if (lane == src) {
        if (bt >= 0) {
            if (w.btri < 0 || w.ct < w.K) { w.K = w.ct; w.bd = FLT_MAX; w.bkey = INT_MAX; }
            if (d < w.bd || (d == w.bd && key < w.bkey)) { w.bd = d; w.bu = bu; w.bv = bv; w.btri = bt; w.bkey = key; }
        }
        w.ca += n; w.cb = 0; w.have = false;
    }
}

// One thread, one ray: the whole walk.  LEAFONLY: every MeshOctree of the scene is a single leaf
// (<= itemTreshold triangles per mesh, e.g. the Free_crate scene) — the expansion code and the stack
// are compiled out, which is worth a third of the registers of the level kernel.
template <bool COUNT, bool LEAFONLY, bool MONO = false>
__device__ __forceinline__ bool traceMesh(const DevScene& S, const DevMesh* __restrict__ M, const RayPre& r, int ignoreTri,
                                          MeshHit& out, Ctr& ctr) {
    Walk w;
    if (!walkStart<COUNT, MONO>(M, r, w, ctr)) return false;
    if (LEAFONLY) {
        while (w.have) walkLeaf<COUNT>(S, r, ignoreTri, w, ctr);
    } else {
        WalkEntry stack[kMeshStack];
        for (;;) {
            if (walkDescend<COUNT, XN_DESCEND_MAX, false, MONO>(S, r, w, stack, ctr)) break;
            walkLeaf<COUNT>(S, r, ignoreTri, w, ctr);
        }
    }
    out.tri = w.btri; out.key = w.bkey; out.u = w.bu; out.v = w.bv; out.d = w.bd;
    return w.btri >= 0;
}

struct Hit {
    int obj, mesh, tri;
    int key;  // position of the hit triangle's leaf reference (its 64-byte record in leafTri)
    float u, v, d;
    V3 pos;  // worldPosition
};

// Per-object block of OctreeSpatialManager.GetRayIntersection (OctreeSpatialManager.cs:347-428).
// The object-space ray depends on the object through InverseWorld only, so objects whose matrices are
// bit-identical (DevObj.xformClass: e.g. every object placed with the identity) share it: `cls`/`r`/`rOk`
// carry the ray of the class seen last from one call to the next.
template <bool COUNT, bool LEAFONLY, bool MONO = false>
__device__ __forceinline__ void objectQuery(const DevScene& S, int objIdx, V3 ro, V3 rd, int ignMesh, int ignTri,
                                            float& minDistance, bool& found, Hit& best, Ctr& ctr, int& cls, RayPre& r, bool& rOk) {
    const DevObj* __restrict__ O = &S.objs[objIdx];
    const int4 om = __ldg(reinterpret_cast<const int4*>(&O->meshStart));  // meshStart, meshCount, xformClass, -
    if (om.z != cls) {
        cls = om.z;
        // v1 = P * W^-1 ; v2 = (P + D) * W^-1 ; D' = Normalize(v2 - v1)   (:358-364)
        V3 rdp = add(ro, rd);
        const M43 Wi = loadM43(O->inv);
        V3 v1 = transform43(ro, Wi);
        V3 v2 = transform43(rdp, Wi);
        V3 dd = normalize(sub(v2, v1));
        rOk = finite3(v1) && finite3(dd);  // a non-finite ray fails every triangle test (NaN comparisons)
        if (rOk) r = prepRay(v1, dd);
    }
    if (!rOk) return;
    const int ms = om.x, mc = om.y;
    for (int mi = 0; mi < mc; mi++) {
        int meshId = __ldg(&S.objMeshIds[ms + mi]);
        const DevMesh* __restrict__ M = &S.meshes[meshId];
        if (COUNT) ctr.nBox++;
        if (!meshBoxSlab<MONO>(M, r)) continue;
        MeshHit mh;
        int ign = (ignMesh == meshId) ? ignTri : -1;
        const bool ok = traceMesh<COUNT, LEAFONLY, MONO>(S, M, r, ign, mh, ctr);
        if (ok && mh.d < minDistance) {
            minDistance = mh.d;
            best.obj = objIdx; best.mesh = meshId; best.tri = mh.tri; best.key = mh.key;
            best.u = mh.u; best.v = mh.v; best.d = mh.d;
            found = true;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// OctreeSpatialManager.GetRayIntersection.  Same group/early-exit semantics one level up; with
// <= itemTreshold (20) objects the tree is the root leaf alone.
// ---------------------------------------------------------------------------------------------
template <bool COUNT, bool LEAFONLY = false, bool MONO = false>
__device__ __forceinline__ bool traceScene(const DevScene& S, V3 ro, V3 rd, int ignMesh, int ignTri, Hit& out, Ctr& ctr) {
    if (!finite3(ro) || !finite3(rd)) return false;
    const bool cnt = COUNT;
    const float4* __restrict__ nodes = S.nodes;
    RayPre r = prepRay(ro, rd);
    float4 lo = __ldg(&nodes[2 * S.sceneRoot]), hi = __ldg(&nodes[2 * S.sceneRoot + 1]);
    float t0;
    if (cnt) ctr.nBox++;
    if (!slabF<MONO>(make_float3(lo.x, lo.y, lo.z), make_float3(hi.x, hi.y, hi.z), r, t0)) return false;
    int w0 = __float_as_int(lo.w), w1 = __float_as_int(hi.w);
    bool found = false;
    Hit best;
    best.obj = best.mesh = best.tri = -1; best.key = 0; best.u = best.v = best.d = 0.0f;
    StackEntry stack[kSceneStack];
    float K = INFINITY;
    float bestD = 0.0f;
    int bestKey = 0;
    int sp = 0;
    float tcur = t0;
    // With <= itemTreshold objects the root is a leaf and this loop runs its body once.
    for (;;) {
        if (w1 >= 0) {
            // leaf: every object in list order; strict `<` keeps the first of equal distances (:371)
            float ld = FLT_MAX;
            bool lf = false;
            Hit lh;
            lh.obj = lh.mesh = lh.tri = -1; lh.key = 0; lh.u = lh.v = lh.d = 0.0f;
            int lkey = 0;
            int cls = -1;
            RayPre ro_;   // object-space ray of the transform class seen last
            ro_.o = ro_.d = ro_.inv = mk(0.0f, 0.0f, 0.0f); ro_.par = ro_.oct = 0;
            bool roOk = false;
            for (int i = 0; i < w1; i++) {
                const float before = ld;
                const bool fb = lf;
                objectQuery<COUNT, LEAFONLY, MONO>(S, __ldg(&S.refs[w0 + i]), ro, rd, ignMesh, ignTri, ld, lf, lh, ctr, cls, ro_, roOk);
                if (lf && (!fb || ld < before)) lkey = w0 + i;
            }
            if (lf) {
                // order inside a group of equal-key leaves: position of the object reference; the same
                // object reached through two leaves reports the same hit, so that is enough
                if (!found || tcur < K) { found = true; K = tcur; best = lh; bestD = ld; bestKey = lkey; }
                else if (tcur == K && (ld < bestD || (ld == bestD && lkey < bestKey))) { best = lh; bestD = ld; bestKey = lkey; }
            }
        } else {
            const int oct = r.oct;
            for (int i = 7; i >= 0; i--) {
                int c = w0 + (i ^ oct);
                float4 clo = __ldg(&nodes[2 * c]), chi = __ldg(&nodes[2 * c + 1]);
                float tc;
                if (cnt) ctr.nBox++;
                if (slabF<MONO>(make_float3(clo.x, clo.y, clo.z), make_float3(chi.x, chi.y, chi.z), r, tc)) {
                    if (!(tc > K) || __float_as_int(chi.w) < 0) { stack[sp].node = c; stack[sp].t = tc; sp++; }
                }
            }
        }
        bool got = false;
        while (sp > 0) {
            sp--;
            int n = stack[sp].node;
            float tn = stack[sp].t;
            float4 nlo = __ldg(&nodes[2 * n]), nhi = __ldg(&nodes[2 * n + 1]);
            int nw1 = __float_as_int(nhi.w);
            if (tn > K) {
                if (nw1 >= 0) continue;
                float4 cm = __ldg(&S.consMax[-1 - nw1]);
                float tcns;
                bool h = slabF<MONO>(make_float3(nlo.x, nlo.y, nlo.z), make_float3(cm.x, cm.y, cm.z), r, tcns);
                if (h && tcns > K) continue;
            }
            w0 = __float_as_int(nlo.w); w1 = nw1; tcur = tn;
            got = true;
            break;
        }
        if (!got) break;
    }
    if (found) {
        // objectSpacePosition = v1 + (v2-v1)*u + (v3-v1)*v (MeshOctree.cs:310-312); world = Po * World (:441-443)
        const float4* __restrict__ T4 = S.leafTri + 4 * (size_t)best.key;
        float4 a = __ldg(&T4[1]), b = __ldg(&T4[2]), c = __ldg(&T4[3]);
        V3 po = add(add(mk(a.x, a.y, a.z), mul(mk(b.x, b.y, b.z), best.u)), mul(mk(c.x, c.y, c.z), best.v));
        best.pos = transform43(po, S.objs[best.obj].world);
        out = best;
    }
    return found;
}

// ---------------------------------------------------------------------------------------------
// Material.LookupUV
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int toIntTrunc(float f) {
    // C# (int)float: out-of-range / NaN -> 0x80000000; the oracle clamps that to texel 0
    if (!(f > -2147483904.0f && f < 2147483648.0f)) return INT_MIN;
    return (int)f;
}
__device__ __forceinline__ void wrapUV(float& u, float& v) {  // Material.cs:125-136 (C# % == fmodf)
    if (u > 1.0f) u = fmodf(u, 1.0f);
    if (v > 1.0f) v = fmodf(v, 1.0f);
    if (u < 0.0f) u = 1.0f + fmodf(u, 1.0f);
    if (v < 0.0f) v = 1.0f + fmodf(v, 1.0f);
}
__device__ __forceinline__ uint32_t texel(const uint32_t* __restrict__ tex, int W, int H, int x, int y) {
    x = max(0, min(x, W - 1));
    y = max(0, min(y, H - 1));
    return __ldg(&tex[(size_t)W * (size_t)y + (size_t)x]);
}
__device__ __forceinline__ V3 lookupUV(const uint32_t* __restrict__ tex, int W, int H, float u, float v, int addressMode,
                                       int filtering) {
    const float BYTE_RECIPROCAL = 1.0f / 255.0f;  // Material.cs:28
    if (addressMode == XNART_ADDRESS_CLAMP) {
        u = (u > 1.0f) ? 1.0f : u; u = (u < 0.0f) ? 0.0f : u;
        v = (v > 1.0f) ? 1.0f : v; v = (v < 0.0f) ? 0.0f : v;
    } else if (addressMode == XNART_ADDRESS_WRAP) {
        wrapUV(u, v);
    } else {  // Mirror, Material.cs:102-123
        float ou = u, ov = v;
        wrapUV(u, v);
        if (toIntTrunc(ou - u) % 2 == 0) u = 1.0f - u;
        if (toIntTrunc(ov - v) % 2 == 0) v = 1.0f - v;
    }
    if (filtering == XNART_FILTER_POINT) {  // GetColorPoint, Material.cs:145-160
        int x = toIntTrunc(u * (float)(W - 1));
        int y = toIntTrunc(v * (float)(H - 1));
        uint32_t a = texel(tex, W, H, x, y);
        return mk((float)((a >> 16) & 255u) * BYTE_RECIPROCAL, (float)((a >> 8) & 255u) * BYTE_RECIPROCAL,
                  (float)(a & 255u) * BYTE_RECIPROCAL);
    }
    // GetColorBilinear, Material.cs:162-232 (IEEERemainder in double)
    float tdx = 1.0f / (float)W, tdy = 1.0f / (float)H;
    double remX = remainder((double)u, (double)tdx);
    double remY = remainder((double)v, (double)tdy);
    u -= (float)remX;
    v -= (float)remY;
    int x = toIntTrunc(u * (float)(W - 1));
    int y = toIntTrunc(v * (float)(H - 1));
    int x2 = toIntTrunc((u + tdx) * (float)(W - 1));
    int y2 = toIntTrunc((v + tdy) * (float)(H - 1));
    uint32_t c00 = texel(tex, W, H, x, y), c10 = texel(tex, W, H, x2, y), c01 = texel(tex, W, H, x, y2), c11 = texel(tex, W, H, x2, y2);
    V3 base = mk((float)((c00 >> 16) & 255u), (float)((c00 >> 8) & 255u), (float)(c00 & 255u));
    V3 bx = mk((float)((c10 >> 16) & 255u), (float)((c10 >> 8) & 255u), (float)(c10 & 255u));
    V3 by = mk((float)((c01 >> 16) & 255u), (float)((c01 >> 8) & 255u), (float)(c01 & 255u));
    V3 bxy = mk((float)((c11 >> 16) & 255u), (float)((c11 >> 8) & 255u), (float)(c11 & 255u));
    float dx = (float)(remX * (double)W) + 0.5f;
    float dy = (float)(remY * (double)H) + 0.5f;
    float idx = 1.0f - dx, idy = 1.0f - dy;
    V3 s = add(add(add(mul(mul(base, idx), idy), mul(mul(by, idx), dy)), mul(mul(bx, dx), idy)), mul(mul(bxy, dx), dy));
    return mul(s, BYTE_RECIPROCAL);
}

// ---------------------------------------------------------------------------------------------
// ILight.GetLightForFragment
// ---------------------------------------------------------------------------------------------
// `toL`: Normalize(Position - position) of a spot light when the caller already has it (the shadow ray's direction,
// shade.cuh: shadowDir — the same operations on the same operands).
__device__ __forceinline__ V3 lightForFragmentTo(const DevLight& L, V3 toL, V3 normal) {
    if (L.kind == XNART_LIGHT_DIRECTIONAL) {  // DirectionalLight.cs:23-30 (uses +Direction)
        float sd = dot(mk(L.dx, L.dy, L.dz), normal);
        if (sd < 0.0f) sd = 0.0f;
        return mul(mul(mk(L.cr, L.cg, L.cb), sd), L.intensity);
    }
    float sd = dot(toL, normal);
    if (sd < 0.0f) return mk(0.0f, 0.0f, 0.0f);
    float ld = dot(neg(toL), mk(L.dx, L.dy, L.dz));
    if (ld > L.angleCosine) {
        float spot = L.intensity * (float)((double)(ld - L.angleCosine) / L.powDen);
        // (float)Math.Pow(surfaceDot, 12): products in double (error < 2^-51 relative, so the
        // binary32 rounding agrees with a correctly rounded pow except on ~1e-8 of inputs)
        double x = (double)sd;
        double x2 = x * x, x4 = x2 * x2, x8 = x4 * x4;
        float hl = (float)(x8 * x4);
        V3 a = mul(mul(mk(L.cr, L.cg, L.cb), spot), sd);
        return add(a, mk(1.0f * hl, 1.0f * hl, 1.0f * hl));
    }
    return mk(0.0f, 0.0f, 0.0f);
}
__device__ __forceinline__ V3 lightForFragment(const DevLight& L, V3 position, V3 normal) {
    V3 toL = mk(0.0f, 0.0f, 0.0f);
    if (L.kind != XNART_LIGHT_DIRECTIONAL) toL = normalize(sub(mk(L.px, L.py, L.pz), position));
    return lightForFragmentTo(L, toL, normal);
}

}  // namespace xn
