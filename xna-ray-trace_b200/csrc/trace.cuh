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

// One axis of the slab test.  Returns false when the parallel branch rejects.
__device__ __forceinline__ bool slabAxis(float mn, float mx, float p, float inv, bool par, float& tmin, float& tmax) {
    if (par) {
        if (p < mn || p > mx) return false;
    } else {
        float t1 = (mn - p) * inv;
        float t2 = (mx - p) * inv;
        float tn = (t1 > t2) ? t2 : t1;
        float tf = (t1 > t2) ? t1 : t2;
        tmin = maxx(tn, tmin);
        tmax = minx(tf, tmax);
    }
    return true;
}

// BoundingBox.Intersects(ref Ray, out float?) — XNA 4.0 form.  tmin only grows and tmax only
// shrinks, so the framework's per-axis `if (tmin > tmax) return null` equals one test at the end.
__device__ __forceinline__ bool slab(float3 mn, float3 mx, const RayPre& r, float& t) {
    float tmin = 0.0f, tmax = FLT_MAX;
    bool ok = slabAxis(mn.x, mx.x, r.o.x, r.inv.x, r.par & 1, tmin, tmax);
    ok = slabAxis(mn.y, mx.y, r.o.y, r.inv.y, r.par & 2, tmin, tmax) && ok;
    ok = slabAxis(mn.z, mx.z, r.o.z, r.inv.z, r.par & 4, tmin, tmax) && ok;
    t = tmin;
    return ok && !(tmin > tmax);
}

struct MeshHit {
    int tri;       // mesh-relative triangle index, -1 = none
    int key;       // position of the winning leaf reference in refs[] / leafTri[]
    float u, v, d;
};

struct StackEntry { int node; float t; };

// ---------------------------------------------------------------------------------------------
// MeshOctree.GetRayIntersection
// ---------------------------------------------------------------------------------------------
template <bool COUNT>
__device__ __forceinline__ bool traceMesh(const DevScene& S, int rootNode, int triBase, const RayPre& r, int ignoreTri,
                                          MeshHit& out, Ctr& ctr) {
    StackEntry stack[kMeshStack];
    const float4* __restrict__ nodes = S.nodes;
    float K = INFINITY;
    bool found = false;
    float bd = 0.0f, bu = 0.0f, bv = 0.0f;
    int btri = -1, bkey = 0;

    float4 lo = __ldg(&nodes[2 * rootNode]), hi = __ldg(&nodes[2 * rootNode + 1]);
    float t0;
    if (COUNT) ctr.nBox++;
    if (!slab(make_float3(lo.x, lo.y, lo.z), make_float3(hi.x, hi.y, hi.z), r, t0)) return false;
    int sp = 0;
    int w0 = __float_as_int(lo.w), w1 = __float_as_int(hi.w);
    float tcur = t0;
    for (;;) {
        if (w1 >= 0) {
            // ---- leaf: every triangle of containingObjects (MeshOctree.cs:286-304)
            bool lf = false;
            float ld = FLT_MAX, lu = 0.0f, lv = 0.0f;
            int lt = -1, lk = 0;
            const float4* __restrict__ R = S.leafTri + 4 * (size_t)w0;
            for (int i = 0; i < w1; i++) {
                // all four loads of the record are issued together (one round trip); a back-facing
                // triangle wastes three of them, which costs L1 bandwidth the kernel has to spare
                const float4* __restrict__ T4 = R + 4 * (size_t)i;
                float4 n4 = __ldg(&T4[0]), a = __ldg(&T4[1]), b = __ldg(&T4[2]), c = __ldg(&T4[3]);
                const int ti = __float_as_int(n4.w);
                if (COUNT) ctr.nTriRef++;
                if (ti == ignoreTri) continue;  // ignoreTriangle != triangles[i] (reference identity)
                // dot(surfaceNormal, D) > 0 -> back face (RayExtensions.cs:48-51)
                if (dot(mk(n4.x, n4.y, n4.z), r.d) > 0.0f) { if (COUNT) ctr.nTriCulled++; continue; }
                if (COUNT) ctr.nTriFull++;
                V3 E1 = mk(b.x, b.y, b.z), E2 = mk(c.x, c.y, c.z);
                V3 T = sub(r.o, mk(a.x, a.y, a.z));
                V3 P = cross(r.d, E2);
                V3 Q = cross(T, E1);
                float row1 = dot(Q, E2), row2 = dot(P, T), row3 = dot(Q, r.d);
                float inv = 1.0f / dot(P, E1);
                float dist = inv * row1, uu = inv * row2, vv = inv * row3;
                if (uu >= 0.0f && vv >= 0.0f && dist >= 0.0f && (uu + vv) <= 1.0f && dist < ld) {
                    ld = dist; lu = uu; lv = vv; lt = ti; lk = w0 + i; lf = true;
                }
            }
            if (lf) {
                if (!found || tcur < K) { found = true; K = tcur; bd = ld; bu = lu; bv = lv; btri = lt; bkey = lk; }
                else if (tcur == K && (ld < bd || (ld == bd && lk < bkey))) { bd = ld; bu = lu; bv = lv; btri = lt; bkey = lk; }
            }
        } else {
            // ---- internal: slab-test all 8 children (MeshOctree.cs:336-339), push the accepted ones
            // farthest first so the nearest pops first
            const int oct = r.oct;
#pragma unroll
            for (int i = 7; i >= 0; i--) {
                int c = w0 + (i ^ oct);
                float4 clo = __ldg(&nodes[2 * c]), chi = __ldg(&nodes[2 * c + 1]);
                float tc;
                if (COUNT) ctr.nBox++;
                if (slab(make_float3(clo.x, clo.y, clo.z), make_float3(chi.x, chi.y, chi.z), r, tc)) {
                    // a leaf entered after K can never matter; internal nodes are re-examined at pop
                    // time against their conservative box
                    if (!(tc > K) || __float_as_int(chi.w) < 0) {
                        stack[sp].node = c; stack[sp].t = tc; sp++;
                    }
                }
            }
        }
        // ---- pop the next node that survives the prune test
        bool got = false;
        while (sp > 0) {
            sp--;
            int n = stack[sp].node;
            float tn = stack[sp].t;
            int nw0 = __float_as_int(__ldg(&nodes[2 * n]).w);
            float4 nhi = __ldg(&nodes[2 * n + 1]);
            int nw1 = __float_as_int(nhi.w);
            if (tn > K) {
                if (nw1 >= 0) continue;  // leaf entered after K
                // internal: entry distance on the conservative box (own Min, consMax)
                float4 nlo = __ldg(&nodes[2 * n]);
                float4 cm = __ldg(&S.consMax[-1 - nw1]);
                float tcns;
                bool h = slab(make_float3(nlo.x, nlo.y, nlo.z), make_float3(cm.x, cm.y, cm.z), r, tcns);
                if (h && tcns > K) continue;
            }
            w0 = nw0; w1 = nw1; tcur = tn;
            got = true;
            break;
        }
        if (!got) break;
    }
    out.tri = btri; out.key = bkey; out.u = bu; out.v = bv; out.d = bd;
    return found;
}

// ---------------------------------------------------------------------------------------------
// MeshOctree.GetRayIntersection, cooperative form: the 8 lanes of an octet walk ONE ray together.
//
// The per-thread walk above diverges badly on deep trees (on the 1M-triangle grid ncu shows ~9 of
// 32 lanes active: one lane is in a 50-triangle leaf while its neighbours test child boxes).  Here
// every step is uniform across the octet: an internal step slab-tests the node's 8 children at
// once (lane j takes the child at front-to-back position j, i.e. child j ^ octant; the 8 x 32 B
// child block is one contiguous 256-byte read), `ballot` yields the mask of accepted children, and
// the walk keeps ONE entry per tree level (child block, pending mask, each lane's own entry
// distance) instead of a stack of nodes; a leaf step tests 8 triangles at a time.  Same visiting
// order, same K rule and same tie-break as the per-thread walk, so results and work counters are
// identical; only the lane-level schedule differs.
// ---------------------------------------------------------------------------------------------
typedef cg::thread_block_tile<8> Octet;
constexpr int kMeshLevels = kMaxMeshDepth + 2;

template <bool COUNT>
__device__ __forceinline__ bool traceMeshOctet(const Octet& g, const DevScene& S, int rootNode, int triBase, const RayPre& r,
                                               int ignoreTri, MeshHit& out, Ctr& ctr) {
    const int lane = (int)g.thread_rank();
    const float4* __restrict__ nodes = S.nodes;
    float lvT[kMeshLevels];
    int lvBase[kMeshLevels];
    unsigned lvMask[kMeshLevels];
    float K = INFINITY;
    bool found = false;      // uniform
    bool rv = false;         // this lane holds a candidate of the current K group
    float bd = 0.0f, bu = 0.0f, bv = 0.0f;
    int btri = -1, bkey = 0;

    float4 lo = __ldg(&nodes[2 * rootNode]), hi = __ldg(&nodes[2 * rootNode + 1]);
    float t0;
    if (COUNT && lane == 0) ctr.nBox++;
    if (!slab(make_float3(lo.x, lo.y, lo.z), make_float3(hi.x, hi.y, hi.z), r, t0)) return false;
    int level = -1;
    int w0 = __float_as_int(lo.w), w1 = __float_as_int(hi.w);
    float tcur = t0;
    for (;;) {
        if (w1 >= 0) {
            // ---- leaf: 8 triangles per step; each lane keeps the best of the ones it tested
            bool lf = false;
            float ld = FLT_MAX, lu = 0.0f, lv = 0.0f;
            int lt = -1, lk = 0;
            const float4* __restrict__ R = S.leafTri + 4 * (size_t)w0;
            for (int i = lane; i < w1; i += 8) {
                const float4* __restrict__ T4 = R + 4 * (size_t)i;
                float4 n4 = __ldg(&T4[0]), a = __ldg(&T4[1]), b = __ldg(&T4[2]), c = __ldg(&T4[3]);
                const int ti = __float_as_int(n4.w);
                if (COUNT) ctr.nTriRef++;
                if (ti == ignoreTri) continue;
                if (dot(mk(n4.x, n4.y, n4.z), r.d) > 0.0f) { if (COUNT) ctr.nTriCulled++; continue; }
                if (COUNT) ctr.nTriFull++;
                V3 E1 = mk(b.x, b.y, b.z), E2 = mk(c.x, c.y, c.z);
                V3 T = sub(r.o, mk(a.x, a.y, a.z));
                V3 P = cross(r.d, E2);
                V3 Q = cross(T, E1);
                float row1 = dot(Q, E2), row2 = dot(P, T), row3 = dot(Q, r.d);
                float inv = 1.0f / dot(P, E1);
                float dist = inv * row1, uu = inv * row2, vv = inv * row3;
                // within a lane positions increase, so strict `<` keeps the earliest of equal distances
                if (uu >= 0.0f && vv >= 0.0f && dist >= 0.0f && (uu + vv) <= 1.0f && dist < ld) {
                    ld = dist; lu = uu; lv = vv; lt = ti; lk = w0 + i; lf = true;
                }
            }
            if (g.any(lf)) {
                if (!found || tcur < K) {  // this leaf opens a nearer group: older candidates are void
                    found = true; K = tcur;
                    rv = lf; bd = ld; bu = lu; bv = lv; btri = lt; bkey = lk;
                } else if (tcur == K && lf && (!rv || ld < bd || (ld == bd && lk < bkey))) {
                    rv = true; bd = ld; bu = lu; bv = lv; btri = lt; bkey = lk;
                }
            }
        } else {
            // ---- internal: lane j tests the child at front-to-back position j
            const int c = w0 + (lane ^ r.oct);
            float4 clo = __ldg(&nodes[2 * c]), chi = __ldg(&nodes[2 * c + 1]);
            float tc;
            if (COUNT) ctr.nBox++;
            bool keep = slab(make_float3(clo.x, clo.y, clo.z), make_float3(chi.x, chi.y, chi.z), r, tc);
            keep = keep && (!(tc > K) || __float_as_int(chi.w) < 0);
            unsigned mask = g.ballot(keep);
            if (mask) { level++; lvT[level] = tc; lvBase[level] = w0; lvMask[level] = mask; }
        }
        // ---- next node: nearest pending child of the deepest level that still has one
        bool got = false;
        while (level >= 0) {
            unsigned mask = lvMask[level];
            if (!mask) { level--; continue; }
            const int j = __ffs((int)mask) - 1;
            lvMask[level] = mask & (mask - 1u);
            const float tn = g.shfl(lvT[level], j);
            const int n = lvBase[level] + (j ^ r.oct);
            float4 nlo = __ldg(&nodes[2 * n]), nhi = __ldg(&nodes[2 * n + 1]);
            const int nw1 = __float_as_int(nhi.w);
            if (tn > K) {
                if (nw1 >= 0) continue;
                float4 cm = __ldg(&S.consMax[-1 - nw1]);
                float tcns;
                bool h = slab(make_float3(nlo.x, nlo.y, nlo.z), make_float3(cm.x, cm.y, cm.z), r, tcns);
                if (h && tcns > K) continue;
            }
            w0 = __float_as_int(nlo.w); w1 = nw1; tcur = tn;
            got = true;
            break;
        }
        if (!got) break;
    }
    if (!found) return false;
    // minimum (distance, position) over the lanes' candidates; every lane ends with the same winner
    float d = rv ? bd : INFINITY;
    int key = rv ? bkey : INT_MAX, src = lane;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        float d2 = g.shfl_xor(d, o);
        int k2 = g.shfl_xor(key, o), s2 = g.shfl_xor(src, o);
        if (d2 < d || (d2 == d && k2 < key)) { d = d2; key = k2; src = s2; }
    }
    out.d = d;
    out.key = key;
    out.u = g.shfl(bu, src);
    out.v = g.shfl(bv, src);
    out.tri = g.shfl(btri, src);
    return true;
}

struct Hit {
    int obj, mesh, tri;
    int key;  // position of the hit triangle's leaf reference (its 64-byte record in leafTri)
    float u, v, d;
    V3 pos;  // worldPosition
};

// Per-object block of OctreeSpatialManager.GetRayIntersection (OctreeSpatialManager.cs:347-428).
// COOP: called by all 8 lanes of an octet with identical arguments; everything but the mesh walk
// is computed redundantly (and counted once).
template <bool COUNT, bool COOP>
__device__ __forceinline__ void objectQuery(const Octet& g, const DevScene& S, int objIdx, V3 ro, V3 rd, int ignMesh, int ignTri,
                                            float& minDistance, bool& found, Hit& best, Ctr& ctr) {
    const DevObj* __restrict__ O = &S.objs[objIdx];
    // v1 = P * W^-1 ; v2 = (P + D) * W^-1 ; D' = Normalize(v2 - v1)   (:358-364)
    V3 rdp = add(ro, rd);
    V3 v1 = transform43(ro, O->inv);
    V3 v2 = transform43(rdp, O->inv);
    V3 dd = normalize(sub(v2, v1));
    if (!finite3(v1) || !finite3(dd)) return;  // a non-finite ray fails every triangle test (NaN comparisons)
    RayPre r = prepRay(v1, dd);
    const int ms = O->meshStart, mc = O->meshCount;
    for (int mi = 0; mi < mc; mi++) {
        int meshId = __ldg(&S.objMeshIds[ms + mi]);
        const DevMesh* __restrict__ M = &S.meshes[meshId];
        float tb;
        if (COUNT && (!COOP || g.thread_rank() == 0)) ctr.nBox++;
        // Mesh.RayIntersects (Mesh.cs:34-39)
        if (!slab(make_float3(M->box[0], M->box[1], M->box[2]), make_float3(M->box[3], M->box[4], M->box[5]), r, tb)) continue;
        MeshHit mh;
        int ign = (ignMesh == meshId) ? ignTri : -1;
        const bool ok = COOP ? traceMeshOctet<COUNT>(g, S, M->rootNode, M->triBase, r, ign, mh, ctr)
                             : traceMesh<COUNT>(S, M->rootNode, M->triBase, r, ign, mh, ctr);
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
template <bool COUNT, bool COOP>
__device__ __forceinline__ bool traceSceneT(const Octet& g, const DevScene& S, V3 ro, V3 rd, int ignMesh, int ignTri, Hit& out, Ctr& ctr) {
    if (!finite3(ro) || !finite3(rd)) return false;
    const bool cnt = COUNT && (!COOP || g.thread_rank() == 0);  // uniform work is counted once per octet
    const float4* __restrict__ nodes = S.nodes;
    RayPre r = prepRay(ro, rd);
    float4 lo = __ldg(&nodes[2 * S.sceneRoot]), hi = __ldg(&nodes[2 * S.sceneRoot + 1]);
    float t0;
    if (cnt) ctr.nBox++;
    if (!slab(make_float3(lo.x, lo.y, lo.z), make_float3(hi.x, hi.y, hi.z), r, t0)) return false;
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
            for (int i = 0; i < w1; i++) {
                const float before = ld;
                const bool fb = lf;
                objectQuery<COUNT, COOP>(g, S, __ldg(&S.refs[w0 + i]), ro, rd, ignMesh, ignTri, ld, lf, lh, ctr);
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
                if (slab(make_float3(clo.x, clo.y, clo.z), make_float3(chi.x, chi.y, chi.z), r, tc)) {
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
                bool h = slab(make_float3(nlo.x, nlo.y, nlo.z), make_float3(cm.x, cm.y, cm.z), r, tcns);
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

// One thread, one ray.
template <bool COUNT>
__device__ __forceinline__ bool traceScene(const DevScene& S, V3 ro, V3 rd, int ignMesh, int ignTri, Hit& out, Ctr& ctr) {
    const Octet g = cg::tiled_partition<8>(cg::this_thread_block());
    return traceSceneT<COUNT, false>(g, S, ro, rd, ignMesh, ignTri, out, ctr);
}

// Every lane of the octet brings (at most) one ray; the octet walks them one after the other, all
// 8 lanes working on the same ray.  Must be called by all 8 lanes together; `want` says whether
// this lane has a ray.  Returns this lane's own result.
template <bool COUNT>
__device__ __forceinline__ bool traceSceneCoop(const Octet& g, const DevScene& S, bool want, V3 ro, V3 rd, int ignMesh,
                                               int ignTri, Hit& out, Ctr& ctr) {
    unsigned pending = g.ballot(want);
    bool myOk = false;
    while (pending) {
        const int src = __ffs((int)pending) - 1;
        pending &= pending - 1u;
        V3 o = mk(g.shfl(ro.x, src), g.shfl(ro.y, src), g.shfl(ro.z, src));
        V3 d = mk(g.shfl(rd.x, src), g.shfl(rd.y, src), g.shfl(rd.z, src));
        const int im = g.shfl(ignMesh, src), it = g.shfl(ignTri, src);
        Hit h;
        const bool ok = traceSceneT<COUNT, true>(g, S, o, d, im, it, h, ctr);
        if ((int)g.thread_rank() == src) { myOk = ok; if (ok) out = h; }
    }
    return myOk;
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
__device__ __forceinline__ V3 lightForFragment(const DevLight& L, V3 position, V3 normal) {
    if (L.kind == XNART_LIGHT_DIRECTIONAL) {  // DirectionalLight.cs:23-30 (uses +Direction)
        float sd = dot(mk(L.dx, L.dy, L.dz), normal);
        if (sd < 0.0f) sd = 0.0f;
        return mul(mul(mk(L.cr, L.cg, L.cb), sd), L.intensity);
    }
    V3 toL = normalize(sub(mk(L.px, L.py, L.pz), position));
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

}  // namespace xn
