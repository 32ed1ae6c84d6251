// path.cu — the trace loop for SMALL opaque scenes (the Free_crate configs of BASELINE.json) as ONE
// persistent kernel per frame (hand-written CUDA, sm_100a).
//
// Eligible scenes (xnart_capi.cu: xnart_build_scene): the OctreeSpatialManager tree is a single leaf
// (<= itemTreshold objects, OctreeSpatialManager.cs:103), every MeshOctree is a single leaf
// (<= itemTreshold triangles, MeshOctree.cs:86), no material is transparent, and the pairs / leaf
// references fit SmallScene.  Everything else keeps the level kernels (kernels.cu) or the wavefront
// path (wavefront.cu).
//
// Why a different schedule.  ncu on the level kernel of round 1 (profiles/r01j_k_level_crate_full.md):
// 80 registers with 650 B of spills (local : global L1 sectors = 6 : 1), 33 % warps active, issue
// active 57 %; a frame was 9 level launches + a resolve, the six deepest levels 3.7 us of pure launch
// latency each, and every ray crossed HBM twice through the ray queue (40 B) and the path arena (48 B).
// Here:
//   * one thread follows one pixel's whole reflection chain (RayTracer.CastRay, RayTracer.cs:506-737,
//     is tail-recursive for opaque materials: one reflection ray per hit, :545-559); a lane whose path
//     has ended takes the next pixel from the frame (persistent warps, batched refill), so lanes at
//     different recursion depths share a warp — the code of a depth is the same;
//   * the scene is a kernel PARAMETER (SmallScene, device_types.h): the pairs and triangles are walked
//     in the same order by all 32 lanes, so their data are uniform constant-bank operands;
//   * the per-level records CastRay needs after its recursive call returns (lightResult, surfaceColor,
//     1 - Reflectiveness, :584) live in a per-thread local array and are folded bottom-up by the same
//     thread when the chain ends, packing to the 8-bit Color at every level exactly where the recursion
//     does (:705, :726, :732): no ray queue, no path arena, no resolve launch.
// Per-ray results do not depend on which lane traces a ray or when, so frames are bit-identical to the
// level kernels and to the oracle, and the instrumented build's counters equal the oracle's.
#include "kernels.h"
#include "shade.cuh"

#ifndef XN_PATH_THREADS
#define XN_PATH_THREADS 128
#endif
#ifndef XN_PATH_MIN_BLOCKS
#define XN_PATH_MIN_BLOCKS 7
#endif
#ifndef XN_PATH_BATCH
#define XN_PATH_BATCH 32        // pixels a warp takes from the frame per atomic (one 8x4 tile)
#endif
#ifndef XN_PATH_UNROLL
#define XN_PATH_UNROLL 2       // triangles of the back-face pass in flight per lane
#endif
#ifndef XN_PATH_EARLY
#define XN_PATH_EARLY 1         // leave the full triangle test after u when u is out of range
#endif
#ifndef XN_PATH_REFILL_MIN
#define XN_PATH_REFILL_MIN 24   // idle lanes that trigger a refill while other lanes are still on a path
#endif

namespace xn {

constexpr int kPathLevels = 32;   // >= kMaxLevels + 1 of xnart_capi.cu
constexpr int kPathUnroll = XN_PATH_UNROLL;

// Vector3.Transform with a matrix packed as 4 rows x 3 columns (SmallPair): m[3*r + c].
__device__ __forceinline__ V3 transform43p(V3 p, const float* m) {
    return mk((((p.x * m[0]) + (p.y * m[3])) + (p.z * m[6])) + m[9],
              (((p.x * m[1]) + (p.y * m[4])) + (p.z * m[7])) + m[10],
              (((p.x * m[2]) + (p.y * m[5])) + (p.z * m[8])) + m[11]);
}
// BoundingBox.Intersects(ref Ray) against a box of the SmallScene.  `anyPar` (warp-uniform): some lane's ray has a
// component below the framework's 1e-6 threshold; otherwise every lane takes the three non-parallel axes of slabAxis
// (trace.cuh) — the same operations in the same order, without the predicated parallel-axis code, which made up half
// of the slab test's instructions (ncu, profiles/r02k: 55 warp instructions per box, 3.7 boxes per query).
__device__ __forceinline__ bool slab6(const float* b, const RayPre& r, bool anyPar, float& t) {
    if (anyPar) { float tExit; return slabGeneral(make_float3(b[0], b[1], b[2]), make_float3(b[3], b[4], b[5]), r, t, tExit); }
    float tmin = 0.0f, tmax = FLT_MAX;
    const float x1 = (b[0] - r.o.x) * r.inv.x, x2 = (b[3] - r.o.x) * r.inv.x;
    tmin = fmaxf(fminf(x1, x2), tmin); tmax = fminf(fmaxf(x1, x2), tmax);
    const float y1 = (b[1] - r.o.y) * r.inv.y, y2 = (b[4] - r.o.y) * r.inv.y;
    tmin = fmaxf(fminf(y1, y2), tmin); tmax = fminf(fmaxf(y1, y2), tmax);
    const float z1 = (b[2] - r.o.z) * r.inv.z, z2 = (b[5] - r.o.z) * r.inv.z;
    tmin = fmaxf(fminf(z1, z2), tmin); tmax = fminf(fmaxf(z1, z2), tmax);
    t = tmin;
    return !(tmin > tmax);
}

// OctreeSpatialManager.GetRayIntersection (OctreeSpatialManager.cs:312-455) for a SmallScene, called by
// all 32 lanes of a converged warp; `act` = this lane has a ray.  The loops over pairs and triangles are
// warp-uniform, per-lane decisions are predicates.
//
// Reference semantics restated (one leaf per tree, so there is exactly one group per SortedList):
//   scene root slab test (:457-461), then every object of the leaf in list order (:337-341), the ray
//   moved to object space (:358-364), every mesh: Mesh.RayIntersects (:368), MeshOctree root slab test
//   (MeshOctree.cs:331), every triangle of the leaf in list order (:286-304) with the reference-identity
//   skip (:290) and IntersectsTriangleBackfaceCulling; `distance < minDistance` keeps the first of equal
//   distances both inside a mesh (MeshOctree.cs:294) and across meshes / objects
//   (OctreeSpatialManager.cs:371), which is ONE running strict minimum over the whole visiting order.
template <bool COUNT>
__device__ __forceinline__ bool traceSmall(const SmallScene& SS, bool act, V3 ro, V3 rd, int ignM, int ignT, Hit& out, Ctr& ctr) {
    const unsigned FULL = 0xffffffffu;
    bool live = act && finite3(ro) && finite3(rd);
    {
        // The scene root's slab test (:457-461).  A finite ray that starts inside the box (bounds included) passes it
        // whatever its direction: on a non-parallel axis (mn - p) <= 0 <= (mx - p) exactly, the reciprocal is finite and
        // non-zero, so one product is <= 0 and the other >= 0, tmin stays 0 and tmax >= 0; a parallel axis tests p in
        // [mn, mx].  Shadow and reflection rays start on a surface, i.e. inside: when no lane of the warp starts outside,
        // the world-space reciprocals and the test are skipped (the verdict is known), else all lanes run it.
        const float* sb = SS.sceneBox;
        const bool inside = ro.x >= sb[0] && ro.x <= sb[3] && ro.y >= sb[1] && ro.y <= sb[4] && ro.z >= sb[2] && ro.z <= sb[5];
        if (COUNT && live) ctr.nBox++;
        if (__any_sync(FULL, live && !inside)) {
            const RayPre rw = prepRay(ro, rd);
            float t0;
            live = slab6(SS.sceneBox, rw, __any_sync(FULL, rw.par != 0), t0) && live;
        }
    }
    float bd = FLT_MAX;
    // `out` is written by the first accepted triangle and only read by the caller after a hit; `r` is written in the first
    // iteration below (no class is -1): neither is initialised here — zeroing them cost 20 instructions per query
    if (!__any_sync(FULL, live)) return false;
    int cls = -1;
    RayPre r;
    bool rOk = false, rPar = false;
#pragma unroll 1
    for (int p = 0; p < SS.nPairs; p++) {
        const SmallPair& P = SS.pair[p];
        if (P.cls != cls) {   // objects with bit-identical InverseWorld share the object-space ray
            cls = P.cls;
            // v1 = P * W^-1 ; v2 = (P + D) * W^-1 ; D' = Normalize(v2 - v1)   (OctreeSpatialManager.cs:358-364)
            const V3 v1 = transform43p(ro, P.inv);
            const V3 v2 = transform43p(add(ro, rd), P.inv);
            const V3 dd = normalize(sub(v2, v1));
            rOk = finite3(v1) && finite3(dd);   // a non-finite ray fails every triangle test (NaN comparisons)
            r = prepRay(v1, dd);
            rPar = __any_sync(FULL, r.par != 0);
        }
        bool go = live && rOk;
        float tb, ct;
        // Mesh.RayIntersects (Mesh.cs:34-39).  When the MeshOctree root box lies inside Mesh.MeshBoundingBox
        // (always, for boxes made by TracerModelProcessor: the mesh box is the vertex box grown to the origin) a ray
        // that passes the root test passes this one too — every operation of the slab test is monotone in the box
        // bounds — so the verdict of the pair is the root test's alone; the instrumented build still runs it.
        if (COUNT || !P.boxCoversRoot) {
            if (COUNT && go) ctr.nBox++;
            go = slab6(P.box, r, rPar, tb) && go;
        }
        if (COUNT && go) ctr.nBox++;
        go = slab6(P.rootBox, r, rPar, ct) && go;     // root of the MeshOctree (MeshOctree.cs:331)
        if (!__any_sync(FULL, go)) continue;
        const int ign = (ignM == P.mesh) ? ignT : -1;
        V3 po;   // set with `improved`
        bool improved = false;
#pragma unroll 1
        for (int c0 = 0; c0 < P.count; c0 += 32) {
            const int n = min(32, P.count - c0);
            const float4* __restrict__ T4 = SS.tri + 4 * (P.first + c0);
            // pass 1: reference identity (:290) and back-face test (RayExtensions.cs:46-51) -> bit mask
            unsigned m = 0u;
            if (!COUNT && P.identity) {
                // the k-th reference of the run is triangle k (a single-leaf MeshOctree built from the mesh's own list):
                // the ignored triangle is bit `ign - c0`, cleared once after the loop; groups of four with constant bit
                // positions (SmallScene.tri is padded, the bits past n are cleared below)
#pragma unroll 1
                for (int i0 = 0; i0 < n; i0 += 4) {
                    unsigned nib = 0u;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const float4 nt = T4[4 * (i0 + k)];
                        if (!(dot(mk(nt.x, nt.y, nt.z), r.d) > 0.0f)) nib |= 1u << k;
                    }
                    m |= nib << i0;
                }
                if (n < 32) m &= (1u << n) - 1u;
                const unsigned ib = (unsigned)(ign - c0);
                if (ib < 32u) m &= ~(1u << ib);
                if (!go) m = 0u;
            } else {
#pragma unroll kPathUnroll
                for (int i = 0; i < n; i++) {
                    const float4 nt = T4[4 * i];
                    const bool ig = __float_as_int(nt.w) == ign;
                    const bool back = dot(mk(nt.x, nt.y, nt.z), r.d) > 0.0f;
                    if (COUNT && go) { ctr.nTriRef++; if (!ig) { if (back) ctr.nTriCulled++; else ctr.nTriFull++; } }
                    if (go && !ig && !back) m |= 1u << i;
                }
            }
            // pass 2: the full test (RayExtensions.cs:52-75), warp-uniform over the union of the lanes' survivors
            unsigned any = __reduce_or_sync(FULL, m);
            while (any) {
                const int i = __ffs((int)any) - 1;
                any &= any - 1u;
                if ((m >> i) & 1u) {
                    const float4 nt = T4[4 * i], a = T4[4 * i + 1], b = T4[4 * i + 2], c = T4[4 * i + 3];
                    const V3 E1 = mk(b.x, b.y, b.z), E2 = mk(c.x, c.y, c.z), V1 = mk(a.x, a.y, a.z);
                    const V3 T = sub(r.o, V1);
                    const V3 Pv = cross(r.d, E2);
                    const float row2 = dot(Pv, T);
                    const float inv = 1.0f / dot(Pv, E1);
                    const float uu = inv * row2;
#if XN_PATH_EARLY
                    // u alone rejects most candidates: 0 <= u, and u <= u + v <= 1 for every accepted v >= 0 (rounding
                    // is monotone), so a lane outside [0, 1] cannot pass the test below whatever v and dist are
                    if (!(uu >= 0.0f && uu <= 1.0f)) continue;
#endif
                    const V3 Q = cross(T, E1);
                    const float row1 = dot(Q, E2), row3 = dot(Q, r.d);
                    const float dist = inv * row1, vv = inv * row3;
                    if (uu >= 0.0f && vv >= 0.0f && dist >= 0.0f && (uu + vv) <= 1.0f && dist < bd) {
                        bd = dist;
                        out.u = uu; out.v = vv; out.tri = __float_as_int(nt.w); out.key = P.refBase + c0 + i;
                        out.obj = P.obj; out.mesh = P.mesh;
                        // objectSpacePosition = v1 + (v2-v1)*u + (v3-v1)*v (MeshOctree.cs:310-312)
                        po = add(add(V1, mul(E1, uu)), mul(E2, vv));
                        improved = true;
                    }
                }
            }
        }
        if (improved) out.pos = transform43p(po, P.world);   // world = Po * World (:441-443)
    }
    out.d = bd;
    return bd < FLT_MAX;
}

// shard-local pixel index -> (x, y), 32-bit arithmetic (shade.cuh: pixelOf divides 64-bit indices)
__device__ __forceinline__ void pixelOf32(const DevFrame& F, unsigned q, int& x, int& y) {
    if (F.samples) { x = (int)q; y = 0; return; }   // sample batch: out[q]
    unsigned lr;
    if (F.tiled) {
        const unsigned rowW = 4u * (unsigned)F.vpW;
        const unsigned tr = q / rowW;
        const unsigned within = q - tr * rowW;
        const unsigned l = within & 31u;
        x = (int)((within >> 5) * 8u + (l & 7u));
        lr = tr * 4u + (l >> 3);
    } else {
        lr = q / (unsigned)F.vpW;
        x = (int)(q - lr * (unsigned)F.vpW);
    }
    if (F.shardCount > 1) {
        const unsigned lb = lr / (unsigned)F.shardBandRows, within = lr - lb * (unsigned)F.shardBandRows;
        y = (int)((lb * (unsigned)F.shardCount + (unsigned)F.shardIndex) * (unsigned)F.shardBandRows + within);
    } else {
        y = (int)lr;
    }
}

// Level-0 ray of shard-local pixel / sample `idx` (RayTracer.Render, RayTracer.cs:410-421).  Returns where the pixel's
// colour goes (renderTargetData[y*W + x], :425; out[q] for a sample batch): the lane keeps that instead of `idx`, so the
// index -> (x, y) division runs once per pixel.  xnart_capi.cu (usePath) keeps frames of 2^31 pixels or more off this kernel.
__device__ __forceinline__ int primaryRay(const DevFrame& F, unsigned idx, V3& o, V3& d) {
    float sx, sy;
    int at;
    if (F.samples) {   // GetColorForQuadrant's sub-pixel positions (RayTracer.cs:220-232)
        const float2 sp = F.samples[(unsigned long long)F.pixelBase + idx];
        sx = sp.x; sy = sp.y;
        at = (int)((unsigned)F.pixelBase + idx);
    } else {
        int x, y;
        pixelOf32(F, (unsigned)F.pixelBase + idx, x, y);
        sx = (float)x; sy = (float)y;
        at = y * F.vpW + x;
    }
    o = unproject(F, sx, sy, F.ndcZ[0]);
    const V3 farP = unproject(F, sx, sy, F.ndcZ[1]);
    d = normalize(sub(farP, o));
    return at;
}

// One lane = one pixel's path; one iteration of the loop = ONE GetRayIntersection query per lane.  A lane is in
// one of two phases, and lanes of both phases share every query:
//   q < 0    the closest-hit query of CastRay (:512) for the ray (ro, rd) ignoring triangle (ignM, ignT);
//            on a hit the lane computes the fragment normal (:520-531) and the reflection direction (:547-550),
//            and moves ro / ign to the hit: they are the origin / ignore identity both of the shadow rays
//            (:482-485) and of the reflection ray (:551-559);
//   q >= 0   IsLightPathObstructed's query towards light q (:465-502); after the last light the surface colour
//            (:568-581) closes the level: the record goes to the lane's local chain and the reflection ray is the
//            next query, or — deepest level (:708-727) — the chain is folded and the pixel written.
// Keeping ONE copy of the query code in the loop matters: the round-1 structure (query, then a second inlined
// query for the shadow ray) ran at a 73 % instruction-cache hit rate with `no_instruction` the top stall
// (profiles/r02b_k_path_crate_full.md), and lanes that had missed idled through the shadow query.
template <bool COUNT>
__global__ void __launch_bounds__(XN_PATH_THREADS, XN_PATH_MIN_BLOCKS)
k_path(const __grid_constant__ SmallScene SS, DevScene S, const __grid_constant__ DevFrame F, const DevLight* __restrict__ lights,
       uint32_t* __restrict__ out, unsigned* __restrict__ fetch, unsigned long long* __restrict__ counters) {
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned total = (unsigned)F.chunkPixels;
    Ctr ctr;
    ctr.nBox = ctr.nTriRef = ctr.nTriCulled = ctr.nTriFull = 0;
    unsigned nHits = 0, nRefl = 0;

    // per-level records of the chain this lane is following (RayTracer.cs:584 needs them after the recursive call)
    float4 recA[kPathLevels];   // (lightResult.xyz, 1 - Reflectiveness)
    float4 recB[kPathLevels];   // (surfaceColor.xyz, -)

    int pix = -1;       // the pixel this lane is working on (its index in `out`), -1 = idle
    int level = 0;      // `iteration` of CastRay
    int q = -1;         // phase: -1 closest hit, >= 0 shadow query of light q
    // idle lanes carry a harmless ray so that the predicated arithmetic never takes a division slow path
    V3 ro = mk(0.0f, 0.0f, 0.0f), rd = mk(0.48f, 0.6f, 0.64f);
    int ignM = -1, ignT = -1;
    V3 N = mk(0.0f, 0.0f, 1.0f), lightResult = mk(0.0f, 0.0f, 0.0f);
    float hu = 0.0f, hv = 0.0f;
    unsigned next = 0u, end = 0u;   // warp-uniform: the batch of pixels this warp is handing out
    bool exhausted = false;         // warp-uniform: the frame has no more batches

    for (;;) {
        // ---- refill: idle lanes take the next pixels of the warp's batch, the warp takes batches from the frame
        unsigned idle = __ballot_sync(FULL, pix < 0);
        if (idle == FULL || (!exhausted && __popc(idle) >= XN_PATH_REFILL_MIN)) {
            bool fresh = false;
            while (idle != 0u && !exhausted) {
                if (next >= end) {
                    unsigned b = 0u;
                    if (lane == 0) {
                        b = atomicAdd(fetch, (unsigned)XN_PATH_BATCH);
                        // RayTracer.Progress (RayTracer.cs:43-46): pixels handed out so far, for xnart_progress
                        if (F.progress && (b & (64u * XN_PATH_BATCH - 1u)) == 0u)
                            *reinterpret_cast<volatile unsigned long long*>(F.progress) = F.progressBase + (unsigned long long)F.pixelBase + (b < total ? b : total);
                    }
                    b = __shfl_sync(FULL, b, 0);
                    if (b >= total) { exhausted = true; break; }
                    next = b;
                    end = min(b + (unsigned)XN_PATH_BATCH, total);
                }
                const unsigned avail = end - next;
                const unsigned rank = (unsigned)__popc(idle & ((1u << lane) - 1u));
                if (pix < 0 && rank < avail) { pix = (int)(next + rank); fresh = true; }
                next += min((unsigned)__popc(idle), avail);
                idle = __ballot_sync(FULL, pix < 0);
            }
            if (fresh) {
                pix = primaryRay(F, (unsigned)pix, ro, rd);   // from here on: the pixel's position in `out`
                level = 0; q = -1; ignM = -1; ignT = -1;
            }
            if (idle == FULL) break;   // nothing left in the frame and every lane has finished its path
        }
        __syncwarp();

        // ---- one GetRayIntersection query for every lane that has a ray
        const bool act = pix >= 0;
        V3 qd = rd;
        float distToLight = 0.0f;
        if (act && q >= 0) qd = shadowDir(lights[q], ro, distToLight);   // IsLightPathObstructed, :467-482
        Hit h;
        const bool hit = traceSmall<COUNT>(SS, act, ro, qd, ignM, ignT, h, ctr);   // :512 / :485
        if (!act) continue;

        bool shaded = false, done = false;
        uint32_t col = 0u;
        if (q < 0) {
            if (!hit) {
                col = packColor(mk(0.0f, 0.0f, 0.0f), F.semantics);   // :732
                done = true;
            } else {
                nHits++;
                const DevMesh* __restrict__ M = &S.meshes[h.mesh];
                N = fragmentNormal(S, M, M->triBase + h.tri, h);   // :520-531
                // the reflection ray (:547-559) starts at the hit and ignores its triangle, like the shadow rays
                rd = normalize(reflect(rd, N));
                ro = h.pos; ignM = h.mesh; ignT = h.tri;
                hu = h.u; hv = h.v;
                lightResult = mk(0.0f, 0.0f, 0.0f);
                q = 0;
                shaded = F.nLights == 0;
            }
        } else {
            // opaque occluders only: amount is 1 or 0 (:487-501); :538-541
            const float amount = (hit && h.d < distToLight) ? 1.0f : 0.0f;
            if (amount != 1.0f) lightResult = add(lightResult, mul(lightForFragmentTo(lights[q], qd, N), 1.0f - amount));   // qd = Normalize(Position - ro)
            q++;
            shaded = q == F.nLights;
        }
        if (shaded) {
            const DevMesh* __restrict__ M = &S.meshes[ignM];
            float alphaW;
            const V3 surf = surfaceColor(S, F, M, M->triBase + ignT, hu, hv, alphaW);   // :568-581
            if (!(level < F.maxReflections)) {                                           // :545
                col = packColor(mulv(lightResult, surf), F.semantics);                   // :726
                done = true;
            } else {
                recA[level] = make_float4(lightResult.x, lightResult.y, lightResult.z, M->oneMinusReflectiveness);
                recB[level] = make_float4(surf.x, surf.y, surf.z, 0.0f);
                level++;       // the next query is the reflection ray (Q9: cast whatever the reflectiveness)
                nRefl++;
                q = -1;
            }
        }
        if (done) {
            // the recursion unwinds: Lerp(reflectionColor.ToVector3(), surfaceColor, 1 - Reflectiveness) * lightResult
            // (:584) and the 8-bit pack (:705) at every level
            for (int k = level - 1; k >= 0; k--) {
                const float4 a = recA[k], b = recB[k];
                const V3 cv = mulv(lerp(unpackColor(col), mk(b.x, b.y, b.z), a.w), mk(a.x, a.y, a.z));
                col = packColor(cv, F.semantics);
            }
            out[(unsigned)pix] = col;   // renderTargetData[y*W + x] (:425)
            pix = -1;
            ro = mk(0.0f, 0.0f, 0.0f); rd = mk(0.48f, 0.6f, 0.64f);
        }
    }

    const unsigned sH = warpSum(nHits), sR = warpSum(nRefl);
    if (lane == 0) {
        if (sH) atomicAdd(&counters[CTR_HITS], (unsigned long long)sH);
        if (sR) atomicAdd(&counters[CTR_RAYS_REFLECTION], (unsigned long long)sR);
    }
    if (COUNT) {
        const unsigned a = warpSum(ctr.nBox), b = warpSum(ctr.nTriRef), c = warpSum(ctr.nTriCulled), e = warpSum(ctr.nTriFull);
        if (lane == 0) {
            atomicAdd(&counters[CTR_NBOX], (unsigned long long)a);
            atomicAdd(&counters[CTR_NTRIREF], (unsigned long long)b);
            atomicAdd(&counters[CTR_NTRICULLED], (unsigned long long)c);
            atomicAdd(&counters[CTR_NTRIFULL], (unsigned long long)e);
        }
    }
}

int pathKernelBlocksPerSm(bool count) {
    int perSm = 0;
    cudaError_t e = count ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_path<true>, XN_PATH_THREADS, 0)
                          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_path<false>, XN_PATH_THREADS, 0);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return perSm;
}

cudaError_t launchPath(const SmallScene& SS, const DevScene& S, const DevFrame& F, const DevLight* lights, uint32_t* out,
                       unsigned* fetch, unsigned long long* counters, int blocks, bool count, cudaStream_t st) {
    // no more warps than batches: a warp without a batch would only run the exit path
    const long long batches = ((long long)F.chunkPixels + XN_PATH_BATCH - 1) / XN_PATH_BATCH;
    const long long need = (batches + (XN_PATH_THREADS / 32) - 1) / (XN_PATH_THREADS / 32);
    if (need < blocks) blocks = (int)(need < 1 ? 1 : need);
    if (count) k_path<true><<<blocks, XN_PATH_THREADS, 0, st>>>(SS, S, F, lights, out, fetch, counters);
    else k_path<false><<<blocks, XN_PATH_THREADS, 0, st>>>(SS, S, F, lights, out, fetch, counters);
    return cudaGetLastError();
}

}  // namespace xn
