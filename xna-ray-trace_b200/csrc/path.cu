// path.cu — the trace loop for SMALL opaque scenes (the Free_crate configs of BASELINE.json) as ONE
// persistent kernel per frame (hand-written CUDA, sm_100a).
//
// Eligible scenes (xnart_capi.cu: buildSmallScene): the OctreeSpatialManager tree is a single leaf
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
#define XN_PATH_MIN_BLOCKS 5
#endif
#ifndef XN_PATH_BATCH
#define XN_PATH_BATCH 64        // pixels a warp takes from the frame per atomic (two 8x4 tiles)
#endif
#ifndef XN_PATH_REFILL_MIN
#define XN_PATH_REFILL_MIN 8    // idle lanes that trigger a refill while other lanes are still on a path
#endif

namespace xn {

constexpr int kPathLevels = 32;   // >= kMaxLevels + 1 of xnart_capi.cu

// Vector3.Transform with a matrix packed as 4 rows x 3 columns (SmallPair): m[3*r + c].
__device__ __forceinline__ V3 transform43p(V3 p, const float* m) {
    return mk((((p.x * m[0]) + (p.y * m[3])) + (p.z * m[6])) + m[9],
              (((p.x * m[1]) + (p.y * m[4])) + (p.z * m[7])) + m[10],
              (((p.x * m[2]) + (p.y * m[5])) + (p.z * m[8])) + m[11]);
}
__device__ __forceinline__ bool slab6(const float* b, const RayPre& r, float& t) {
    return slab(make_float3(b[0], b[1], b[2]), make_float3(b[3], b[4], b[5]), r, t);
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
// SHADOW: only the distance is wanted (IsLightPathObstructed compares it with the light's, RayTracer.cs:485).
template <bool COUNT, bool SHADOW>
__device__ __forceinline__ bool traceSmall(const SmallScene& SS, bool act, V3 ro, V3 rd, int ignM, int ignT, Hit& out, Ctr& ctr) {
    const unsigned FULL = 0xffffffffu;
    bool live = act && finite3(ro) && finite3(rd);
    {
        const RayPre rw = prepRay(ro, rd);
        float t0;
        if (COUNT && live) ctr.nBox++;
        live = slab6(SS.sceneBox, rw, t0) && live;
    }
    float bd = FLT_MAX;
    out.obj = out.mesh = out.tri = -1; out.key = 0; out.u = out.v = 0.0f; out.d = 0.0f;
    out.pos = mk(0.0f, 0.0f, 0.0f);
    if (!__any_sync(FULL, live)) return false;
    int cls = -1;
    RayPre r;
    r.o = r.d = r.inv = mk(0.0f, 0.0f, 0.0f); r.par = r.oct = 0;
    bool rOk = false;
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
        }
        bool go = live && rOk;
        float tb, ct;
        if (COUNT && go) ctr.nBox++;
        go = slab6(P.box, r, tb) && go;         // Mesh.RayIntersects (Mesh.cs:34-39)
        if (COUNT && go) ctr.nBox++;
        go = slab6(P.rootBox, r, ct) && go;     // root of the MeshOctree (MeshOctree.cs:331)
        if (!__any_sync(FULL, go)) continue;
        const int ign = (ignM == P.mesh) ? ignT : -1;
        V3 po = mk(0.0f, 0.0f, 0.0f);
        bool improved = false;
#pragma unroll 1
        for (int c0 = 0; c0 < P.count; c0 += 32) {
            const int n = min(32, P.count - c0);
            const float4* __restrict__ T4 = SS.tri + 4 * (P.first + c0);
            // pass 1: reference identity (:290) and back-face test (RayExtensions.cs:46-51) -> bit mask
            unsigned m = 0u;
#pragma unroll 4
            for (int i = 0; i < n; i++) {
                const float4 nt = T4[4 * i];
                const bool ig = __float_as_int(nt.w) == ign;
                const bool back = dot(mk(nt.x, nt.y, nt.z), r.d) > 0.0f;
                if (COUNT && go) { ctr.nTriRef++; if (!ig) { if (back) ctr.nTriCulled++; else ctr.nTriFull++; } }
                if (go && !ig && !back) m |= 1u << i;
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
                    const V3 Q = cross(T, E1);
                    const float row1 = dot(Q, E2), row2 = dot(Pv, T), row3 = dot(Q, r.d);
                    const float inv = 1.0f / dot(Pv, E1);
                    const float dist = inv * row1, uu = inv * row2, vv = inv * row3;
                    if (uu >= 0.0f && vv >= 0.0f && dist >= 0.0f && (uu + vv) <= 1.0f && dist < bd) {
                        bd = dist;
                        if (!SHADOW) {
                            out.u = uu; out.v = vv; out.tri = __float_as_int(nt.w); out.key = P.refBase + c0 + i;
                            out.obj = P.obj; out.mesh = P.mesh;
                            // objectSpacePosition = v1 + (v2-v1)*u + (v3-v1)*v (MeshOctree.cs:310-312)
                            po = add(add(V1, mul(E1, uu)), mul(E2, vv));
                            improved = true;
                        }
                    }
                }
            }
        }
        if (!SHADOW && improved) out.pos = transform43p(po, P.world);   // world = Po * World (:441-443)
    }
    out.d = bd;
    return bd < FLT_MAX;
}

// Level-0 ray of shard-local pixel / sample `idx` (RayTracer.Render, RayTracer.cs:410-421).
__device__ __forceinline__ void primaryRay(const DevFrame& F, unsigned idx, V3& o, V3& d) {
    float sx, sy;
    if (F.samples) {   // GetColorForQuadrant's sub-pixel positions (RayTracer.cs:220-232)
        const float2 sp = F.samples[F.pixelBase + idx];
        sx = sp.x; sy = sp.y;
    } else {
        int x, y;
        pixelOf(F, F.pixelBase + idx, x, y);
        sx = (float)x; sy = (float)y;
    }
    o = unproject(F, sx, sy, 0.0f);
    const V3 farP = unproject(F, sx, sy, 1.0f);
    d = normalize(sub(farP, o));
}

template <bool COUNT>
__global__ void __launch_bounds__(XN_PATH_THREADS, XN_PATH_MIN_BLOCKS)
k_path(const __grid_constant__ SmallScene SS, DevScene S, const __grid_constant__ DevFrame F, const DevLight* __restrict__ lights,
       uint32_t* __restrict__ out, unsigned* __restrict__ fetch, unsigned long long* __restrict__ counters,
       volatile unsigned* __restrict__ progress) {
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned total = (unsigned)F.chunkPixels;
    Ctr ctr;
    ctr.nBox = ctr.nTriRef = ctr.nTriCulled = ctr.nTriFull = 0;
    unsigned nHits = 0, nRefl = 0;

    // per-level records of the chain this lane is following (RayTracer.cs:584 needs them after the recursive call)
    float4 recA[kPathLevels];   // (lightResult.xyz, 1 - Reflectiveness)
    float4 recB[kPathLevels];   // (surfaceColor.xyz, -)

    int pix = -1;       // shard-local pixel this lane is working on, -1 = idle
    int level = 0;      // `iteration` of CastRay
    // idle lanes carry a harmless ray so that the predicated arithmetic never takes a division slow path
    V3 ro = mk(0.0f, 0.0f, 0.0f), rd = mk(0.48f, 0.6f, 0.64f);
    int ignM = -1, ignT = -1;
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
                        if (progress && (b & (64u * XN_PATH_BATCH - 1u)) == 0u) *progress = b < total ? b : total;
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
                primaryRay(F, (unsigned)pix, ro, rd);
                level = 0; ignM = -1; ignT = -1;
            }
            if (idle == FULL) break;   // nothing left in the frame and every lane has finished its path
        }
        __syncwarp();

        // ---- CastRay, one recursion depth for every lane that has a ray
        const bool act = pix >= 0;
        Hit h0;
        const bool hit = traceSmall<COUNT, false>(SS, act, ro, rd, ignM, ignT, h0, ctr);   // :512
        V3 N = mk(0.0f, 0.0f, 1.0f), lightResult = mk(0.0f, 0.0f, 0.0f);
        const DevMesh* __restrict__ M = &S.meshes[hit ? h0.mesh : 0];
        const int triG = hit ? M->triBase + h0.tri : 0;
        if (hit) N = fragmentNormal(S, M, triG, h0);   // :520-531
        for (int q = 0; q < F.nLights; q++) {          // :535-542
            const DevLight& L = lights[q];
            float distToLight = 0.0f;
            V3 sd = mk(0.48f, 0.6f, 0.64f);
            if (hit) sd = shadowDir(L, h0.pos, distToLight);   // IsLightPathObstructed, :467-482
            Hit sh;
            const bool shHit = traceSmall<COUNT, true>(SS, hit, h0.pos, sd, h0.mesh, h0.tri, sh, ctr);   // :485
            if (hit) {
                // opaque occluders only: amount is 1 or 0 (:487-501); :538-541
                const float amount = (shHit && sh.d < distToLight) ? 1.0f : 0.0f;
                if (amount != 1.0f) lightResult = add(lightResult, mul(lightForFragment(L, h0.pos, N), 1.0f - amount));
            }
        }
        if (act) {
            uint32_t col = 0u;
            bool done = true;
            if (!hit) {
                col = packColor(mk(0.0f, 0.0f, 0.0f), F.semantics);   // :732
            } else {
                nHits++;
                float alphaW;
                const V3 surf = surfaceColor(S, F, M, triG, h0.u, h0.v, alphaW);   // :568-581
                if (!(level < F.maxReflections)) {                                   // :545
                    col = packColor(mulv(lightResult, surf), F.semantics);           // :726
                } else {
                    recA[level] = make_float4(lightResult.x, lightResult.y, lightResult.z, M->oneMinusReflectiveness);
                    recB[level] = make_float4(surf.x, surf.y, surf.z, 0.0f);
                    // the reflection ray (:547-559) — always cast, whatever the reflectiveness (Q9)
                    rd = normalize(reflect(rd, N));
                    ro = h0.pos; ignM = h0.mesh; ignT = h0.tri;
                    level++;
                    nRefl++;
                    done = false;
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
                int x, y;
                pixelOf(F, F.pixelBase + (long long)pix, x, y);
                out[(size_t)y * (size_t)F.vpW + (size_t)x] = col;   // renderTargetData[y*W + x] (:425)
                pix = -1;
                rd = mk(0.48f, 0.6f, 0.64f);
            }
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
                       unsigned* fetch, unsigned long long* counters, unsigned* progress, int blocks, bool count, cudaStream_t st) {
    // no more warps than batches: a warp without a batch would only run the exit path
    const long long batches = ((long long)F.chunkPixels + XN_PATH_BATCH - 1) / XN_PATH_BATCH;
    const long long need = (batches + (XN_PATH_THREADS / 32) - 1) / (XN_PATH_THREADS / 32);
    if (need < blocks) blocks = (int)(need < 1 ? 1 : need);
    if (count) k_path<true><<<blocks, XN_PATH_THREADS, 0, st>>>(SS, S, F, lights, out, fetch, counters, progress);
    else k_path<false><<<blocks, XN_PATH_THREADS, 0, st>>>(SS, S, F, lights, out, fetch, counters, progress);
    return cudaGetLastError();
}

}  // namespace xn
