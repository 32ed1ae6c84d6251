// wavefront.cu — the trace loop for scenes with deep MeshOctrees (hand-written CUDA, sm_100a).
//
// The cost of one GetRayIntersection query on a deep octree is heavy-tailed: most rays touch a
// few leaves, some graze the surface through hundreds.  With one thread per ray and a whole
// CastRay per thread (kernels.cu) the stragglers hold their warps hostage — ncu on the
// 1M-triangle grid shows 5 of 32 lanes active (profiles/r01c).  Here the recursion level is split
// into its queries and every query batch runs through ONE persistent kernel in which a lane that
// finishes its ray immediately takes the next one from the batch:
//
//   level L:  k_gen_primary (L = 0)      RayTracer.Render's Unproject pair (RayTracer.cs:410-421) -> ray queue
//             k_trace      (level rays)  OctreeSpatialManager / MeshOctree.GetRayIntersection (RayTracer.cs:512)
//             k_shade_a                  hit position, fragment normal (:520-531), shadow rays (:467-482) and
//                                        the reflection / refraction rays of level L+1 (:545-702)
//       side: k_trace      (shadow rays) IsLightPathObstructed's query (:485), nLights segments
//       side: k_shade_b                  light sum (:535-542), surface colour, path node
//   then k_resolve (kernels.cu) bottom-up, as in the per-thread path.
// The rays of level L+1 depend on the hits of level L only, not on its shadow verdicts, so the shadow
// batch and shade B of a level run on a side stream while the main stream already traces the next
// level: the serial tail of one batch (a few grazing rays) overlaps the bulk of the other.
//
// k_trace: persistent warps; per lane a small state machine (idle -> pick the next (object, mesh)
// pair -> walk its octree, alternating descend and leaf steps (trace.cuh) -> merge -> next pair ->
// write the hit).  All 32 lanes stay in the outer loop until the batch is drained, so the phases
// reconverge at __syncwarp() and idle lanes refill from the batch through one warp-aggregated
// atomicAdd.  Per-ray results do not depend on which lane or when a ray is traced, so frames stay
// bit-identical to the per-thread path and to the oracle.
//
// Available when the scene octree is a single leaf (<= itemTreshold objects: every BASELINE
// configuration); scenes with a deeper OctreeSpatialManager tree use the per-thread kernels.
#include "kernels.h"
#include "shade.cuh"

#ifndef XN_TRACE_MIN_BLOCKS
#define XN_TRACE_MIN_BLOCKS 4
#endif
#ifndef XN_TRACE_DESCEND
#define XN_TRACE_DESCEND 2  // node expansions per DESCEND round of the persistent kernel
#endif
#ifndef XN_COOP_LEAF_MAX
#define XN_COOP_LEAF_MAX 12 // with at most this many lanes holding leaves, the warp tests each leaf cooperatively
#endif
#ifndef XN_LEAF_PCT
#define XN_LEAF_PCT 50      // a round runs the LEAF step when at least this share of the walking lanes hold a leaf, else DESCEND
#endif
#ifndef XN_REFILL_MIN
#define XN_REFILL_MIN 24    // idle lanes that trigger a refill while other lanes are still walking
#endif

namespace xn {

// ---------------------------------------------------------------------------------------------
// level 0: primary rays into the queue (tile order, see pixelOf)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gen_primary(DevFrame F, DevArena A) {
    const unsigned count = (unsigned)F.chunkPixels;
    if (blockIdx.x == 0 && threadIdx.x == 0) { A.levelCount[0] = count; A.levelBase[0] = 0u; }
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x) {
        const RayIn r = loadRay(F, A, 0, idx);
        A.rayO[0][idx] = make_float4(r.o.x, r.o.y, r.o.z, r.refIndex);
        A.rayD[0][idx] = make_float4(r.d.x, r.d.y, r.d.z, 0.0f);
        A.rayIgn[0][idx] = make_int2(-1, -1);
    }
}

// ---------------------------------------------------------------------------------------------
// persistent trace kernel
// ---------------------------------------------------------------------------------------------
// SPLIT: every ray of a SMALL batch is traced as kSplit segment queries (trace.cuh: Seg), each by its own
// lane.  Deep recursion levels hold few rays, and a few of them graze the surface through hundreds of
// leaves: one lane walking such a ray alone kept the GPU waiting for most of the launch (ncu launch
// list, profiles/r01h: 42 M warp instructions in 0.68 ms).  Cutting the ray into segments by entry
// distance divides that serial walk by up to kSplit and gives the same answer (the first non-empty
// segment's).  The instrumented build never splits: its counters are those of the canonical walk.
// Is this batch traced as split queries?  Decided on the device (the batch size is only known there); the shade kernels
// that read the results evaluate the same rule (isSplitBatch below).
__device__ __forceinline__ bool splitRule(int nPairs, unsigned splitMax, unsigned nRays, const unsigned* hintHits, const unsigned* hintRays,
                                          unsigned pct) {
    if (splitMax == 0u || nRays > splitMax || nPairs != 1) return false;
    if (hintHits == nullptr || hintRays == nullptr) return true;
    return (unsigned long long)(*hintHits) * 100ull <= (unsigned long long)(*hintRays) * (unsigned long long)pct;
}
__device__ __forceinline__ bool isSplitJob(const DevScene& S, const TraceJob& J, unsigned nRays) {
    return splitRule(S.nPairs, J.splitMax, nRays, J.hintHits, J.hintRays, J.splitHitPct);
}

template <bool COUNT, bool SPLIT>
__global__ void __launch_bounds__(256, XN_TRACE_MIN_BLOCKS) k_trace(DevScene S, TraceJob J, unsigned long long* __restrict__ counters) {
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned segCount = *J.segCount;
    const unsigned nRays = segCount * J.nSeg;
    const bool splitBatch = !COUNT && isSplitJob(S, J, nRays);
    if (splitBatch != SPLIT) return;
    const unsigned total = SPLIT ? nRays * (unsigned)kSplit : nRays;
    Seg seg;
    seg.a = seg.aLo = -INFINITY; seg.b = seg.bHi = INFINITY;
    int mySeg = 0;
    size_t outIdx = 0;
    WalkEntry stack[kMeshStack];
    Ctr ctr;
    ctr.nBox = ctr.nTriRef = ctr.nTriCulled = ctr.nTriFull = 0;

    // Per-lane state kept small on purpose (64 registers, 4 blocks per SM): the world-space ray is
    // re-read from the batch when a lane moves on to another (object, mesh) pair, and the best hit so
    // far lives in the output record (written when a pair improves it), only its distance in a register.
    int slot = -1;            // ray slot this lane owns, -1 = idle
    int pi = 0;               // (object, mesh) pair being tried / walked
    bool walking = false;
    bool exhausted = false;   // warp-uniform: the batch has no more rays to hand out
    RayPre r;                 // object-space ray of the pair being walked
    r.o = r.d = r.inv = mk(0.0f, 0.0f, 0.0f); r.par = r.oct = 0;
    int ignoreTri = -1;
    Walk w;
    w.K = INFINITY; w.bd = w.bu = w.bv = 0.0f; w.btri = -1; w.bkey = 0;
    w.ca = w.cb = 0; w.ct = 0.0f; w.have = false; w.sp = 0;
    float bestD = FLT_MAX;    // minDistance over the pairs tried so far (OctreeSpatialManager.cs:371-377)

    const float4 slo = __ldg(&S.nodes[2 * S.sceneRoot]), shi = __ldg(&S.nodes[2 * S.sceneRoot + 1]);

    for (;;) {
        // ---- FETCH: idle lanes take the next rays of the batch
        // (ray set-up runs with only the refilled lanes active, so refills are batched: at least
        // XN_REFILL_MIN idle lanes, or nothing else to do)
        const unsigned idleMask = __ballot_sync(FULL, slot < 0);
        if (!exhausted && (__popc(idleMask) >= XN_REFILL_MIN || __all_sync(FULL, !walking))) {
            const unsigned n = (unsigned)__popc(idleMask);
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(J.fetch, n);
            base = __shfl_sync(FULL, base, 0);
            if (slot < 0) {
                const unsigned i = base + (unsigned)__popc(idleMask & ((1u << lane) - 1u));
                if (i < total) {
                    unsigned iray = i;
                    if (SPLIT) { iray = i / (unsigned)kSplit; mySeg = (int)(i - iray * (unsigned)kSplit); }
                    unsigned q = 0, h = iray;
                    if (J.nSeg > 1) { q = iray / segCount; h = iray - q * segCount; }
                    slot = (int)(q * J.segStride + h);
                    outIdx = SPLIT ? (size_t)i : (size_t)slot;
                    pi = 0; walking = false;
                    bestD = FLT_MAX;
                    // a non-finite ray fails every triangle test; the scene root must be hit
                    // (OctreeSpatialManager.cs:457-461)
                    const float4 o4 = J.rayO[slot], d4 = J.rayD[slot];
                    const V3 wo = mk(o4.x, o4.y, o4.z), wd = mk(d4.x, d4.y, d4.z);
                    bool ok = finite3(wo) && finite3(wd);
                    if (ok) {
                        const RayPre rw = prepRay(wo, wd);
                        float t0;
                        if (COUNT) ctr.nBox++;
                        ok = slab(make_float3(slo.x, slo.y, slo.z), make_float3(shi.x, shi.y, shi.z), rw, t0);
                    }
                    if (!ok) pi = S.nPairs;
                }
            }
            exhausted = base + n >= total;
        }
        if (__all_sync(FULL, slot < 0)) break;

        // ---- SETUP: lanes between walks pick their next (object, mesh) pair, or finish the ray
        if (slot >= 0 && !walking) {
            if (pi < S.nPairs) {
                const float4 o4 = J.rayO[slot], d4 = J.rayD[slot];
                const int2 ig = J.rayIgn[slot];
                const V3 wo = mk(o4.x, o4.y, o4.z), wd = mk(d4.x, d4.y, d4.z);
                int cls = -1;
                bool rOk = false;
                do {
                    const int2 pr = __ldg(&S.pairs[pi]);
                    const DevObj* __restrict__ O = &S.objs[pr.x];
                    const int oc = __ldg(&O->xformClass);
                    if (oc != cls) {   // objects with bit-identical InverseWorld share the object-space ray
                        cls = oc;
                        // v1 = P * W^-1 ; v2 = (P + D) * W^-1 ; D' = Normalize(v2 - v1)   (OctreeSpatialManager.cs:358-364)
                        const M43 Wi = loadM43(O->inv);
                        const V3 v1 = transform43(wo, Wi);
                        const V3 v2 = transform43(add(wo, wd), Wi);
                        const V3 dd = normalize(sub(v2, v1));
                        rOk = finite3(v1) && finite3(dd);
                        if (rOk) r = prepRay(v1, dd);
                    }
                    if (!rOk) { pi++; continue; }
                    const DevMesh* __restrict__ M = &S.meshes[pr.y];
                    if (COUNT) ctr.nBox++;
                    if (!meshBoxSlab(M, r)) { pi++; continue; }
                    ignoreTri = (ig.x == pr.y) ? ig.y : -1;
                    float tExit = 0.0f;
                    if (!walkStart<COUNT>(M, r, w, ctr, &tExit)) { pi++; continue; }
                    if (SPLIT) {
                        // segment mySeg of [root entry, root exit]; the first reaches back to -inf, the last to +inf
                        const float len = tExit - w.ct;
                        const float lo = w.ct + len * ((float)mySeg * (1.0f / kSplit));
                        const float hi = w.ct + len * ((float)(mySeg + 1) * (1.0f / kSplit));
                        seg.a = (mySeg == 0) ? -INFINITY : lo;
                        seg.b = (mySeg == kSplit - 1) ? INFINITY : hi;
                        const float maxInv = fmaxf((r.par & 1) ? 0.0f : fabsf(r.inv.x), fmaxf((r.par & 2) ? 0.0f : fabsf(r.inv.y), (r.par & 4) ? 0.0f : fabsf(r.inv.z)));
                        const float slack = 8.0f * __ldg(&M->overshoot) * maxInv + 1e-5f * (fabsf(w.ct) + fabsf(tExit)) + 1e-30f;
                        seg.aLo = seg.a - slack; seg.bHi = seg.b + slack;
                        // a root that is itself a leaf belongs to the segment that contains its entry distance
                        if (w.cb >= 0 && !(w.ct >= seg.a && w.ct < seg.b)) { pi++; continue; }
                    }
                    walking = true;
                } while (!walking && pi < S.nPairs);
            }
            if (!walking) {
                // every pair tried: the ray is finished (a hit was written when it became the best)
                if (bestD == FLT_MAX) {
                    int4* __restrict__ o = SPLIT ? J.outSplit : J.out;
                    o[2 * outIdx] = make_int4(-1, -1, -1, 0);
                    o[2 * outIdx + 1] = make_int4(0, 0, 0, 0);
                }
                slot = -1;
            }
        }
        __syncwarp();

        // ---- one phase per round, chosen by majority: LEAF (lanes holding a leaf test a chunk of its
        // triangles) or DESCEND (the others pop / expand one node).  Lanes of the minority wait, so each
        // phase runs with most of the walking lanes active instead of a few.
        const bool wantLeaf = walking && w.have && w.cb >= 0;
        const unsigned nLeaf = (unsigned)__popc(__ballot_sync(FULL, wantLeaf));
        const unsigned nDesc = (unsigned)__popc(__ballot_sync(FULL, walking && !wantLeaf));
        bool over = false;
        if (nLeaf * 100u >= (nLeaf + nDesc) * (unsigned)XN_LEAF_PCT) {
            if (nLeaf <= XN_COOP_LEAF_MAX) {
                // few lanes hold leaves: the whole warp serves them one after the other
                unsigned lm = __ballot_sync(FULL, wantLeaf);
                while (lm) {
                    const int src = __ffs((int)lm) - 1;
                    lm &= lm - 1u;
                    walkLeafWarp<COUNT>(S, src, r, ignoreTri, w, ctr);
                }
            } else if (wantLeaf) {
                walkLeaf<COUNT>(S, r, ignoreTri, w, ctr);
            }
        } else {
            if (walking && !wantLeaf) over = walkDescend<COUNT, XN_TRACE_DESCEND, SPLIT>(S, r, w, stack, ctr, seg);
        }
        __syncwarp();
        if (walking && over) {
            // minDistance bookkeeping of the object loop (OctreeSpatialManager.cs:370-377): strict <,
            // so the first of equal distances stays
            if (w.btri >= 0 && w.bd < bestD) {
                bestD = w.bd;
                const int2 pr = __ldg(&S.pairs[pi]);
                int4* __restrict__ o = SPLIT ? J.outSplit : J.out;
                o[2 * outIdx] = make_int4(pr.x, pr.y, w.btri, w.bkey);
                o[2 * outIdx + 1] = make_int4(__float_as_int(w.bu), __float_as_int(w.bv), __float_as_int(w.bd), 1);
            }
            pi++;
            walking = false;
        }
    }

    if (COUNT) {
        unsigned a = warpSum(ctr.nBox), b = warpSum(ctr.nTriRef), c = warpSum(ctr.nTriCulled), d = warpSum(ctr.nTriFull);
        if (lane == 0) {
            atomicAdd(&counters[CTR_NBOX], (unsigned long long)a);
            atomicAdd(&counters[CTR_NTRIREF], (unsigned long long)b);
            atomicAdd(&counters[CTR_NTRICULLED], (unsigned long long)c);
            atomicAdd(&counters[CTR_NTRIFULL], (unsigned long long)d);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// shade A: closest hits of the level -> hit position, fragment normal, shadow rays
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ Hit hitOf(const DevScene& S, const int4 a, const int4 b) {
    Hit h;
    h.obj = a.x; h.mesh = a.y; h.tri = a.z; h.key = a.w;
    h.u = __int_as_float(b.x); h.v = __int_as_float(b.y); h.d = __int_as_float(b.z);
    h.pos = mk(0.0f, 0.0f, 0.0f);
    return h;
}

// Result of ray `iray` of a batch of `nRays`: the record itself, or — when the batch was traced as split
// queries — the record of its first segment that found anything.
__device__ __forceinline__ void fetchHit(const int4* __restrict__ plain, const int4* __restrict__ split, bool splitBatch, size_t plainIdx,
                                         size_t iray, int4& a, int4& b) {
    if (!splitBatch) { a = plain[2 * plainIdx]; b = plain[2 * plainIdx + 1]; return; }
    const int4* __restrict__ p = split + 2 * iray * kSplit;
    a = make_int4(-1, -1, -1, 0); b = make_int4(0, 0, 0, 0);
#pragma unroll 1
    for (int sgm = 0; sgm < kSplit; sgm++) {
        const int4 x = p[2 * sgm];
        if (x.x >= 0) { a = x; b = p[2 * sgm + 1]; return; }
    }
}
// level >= 0: the closest-hit batch of that level (hint = hits / rays of level - 1, none at level 0);
// shadowOf >= 0: the shadow batch of that level (hint = hits / rays of the level's own closest-hit batch).
__device__ __forceinline__ bool isSplitBatch(const DevScene& S, const DevArena& A, unsigned nRays, int level, int shadowOf) {
    const int h = shadowOf >= 0 ? shadowOf : level - 1;
    if (h < 0) return splitRule(S.nPairs, A.splitMax, nRays, nullptr, nullptr, A.splitHitPct);
    return splitRule(S.nPairs, A.splitMax, nRays, &A.shadowHits[h], &A.levelCount[h], A.splitHitPct);
}

__global__ void __launch_bounds__(256) k_shade_a(DevScene S, DevFrame F, DevArena A, const DevLight* __restrict__ lights, int level) {
    const unsigned count = A.levelCount[level];
    const unsigned base = A.levelBase[level];
    const unsigned nextBase = base + count;
    if (blockIdx.x == 0 && threadIdx.x == 0) { A.levelBase[level + 1] = nextBase; reportProgress(F, level); }
    const bool last = !(level < F.maxReflections);  // RayTracer.cs:545
    const int par = level & 1;
    const bool splitL = isSplitBatch(S, A, count, level, -1);
    unsigned nHits = 0, nRefl = 0, nRefr = 0;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x) {
        const unsigned node = base + idx;
        int4 ha, hb;
        fetchHit(A.hit[par], A.hitSplit[par], splitL, idx, idx, ha, hb);
        if (ha.x < 0) {
            A.color[node] = packColor(mk(0.0f, 0.0f, 0.0f), F.semantics);  // RayTracer.cs:732
            A.flags[node] = 0u;
            continue;
        }
        nHits++;
        Hit h = hitOf(S, ha, hb);
        // objectSpacePosition = v1 + (v2-v1)*u + (v3-v1)*v (MeshOctree.cs:310-312); world = Po * World (:441-443)
        const float4* __restrict__ T4 = S.leafTri + 4 * (size_t)h.key;
        const float4 a = __ldg(&T4[1]), b = __ldg(&T4[2]), c = __ldg(&T4[3]);
        const V3 po = add(add(mk(a.x, a.y, a.z), mul(mk(b.x, b.y, b.z), h.u)), mul(mk(c.x, c.y, c.z), h.v));
        h.pos = transform43(po, S.objs[h.obj].world);
        const DevMesh* __restrict__ M = &S.meshes[h.mesh];
        const V3 N = fragmentNormal(S, M, M->triBase + h.tri, h);
        // compact hit index: the shadow rays of hit k sit at slot q*rayCapacity + k, one segment per light
        unsigned k;
        {
            cg::coalesced_group g = cg::coalesced_threads();
            unsigned b0 = 0;
            if (g.thread_rank() == 0) b0 = atomicAdd(&A.shadowHits[level], g.size());
            k = g.shfl(b0, 0) + g.thread_rank();
        }
        for (int q = 0; q < F.nLights; q++) {
            float distToLight;
            const V3 sd = shadowDir(lights[q], h.pos, distToLight);
            const size_t s = (size_t)q * A.rayCapacity + k;
            A.shO[par][s] = make_float4(h.pos.x, h.pos.y, h.pos.z, distToLight);
            A.shD[par][s] = make_float4(sd.x, sd.y, sd.z, 0.0f);
            A.shIgn[par][s] = make_int2(h.mesh, h.tri);
        }
        // the rays of the next level depend on the hit alone (:547-559, :656-698): emitted here, so the next
        // level can be traced while this level's shadow batch is still running
        if (!last) {
            const float4 o4 = A.rayO[par][idx], d4 = A.rayD[par][idx];
            A.nodeC[node] = emitChildren(S, A, level, nextBase, h, M, N, mk(d4.x, d4.y, d4.z), o4.w, nRefl, nRefr);
        }
        A.nodeA[node] = make_float4(h.pos.x, h.pos.y, h.pos.z, __uint_as_float(k));
        A.nodeB[node] = make_float4(N.x, N.y, N.z, 0.0f);
    }
    const unsigned lane = threadIdx.x & 31u;
    const unsigned sH = warpSum(nHits), sR = warpSum(nRefl), sF = warpSum(nRefr);
    if (lane == 0) {
        if (sH) atomicAdd(&A.counters[CTR_HITS], (unsigned long long)sH);
        if (sR) atomicAdd(&A.counters[CTR_RAYS_REFLECTION], (unsigned long long)sR);
        if (sF) atomicAdd(&A.counters[CTR_RAYS_REFRACTION], (unsigned long long)sF);
    }
}

// ---------------------------------------------------------------------------------------------
// shade B: shadow verdicts -> light sum (:535-542), surface colour, path node (:584, :705-727)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_shade_b(DevScene S, DevFrame F, DevArena A, const DevLight* __restrict__ lights, int level) {
    const unsigned count = A.levelCount[level];
    const unsigned base = A.levelBase[level];
    const bool last = !(level < F.maxReflections);  // RayTracer.cs:545
    const int par = level & 1;
    const bool splitL = isSplitBatch(S, A, count, level, -1);
    const unsigned nShadowHits = A.shadowHits[level];
    const bool splitS = isSplitBatch(S, A, nShadowHits * (unsigned)F.nLights, -1, level);
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x) {
        const unsigned node = base + idx;
        int4 ha, hb;
        fetchHit(A.hit[par], A.hitSplit[par], splitL, idx, idx, ha, hb);
        if (ha.x < 0) continue;
        Hit h = hitOf(S, ha, hb);
        const float4 pa = A.nodeA[node], nb = A.nodeB[node];
        h.pos = mk(pa.x, pa.y, pa.z);
        const unsigned k = __float_as_uint(pa.w);
        const V3 N = mk(nb.x, nb.y, nb.z);
        V3 lightResult = mk(0.0f, 0.0f, 0.0f);
        for (int q = 0; q < F.nLights; q++) {
            const size_t s = (size_t)q * A.rayCapacity + k;
            int4 sa, sb;
            fetchHit(A.shHit[par], A.shSplit[par], splitS, s, (size_t)q * nShadowHits + k, sa, sb);
            const Hit sh = hitOf(S, sa, sb);
            lightResult = addLight(S, lights[q], sa.x >= 0, sh, A.shO[par][s].w, h.pos, N, lightResult);
        }
        const DevMesh* __restrict__ M = &S.meshes[h.mesh];
        float alphaW;
        const V3 surf = surfaceColor(S, F, M, M->triBase + h.tri, h.u, h.v, alphaW);
        if (last) {
            A.color[node] = packColor(mulv(lightResult, surf), F.semantics);  // :726
            A.flags[node] = 0u;
        } else {
            A.nodeA[node] = make_float4(lightResult.x, lightResult.y, lightResult.z, M->oneMinusReflectiveness);
            A.nodeB[node] = make_float4(surf.x, surf.y, surf.z, alphaW);
            A.flags[node] = 1u;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// batch API (xnart_trace_rays): user rays -> queue layout, hits -> user arrays
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_rays(const float* __restrict__ rays6, const int* __restrict__ ignore2, long long n,
                                                   float4* __restrict__ rayO, float4* __restrict__ rayD, int2* __restrict__ rayIgn) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        rayO[i] = make_float4(rays6[6 * i], rays6[6 * i + 1], rays6[6 * i + 2], 1.0f);
        rayD[i] = make_float4(rays6[6 * i + 3], rays6[6 * i + 4], rays6[6 * i + 5], 0.0f);
        rayIgn[i] = ignore2 ? make_int2(ignore2[2 * i], ignore2[2 * i + 1]) : make_int2(-1, -1);
    }
}

__global__ void __launch_bounds__(256) k_unpack_hits(DevScene S, const int4* __restrict__ hit, long long n, int* __restrict__ objId,
                                                     int* __restrict__ meshId, int* __restrict__ triId, float* __restrict__ u,
                                                     float* __restrict__ v, float* __restrict__ d, float* __restrict__ pos3,
                                                     unsigned long long* __restrict__ counters) {
    unsigned nHits = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int4 ha = hit[2 * i], hb = hit[2 * i + 1];
        const bool ok = ha.x >= 0;
        if (ok) nHits++;
        objId[i] = ok ? ha.x : -1; meshId[i] = ok ? ha.y : -1; triId[i] = ok ? ha.z : -1;
        u[i] = ok ? __int_as_float(hb.x) : 0.0f; v[i] = ok ? __int_as_float(hb.y) : 0.0f; d[i] = ok ? __int_as_float(hb.z) : 0.0f;
        if (pos3) {
            V3 p = mk(0.0f, 0.0f, 0.0f);
            if (ok) {
                const float4* __restrict__ T4 = S.leafTri + 4 * (size_t)ha.w;
                const float4 a = __ldg(&T4[1]), b = __ldg(&T4[2]), c = __ldg(&T4[3]);
                const V3 po = add(add(mk(a.x, a.y, a.z), mul(mk(b.x, b.y, b.z), __int_as_float(hb.x))), mul(mk(c.x, c.y, c.z), __int_as_float(hb.y)));
                p = transform43(po, S.objs[ha.x].world);
            }
            pos3[3 * i] = p.x; pos3[3 * i + 1] = p.y; pos3[3 * i + 2] = p.z;
        }
    }
    const unsigned sH = warpSum(nHits);
    if ((threadIdx.x & 31u) == 0 && sH) atomicAdd(&counters[CTR_HITS], (unsigned long long)sH);
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

int traceKernelBlocksPerSm(bool count) {
    int perSm = 0;
    cudaError_t e = count ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_trace<true, false>, 256, 0)
                          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_trace<false, false>, 256, 0);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return perSm;
}

cudaError_t launchGenPrimary(const DevFrame& F, const DevArena& A, int maxBlocks, cudaStream_t st) {
    k_gen_primary<<<gridFor(F.chunkPixels, maxBlocks), 256, 0, st>>>(F, A);
    return cudaGetLastError();
}

cudaError_t launchTrace(const DevScene& S, const TraceJob& J, unsigned long long* counters, int blocks, bool count, cudaStream_t st) {
    if (count) {
        k_trace<true, false><<<blocks, 256, 0, st>>>(S, J, counters);
    } else {
        k_trace<false, false><<<blocks, 256, 0, st>>>(S, J, counters);
        // the split build handles the batch instead when it is small (decided on the device, see k_trace)
        if (J.splitMax != 0u && S.nPairs == 1) k_trace<false, true><<<blocks, 256, 0, st>>>(S, J, counters);
    }
    return cudaGetLastError();
}

cudaError_t launchShadeA(const DevScene& S, const DevFrame& F, const DevArena& A, const DevLight* lights, int level,
                         long long upperBound, int maxBlocks, cudaStream_t st) {
    k_shade_a<<<gridFor(upperBound, maxBlocks), 256, 0, st>>>(S, F, A, lights, level);
    return cudaGetLastError();
}

cudaError_t launchShadeB(const DevScene& S, const DevFrame& F, const DevArena& A, const DevLight* lights, int level,
                         long long upperBound, int maxBlocks, cudaStream_t st) {
    k_shade_b<<<gridFor(upperBound, maxBlocks), 256, 0, st>>>(S, F, A, lights, level);
    return cudaGetLastError();
}

cudaError_t launchPackRays(const float* rays6, const int* ignore2, long long n, float4* rayO, float4* rayD, int2* rayIgn,
                           int maxBlocks, cudaStream_t st) {
    k_pack_rays<<<gridFor(n, maxBlocks), 256, 0, st>>>(rays6, ignore2, n, rayO, rayD, rayIgn);
    return cudaGetLastError();
}

cudaError_t launchUnpackHits(const DevScene& S, const int4* hit, long long n, int* objId, int* meshId, int* triId, float* u, float* v,
                             float* d, float* pos3, unsigned long long* counters, int maxBlocks, cudaStream_t st) {
    k_unpack_hits<<<gridFor(n, maxBlocks), 256, 0, st>>>(S, hit, n, objId, meshId, triId, u, v, d, pos3, counters);
    return cudaGetLastError();
}

}  // namespace xn
