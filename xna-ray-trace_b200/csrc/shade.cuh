// shade.cuh — ray generation, shading and path-arena helpers shared by the level kernels
// (kernels.cu: one thread per ray, whole CastRay per thread) and the wavefront kernels
// (wavefront.cu: persistent trace kernel + streaming shade kernels).
//
//   RayTracer.Render        RayTracer.cs:391-429   pixelOf / unproject / loadRay
//   RayTracer.CastRay       RayTracer.cs:506-737   fragmentNormal / shadowDir / addLight / finishHit
#pragma once
#include "trace.cuh"

namespace xn {

__device__ __forceinline__ unsigned warpSum(unsigned v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// shard-local pixel index -> (x, y) of the full frame
// Tiled frames deal pixels to warps as 8x4 tiles (32 consecutive indices = one tile), so the rays
// of a warp start out close together and walk the same octree nodes; otherwise scanline order.
__device__ __forceinline__ void pixelOf(const DevFrame& F, long long q, int& x, int& y) {
    if (F.samples) { x = (int)q; y = 0; return; }   // sample batch: out[q]
    int lr;
    if (F.tiled) {
        const int rowW = 4 * F.vpW;
        const int tr = (int)(q / rowW);
        const int within = (int)(q - (long long)tr * rowW);
        const int l = within & 31;
        x = (within >> 5) * 8 + (l & 7);
        lr = tr * 4 + (l >> 3);
    } else {
        lr = (int)(q / F.vpW);
        x = (int)(q - (long long)lr * F.vpW);
    }
    if (F.shardCount > 1) {
        int lb = lr / F.shardBandRows, within = lr - lb * F.shardBandRows;
        y = (lb * F.shardCount + F.shardIndex) * F.shardBandRows + within;
    } else {
        y = lr;
    }
}

// Viewport.Unproject with the frame-constant inverse.  `ndcZ` is the source depth after
// (z - MinDepth) / (MaxDepth - MinDepth): the two depths Render uses (0 and 1, RayTracer.cs:414-415) are divided once
// per frame on the host (DevFrame.ndcZ) — with MinDepth = 0 the near one is 0 / 1, and a zero dividend sends the
// device's IEEE division through its slow-path subroutine for every pixel.
__device__ __forceinline__ V3 unproject(const DevFrame& F, float sx, float sy, float ndcZ) {
    V3 s;
    s.x = (((sx - (float)F.vpX) / (float)F.vpW) * 2.0f) - 1.0f;
    s.y = -((((sy - (float)F.vpY) / (float)F.vpH) * 2.0f) - 1.0f);
    s.z = ndcZ;
    const float* m = F.inv;
    V3 v = mk((((s.x * m[0]) + (s.y * m[4])) + (s.z * m[8])) + m[12],
              (((s.x * m[1]) + (s.y * m[5])) + (s.z * m[9])) + m[13],
              (((s.x * m[2]) + (s.y * m[6])) + (s.z * m[10])) + m[14]);
    float a = (((s.x * m[3]) + (s.y * m[7])) + (s.z * m[11])) + m[15];
    float num = a - 1.0f;
    bool within = (-1.401298E-45f <= num) && (num <= 1.401298E-45f);
    if (!within) v = divs(v, a);
    return v;
}

struct RayIn {
    V3 o, d;
    int ignM, ignT;
    float refIndex;
};

// Level 0: RayTracer.Render's per-pixel Unproject pair (:412-421).  Level > 0: the queued ray.
__device__ __forceinline__ RayIn loadRay(const DevFrame& F, const DevArena& A, int level, unsigned idx) {
    RayIn r;
    if (level == 0) {
        float sx, sy;
        if (F.samples) {   // GetColorForQuadrant's sub-pixel positions (RayTracer.cs:220-232)
            const float2 sp = F.samples[F.pixelBase + idx];
            sx = sp.x; sy = sp.y;
        } else {
            int x, y;
            pixelOf(F, F.pixelBase + idx, x, y);
            sx = (float)x; sy = (float)y;
        }
        r.o = unproject(F, sx, sy, F.ndcZ[0]);
        V3 farP = unproject(F, sx, sy, F.ndcZ[1]);
        r.d = normalize(sub(farP, r.o));
        r.ignM = -1; r.ignT = -1; r.refIndex = 1.0f;
    } else {
        const int in = level & 1;
        float4 o4 = A.rayO[in][idx], d4 = A.rayD[in][idx];
        int2 ig = A.rayIgn[in][idx];
        r.o = mk(o4.x, o4.y, o4.z); r.refIndex = o4.w;
        r.d = mk(d4.x, d4.y, d4.z);
        r.ignM = ig.x; r.ignT = ig.y;
    }
    return r;
}

// fragment normal (:520-531)
__device__ __forceinline__ V3 fragmentNormal(const DevScene& S, const DevMesh* __restrict__ M, int triG, const Hit& h) {
    if (M->flags & MESH_INTERP_NORMALS) {
        const float4* __restrict__ T = S.triS + 5 * (size_t)triG;
        float4 s1 = __ldg(&T[1]), s2 = __ldg(&T[2]), s3 = __ldg(&T[3]);
        V3 n1 = mk(s1.z, s1.w, s2.x), n2 = mk(s2.y, s2.z, s2.w), n3 = mk(s3.x, s3.y, s3.z);
        return normalize(add(add(n1, mul(sub(n2, n1), h.u)), mul(sub(n3, n1), h.v)));
    }
    float4 n4 = __ldg(&S.leafTri[4 * (size_t)h.key]);
    return mk(n4.x, n4.y, n4.z);
}

// shadow ray towards a light (:467-482): no epsilon offset, ignoring only the hit triangle (Q4)
__device__ __forceinline__ V3 shadowDir(const DevLight& L, V3 pos, float& distToLight) {
    if (L.kind == XNART_LIGHT_SPOT) {
        V3 dir = sub(mk(L.px, L.py, L.pz), pos);
        distToLight = length(dir);
        return normalize(dir);
    }
    distToLight = FLT_MAX;
    return neg(mk(L.dx, L.dy, L.dz));
}

// IsLightPathObstructed's verdict (:485-501) folded into the light sum (:537-541)
__device__ __forceinline__ V3 addLight(const DevScene& S, const DevLight& L, bool shadowHit, const Hit& sh, float distToLight,
                                       V3 pos, V3 N, V3 lightResult) {
    float amount = 0.0f;
    if (shadowHit && sh.d < distToLight) {
        const DevMesh* __restrict__ SM = &S.meshes[sh.mesh];
        amount = (SM->flags & MESH_TRANSPARENT) ? __ldg(&S.triS[5 * (size_t)(SM->triBase + sh.tri) + 4]).z : 1.0f;
    }
    if (amount != 1.0f) lightResult = add(lightResult, mul(lightForFragment(L, pos, N), 1.0f - amount));
    return lightResult;
}

__device__ __forceinline__ unsigned appendRay(const DevArena& A, int level, int outq, unsigned nextBase, V3 o, V3 d, float refIndex,
                                              int mesh, int tri) {
    cg::coalesced_group g = cg::coalesced_threads();
    unsigned b = 0;
    if (g.thread_rank() == 0) b = atomicAdd(&A.levelCount[level + 1], g.size());
    const unsigned slot = g.shfl(b, 0) + g.thread_rank();
    const unsigned child = nextBase + slot;
    if (slot < A.rayCapacity && child < A.nodeCapacity) {
        A.rayO[outq][slot] = make_float4(o.x, o.y, o.z, refIndex);
        A.rayD[outq][slot] = make_float4(d.x, d.y, d.z, 0.0f);
        A.rayIgn[outq][slot] = make_int2(mesh, tri);
    } else {
        *A.overflow = 1u;
    }
    return child;
}

// Surface colour of a hit (:568-581): texture lookup at the interpolated uv, or the triangle colour.
// alphaW receives triangle.color.W.
__device__ __forceinline__ V3 surfaceColor(const DevScene& S, const DevFrame& F, const DevMesh* __restrict__ M, int triG, float u, float v,
                                           float& alphaW) {
    const float4* __restrict__ T = S.triS + 5 * (size_t)triG;
    float4 s3 = __ldg(&T[3]), s4 = __ldg(&T[4]);
    alphaW = s4.z;
    if (M->flags & MESH_USE_TEXTURE) {  // :568-575
        float4 s0 = __ldg(&T[0]), s1 = __ldg(&T[1]);
        float ax = s0.z - s0.x, ay = s0.w - s0.y;  // uv2 - uv1
        float bx = s1.x - s0.x, by = s1.y - s0.y;  // uv3 - uv1
        float tu = (s0.x + (ax * u)) + (bx * v);
        float tv = (s0.y + (ay * u)) + (by * v);
        return lookupUV(S.texels + M->texOffset, M->texW, M->texH, tu, tv, F.addressMode, F.filtering);
    }
    return mk(s3.w, s4.x, s4.y);
}

// The reflection and refraction rays of a hit (:547-559, :656-698), appended to the next level's
// queue; returns the path-node indices of the two children (refraction: 0xFFFFFFFF when none).
__device__ __forceinline__ uint2 emitChildren(const DevScene& S, const DevArena& A, int level, unsigned nextBase, const Hit& h0,
                                              const DevMesh* __restrict__ M, V3 N, V3 rd0, float refIndex, unsigned& nRefl, unsigned& nRefr) {
    const int outq = (level & 1) ^ 1;
    // reflection ray (:547-559) — always cast, whatever the reflectiveness (Q9)
    V3 rdir = normalize(reflect(rd0, N));
    unsigned reflChild = appendRay(A, level, outq, nextBase, h0.pos, rdir, refIndex, h0.mesh, h0.tri);
    unsigned refrChild = 0xFFFFFFFFu;
    nRefl++;
    if (M->flags & MESH_TRANSPARENT) {
        // refraction (:656-698), double-precision island for the Snell term
        float n1, n2;
        if (refIndex == M->refractionIndex) { n1 = 1.0f; n2 = refIndex; }
        else { n1 = M->refractionIndex; n2 = 1.0f; }
        float cos1 = dot(N, neg(rd0));
        float k = n1 / n2;
        double ratio = (double)k, c1 = (double)cos1;
        float cos2 = (float)sqrt(1.0 - (ratio * ratio) * (1.0 - (c1 * c1)));
        float w = (k * cos1) - cos2;
        V3 refr = (cos1 >= 0.0f) ? add(mul(rd0, k), mul(N, w)) : sub(mul(rd0, k), mul(N, w));
        refrChild = appendRay(A, level, outq, nextBase, h0.pos, normalize(refr), n2, h0.mesh, h0.tri);
        nRefr++;
    }
    return make_uint2(reflChild, refrChild);
}

// Everything CastRay does after the light loop (:545-727): surface colour, the path node, and the
// reflection / refraction rays of the next level.
__device__ __forceinline__ void finishHit(const DevScene& S, const DevFrame& F, const DevArena& A, int level, bool last, unsigned node,
                                          unsigned nextBase, const Hit& h0, const DevMesh* __restrict__ M, int triG, V3 N,
                                          V3 lightResult, V3 rd0, float refIndex, unsigned& nRefl, unsigned& nRefr) {
    float alphaW;
    const V3 surf = surfaceColor(S, F, M, triG, h0.u, h0.v, alphaW);
    if (last) {
        A.color[node] = packColor(mulv(lightResult, surf), F.semantics);  // :726
        A.flags[node] = 0u;
        return;
    }
    const uint2 ch = emitChildren(S, A, level, nextBase, h0, M, N, rd0, refIndex, nRefl, nRefr);
    A.nodeA[node] = make_float4(lightResult.x, lightResult.y, lightResult.z, M->oneMinusReflectiveness);
    A.nodeB[node] = make_float4(surf.x, surf.y, surf.z, alphaW);
    A.nodeC[node] = ch;
    A.flags[node] = 1u;
}

// RayTracer.Progress = scanline / Height (RayTracer.cs:43-46).  A frame here is a sequence of recursion levels over a
// chunk of pixels, not a sweep over scanlines: level L of D+1 (plus the resolve) counts as L/(D+2) of the chunk.
__device__ __forceinline__ void reportProgress(const DevFrame& F, int level) {
    if (F.progress)
        *reinterpret_cast<volatile unsigned long long*>(F.progress) =
            F.progressBase + (unsigned long long)F.pixelBase + (unsigned long long)F.chunkPixels * (unsigned)level / (unsigned)(F.maxReflections + 2);
}

template <bool COUNT>
__device__ __forceinline__ void flushCounters(const DevArena& A, const Ctr& ctr, unsigned nHits, unsigned nRefl, unsigned nRefr) {
    // ray statistics: one atomic per warp and counter
    const unsigned lane = threadIdx.x & 31u;
    unsigned sH = warpSum(nHits), sR = warpSum(nRefl), sF = warpSum(nRefr);
    if (lane == 0) {
        if (sH) atomicAdd(&A.counters[CTR_HITS], (unsigned long long)sH);
        if (sR) atomicAdd(&A.counters[CTR_RAYS_REFLECTION], (unsigned long long)sR);
        if (sF) atomicAdd(&A.counters[CTR_RAYS_REFRACTION], (unsigned long long)sF);
    }
    if (COUNT) {
        unsigned a = warpSum(ctr.nBox), b = warpSum(ctr.nTriRef), c = warpSum(ctr.nTriCulled), d = warpSum(ctr.nTriFull);
        if (lane == 0) {
            atomicAdd(&A.counters[CTR_NBOX], (unsigned long long)a);
            atomicAdd(&A.counters[CTR_NTRIREF], (unsigned long long)b);
            atomicAdd(&A.counters[CTR_NTRICULLED], (unsigned long long)c);
            atomicAdd(&A.counters[CTR_NTRIFULL], (unsigned long long)d);
        }
    }
}

__device__ __forceinline__ void levelHeader(const DevFrame& F, const DevArena& A, int level, unsigned& count, unsigned& base) {
    count = (level == 0) ? (unsigned)F.chunkPixels : A.levelCount[level];
    base = (level == 0) ? 0u : A.levelBase[level];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (level == 0) { A.levelCount[0] = count; A.levelBase[0] = 0u; }
        A.levelBase[level + 1] = base + count;
        reportProgress(F, level);
    }
}

}  // namespace xn
