// device_types.h — HBM layout of an uploaded scene and of the per-render path arena.
//
// Octree nodes (MeshOctree.CubeNode, MeshOctree.cs:32-40; OctreeSpatialManager.CubeNode,
// OctreeSpatialManager.cs:37-48) of every mesh tree and of the scene tree live in ONE array in
// breadth order per tree; the 8 children of a node are consecutive (child c = 4i+2j+k,
// MeshOctree.cs:210-222), so one child block is 8 x 32 B = 256 contiguous bytes.
//
//   node n = { float4 lo = (Min.xyz, w0), float4 hi = (Max.xyz, w1) }
//     leaf     : w0 = first position in refs[] (global), w1 = item count (>= 0)
//     internal : w0 = global index of child 0,           w1 = -1 - consIndex (< 0)
//   consMax[consIndex] = component-wise max of Max over the node and all its descendants
//   (a child's Max = (Min + size*i) + size can overshoot its parent's Max by rounding,
//   MeshOctree.cs:217-218; pruning must use the conservative box, descending the node's own).
//
//   refs[] : leaf item lists concatenated in DEPTH-FIRST leaf order (children 0..7), so the
//            position of an item in refs[] reproduces the reference's visiting order inside a
//            group of equal-key leaves (MeshOctree.cs:283-288): ties in distance resolve to the
//            smallest position.  Mesh trees store mesh-relative triangle indices, the scene tree
//            stores object ids.
//
// Internal nodes of the MESH trees additionally get a 128-byte expansion record (one cache line,
// 8 x float4), indexed like consMax[]:
//   inner[8k+0] = x planes (lo0, hi0, lo1, hi1)   child (i,j,k) = 4i+2j+k spans x in [lo_i, hi_i]: the 8
//   inner[8k+1] = y planes (lo0, hi0, lo1, hi1)   children of a node are built from two halves per axis
//   inner[8k+2] = z planes (lo0, hi0, lo1, hi1)   (MeshOctree.cs:207-222), so 12 planes describe all 8 boxes
//   inner[8k+3] = (consMax.xyz, 0)                and one expansion costs 12 sub + 12 mul instead of 48 + 48
//   inner[8k+4..7] = 8 child descriptors (a, b):  leaf  (first position in refs[]/leafTri, count >= 0)
//                                                 inner (index of its own record, -1)
//   The node's own Min is (x.lo0, y.lo0, z.lo0): child 0 starts at Min + size*0.
//
// Triangles (Triangle.cs:14-25) are split into a 64-byte intersection record and an 80-byte
// shading record, both arrays of float4 read with 128-bit loads.
//   leafTri[4p+0] = (n.xyz, bits(triangle index))  [4p+1] = (v1.xyz, 0)  [4p+2] = (E1.xyz, 0)  [4p+3] = (E2.xyz, 0)
//       ONE RECORD PER LEAF REFERENCE, at the reference's own position p in refs[]: a leaf's
//       triangles are one contiguous run of 64-byte records, so the walk needs no dependent
//       index load (refs[p] -> triangle -> record) and can issue all four loads of a record at
//       once.  Triangles referenced by several leaves (vertex-containment binning,
//       MeshOctree.cs:226-228) are stored once per reference (about 2.2x on the 1M-triangle grid).
//       n = surfaceNormal; E1 = v2 - v1, E2 = v3 - v1 are computed once on the host in binary32 —
//       bit-identical to the per-test recomputation at RayExtensions.cs:54-55.
//   triS[5t+0] = (uv1.xy, uv2.xy) triS[5t+1] = (uv3.xy, n1.xy) triS[5t+2] = (n1.z, n2.xyz)
//   triS[5t+3] = (n3.xyz, color.x) triS[5t+4] = (color.yzw, 0)      (by triangle, read once per shaded hit)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace xn {

constexpr int kMaxLights = 16;
constexpr int kMeshStack = 160;   // >= 7*depth+1 -> mesh octree depth <= 22
constexpr int kSceneStack = 64;   // scene octree depth <= 9
constexpr int kMaxMeshDepth = (kMeshStack - 1) / 7;
constexpr int kMaxSceneDepth = (kSceneStack - 1) / 7;
constexpr int kBuilderDepthGuard = 64;  // same guard as the oracle: deeper -> XNART_ERR_OCTREE_DEPTH

struct DevMesh {
    float box[6];        // Mesh.MeshBoundingBox (Mesh.cs:34-39 prefilter)
    float rootBox[6];    // bounds of the MeshOctree root (tight box, MeshOctree.cs:59-80)
    int rootA, rootB;    // root descriptor: leaf (first ref, count) or inner (record index, -1)
    int rootNode;        // global node index of the MeshOctree root
    int triBase;         // first triangle of this mesh in triS
    float reflectiveness;
    float oneMinusReflectiveness;  // 1.0f - Reflectiveness (RayTracer.cs:584), hoisted
    float refractionIndex;
    int flags;           // bit0 useTexture, bit1 interpolateNormals, bit2 transparent
    int texOffset;       // first texel in texels[]
    int texW, texH;
    float overshoot;     // max over internal nodes and axes of (consMax - own Max) >= 0 (split queries: node filter slack)
    int _pad[8];         // 128 bytes
};
static_assert(sizeof(DevMesh) == 128, "DevMesh layout");
enum { MESH_USE_TEXTURE = 1, MESH_INTERP_NORMALS = 2, MESH_TRANSPARENT = 4 };

struct DevObj {
    float world[16];     // SceneObject.World (row-major; column 4 unused by Vector3.Transform)
    float inv[16];       // SceneObject.InverseWorld
    int meshStart, meshCount;  // range in objMeshIds[]
    int xformClass;            // index of the first object whose InverseWorld has the same bits
    int _pad;
};

struct DevLight {
    int kind;
    float px, py, pz, dx, dy, dz, cr, cg, cb;
    float intensity;
    float angleCosine;   // (float)Math.Cos(spotAngle * 0.5f), SpotLight.cs:25
    double powDen;       // Math.Pow(1 - angleCosine, DecayExponent), SpotLight.cs:54 (ray independent)
};

struct DevScene {
    const float4* nodes;
    const float4* consMax;
    const float4* inner;     // 8 x float4 per internal node of the mesh trees (expansion records)
    const int* refs;
    const float4* leafTri;
    const float4* triS;
    const DevMesh* meshes;
    const DevObj* objs;
    const int* objMeshIds;
    const uint32_t* texels;
    // (object, mesh) pairs in the reference's visiting order when the scene octree is a single leaf
    // (<= itemTreshold objects, OctreeSpatialManager.cs:103): list order of the leaf, then mesh order
    const int2* pairs;
    int nPairs;          // 0 when the scene octree has internal nodes (wavefront path unavailable)
    int sceneRoot;
    int nMeshes, nObjs;
};

// Result of one top-level GetRayIntersection query, two 16-byte words per ray slot:
//   out[2s]   = (object, mesh, triangle, position of the leaf reference), object = -1: miss
//   out[2s+1] = (u, v, d, 0) as float bits
// One batch of queries for the persistent trace kernel: nSeg segments of *segCount rays each,
// segment q occupying slots [q*segStride, q*segStride + *segCount).
struct TraceJob {
    const float4* rayO;      // (origin.xyz, -)
    const float4* rayD;      // (direction.xyz, -)
    const int2* rayIgn;      // (mesh, triangle) identity of the origin triangle, (-1,-1) = none
    int4* out;
    const unsigned int* segCount;
    unsigned int* fetch;     // work counter, zero at launch
    unsigned int nSeg, segStride;
    // Split queries: batches of at most splitMax rays are traced as kSplit segment queries per ray
    // (k_trace<.., true>; results in outSplit[(ray * kSplit + segment) * 2 ..]); larger batches by
    // k_trace<.., false>.  Both kernels are launched, the one that does not apply exits at once.
    int4* outSplit;
    unsigned int splitMax;   // 0: never split
    // Splitting pays when rays run their whole path without an accepted triangle (grazing a surface, leaving for the sky):
    // the segments are then work the whole-ray walk would have done anyway.  Where most rays hit early (a dense soup) the
    // segments behind the hit are wasted.  Hint: hits / rays of the closest-hit batch this one descends from (NULL: none);
    // the batch is split only if hits * 100 <= rays * splitHitPct.
    const unsigned int* hintHits;
    const unsigned int* hintRays;
    unsigned int splitHitPct;
};
constexpr int kSplit = 8;

struct DevFrame {
    float inv[16];       // Invert(Identity * view * proj), computed once per frame on the host
    int vpX, vpY, vpW, vpH;
    float minDepth, maxDepth;
    float ndcZ[2];          // (0 - minDepth) / (maxDepth - minDepth) and (1 - minDepth) / (maxDepth - minDepth), IEEE fp32 on the host
    int maxReflections, filtering, addressMode, semantics;
    int shardIndex, shardCount, shardBandRows;
    int nLights;
    // chunk of shard-local pixels handled by this launch sequence
    long long pixelBase;   // first shard-local pixel index of the chunk
    int chunkPixels;
    int tiled;             // 1: shard-local pixel indices run over 8x4 tiles (kernels.cu: pixelOf)
    // xnart_cast_rays: primary rays through arbitrary screen positions (x, y) instead of pixel centres;
    // "pixel" q is sample q and its colour goes to out[q]
    const float2* samples;
    // RayTracer.Progress (RayTracer.cs:43-46): a page-locked host word (mapped) the kernels store "pixels of this call
    // finished so far on this in-flight state" into; progressBase = pixels of the frames enqueued before this one
    unsigned long long* progress;
    unsigned long long progressBase;
};

// Per-render path arena (SoA over "path nodes": one node per traced primary/reflection/refraction ray).
struct DevArena {
    float4* nodeA;       // (lightResult.xyz, 1 - Reflectiveness)
    float4* nodeB;       // (surfaceColor.xyz, triangle.color.W)
    uint2* nodeC;        // (reflection child node, refraction child node or 0xFFFFFFFF)
    uint32_t* color;     // resolved RGBA8 of the node (Color struct, RayTracer.cs:705,726,732)
    uint32_t* flags;     // 0 = colour final, 1 = needs resolve
    // ray queues (ping-pong): rays of level L are read, rays of level L+1 appended
    float4* rayO[2];     // (origin.xyz, currentRefIndex)
    float4* rayD[2];     // (direction.xyz, unused)
    int2* rayIgn[2];     // origin triangle identity (mesh id, triangle index), Q6
    // level bookkeeping, all device resident so a frame needs no host round trip
    unsigned int* levelCount;   // [maxLevels+2] rays queued per level
    unsigned int* levelBase;    // [maxLevels+2] first node index of a level
    unsigned long long* counters;  // kCounterSlots
    unsigned int nodeCapacity, rayCapacity;
    unsigned int* overflow;
    // wavefront path (deep scenes): closest hits of the level's rays, and the shadow-ray batch of the level.
    // Double buffered by level parity: the shadow batch + shade B of level L run on a side stream while
    // the main stream already traces level L+1.
    int4* hit[2];        // 2 per ray slot of the level
    float4* shO[2];      // nLights segments of rayCapacity slots: (hit position, distance to the light)
    float4* shD[2];
    int2* shIgn[2];
    int4* shHit[2];      // 2 per shadow slot
    int4* hitSplit[2];   // split batches: kSplit records per ray of the level (<= splitMax rays)
    int4* shSplit[2];    // ... per shadow ray
    unsigned int splitMax;
    unsigned int splitHitPct;   // see TraceJob
    unsigned int* shadowHits;   // [maxLevels+2] hits per level (= rays per shadow segment)
    unsigned int* fetch;        // [2*(maxLevels+2)] work counters of the trace launches
};


// ---------------------------------------------------------------------------------------------
// Small scenes (the Free_crate configs): when the scene octree is a single leaf, every MeshOctree is
// a single leaf and no material is transparent, the whole intersection data set is a few hundred
// bytes.  It is then passed to the path kernel (path.cu) BY VALUE as a kernel parameter, i.e. it
// lives in the constant bank: all lanes of a warp walk the (object, mesh) pairs and their triangles
// in the same order, so every access is a uniform constant-bank operand — no load instruction, no
// address arithmetic, no register for the triangle data.
// ---------------------------------------------------------------------------------------------
constexpr int kSmallPairs = 6;    // (object, mesh) pairs in visiting order
constexpr int kSmallTris = 32;    // leaf references of all distinct meshes together
struct SmallPair {
    float inv[12];       // SceneObject.InverseWorld, rows 0..3 x columns 0..2: inv[3*r + c]
    float world[12];     // SceneObject.World, same packing
    float box[6];        // Mesh.MeshBoundingBox
    float rootBox[6];    // MeshOctree root bounds
    int obj, mesh;
    int cls;             // transform class (DevObj.xformClass)
    int first, count;    // run of the mesh's leaf in SmallScene.tri (4 float4 per reference)
    int refBase;         // global position of the run's first reference in refs[] / leafTri
    int boxCoversRoot;   // 1: rootBox lies inside box component-wise (the Mesh.RayIntersects prefilter cannot reject alone)
    int identity;        // 1: reference k of the run is triangle k of the mesh (the ignored triangle is found by position)
};
static_assert(sizeof(SmallPair) == 176, "SmallPair layout");
struct SmallScene {
    SmallPair pair[kSmallPairs];
    float4 tri[4 * (kSmallTris + 3)];   // same 64-byte records as leafTri; 3 zero records so that runs can be read in fours
    float sceneBox[6];            // bounds of the scene octree root
    int nPairs;
    int _pad;
};

enum {
    CTR_RAYS_PRIMARY = 0, CTR_RAYS_SHADOW, CTR_RAYS_REFLECTION, CTR_RAYS_REFRACTION, CTR_HITS,
    CTR_NBOX, CTR_NTRIREF, CTR_NTRICULLED, CTR_NTRIFULL, CTR_PIXELS, kCounterSlots = 16
};

}  // namespace xn
