/*
 * xnart.h — C ABI of libxnart.so, the B200-native per-pixel trace loop that replaces the
 * worker fan-out of RayTracer.RenderInternal (RayTraceProject/RayTraceProject/RayTracer.cs:103-126).
 *
 * The reference has no FFI seam on this path (its only DllImports are avifil32/winmm,
 * aviFileWrapper_src/Avi.cs:175-377), so this boundary is new; every entry point cites the
 * reference interface whose work it takes over.  All structs are plain-old-data with natural
 * alignment (no padding surprises: only 4-byte scalars and one pointer at the end of
 * XnartMaterial), all functions are extern "C", take plain pointers and sizes, and return an
 * int32 status (XNART_OK == 0).  Matrices are XNA row-major, row-vector convention
 * (M11..M44 in memory order, translation in M41..M43).
 *
 * Ownership: input pointers are borrowed for the duration of the call and copied; output
 * buffers are caller-allocated; the context owns all device memory.
 * Threading: a context may be used from any thread; calls that touch the device are
 * serialised per context, a render issued while another is in flight returns XNART_ERR_BUSY
 * (RayTracer.cs:26-27,62-63 throws InvalidOperationException); xnart_progress may be called
 * concurrently (Game1.cs:337 polls RayTracer.Progress from the game thread).
 */
#ifndef XNART_H
#define XNART_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------- */
enum {
    XNART_OK = 0,
    XNART_ERR_INVALID_ARGUMENT = 1, /* bad enum / null pointer -> ArgumentException (Material.cs:85,97) */
    XNART_ERR_BUSY = 2,             /* render in flight -> InvalidOperationException (RayTracer.cs:62-63) */
    XNART_ERR_CUDA = 3,             /* any CUDA runtime failure; text in xnart_last_error */
    XNART_ERR_NOT_BUILT = 4,        /* render/trace before xnart_build_scene */
    XNART_ERR_OCTREE_DEPTH = 5,     /* MeshOctree.BuildTree (MeshOctree.cs:84-96) would recurse without bound */
    XNART_ERR_NO_DEVICE = 6,        /* no usable CUDA device: there is no CPU fallback */
    XNART_ERR_OUT_OF_MEMORY = 7,
    XNART_ERR_UNSUPPORTED = 8
};

/* ---- enums mirrored from the reference ------------------------------------------------ */
enum { XNART_FILTER_POINT = 0, XNART_FILTER_BILINEAR = 1 };                /* Material.cs:12-16 TextureFiltering */
enum { XNART_ADDRESS_CLAMP = 0, XNART_ADDRESS_WRAP = 1, XNART_ADDRESS_MIRROR = 2 }; /* Material.cs:18-23 UVAddressMode */
enum { XNART_LIGHT_SPOT = 0, XNART_LIGHT_DIRECTIONAL = 1 };                /* SpotLight.cs / DirectionalLight.cs */
/* Semantics of the closed-source framework calls on the path:
 * XNA40 = Microsoft.Xna.Framework 4.0 (canonical; what the reference links,
 * RayTraceProject.csproj:65); MONOGAME = MonoGame's re-implementation. */
enum { XNART_SEMANTICS_XNA40 = 0, XNART_SEMANTICS_MONOGAME = 1 };

/* ---- scene records -------------------------------------------------------------------- */
/* Triangle.cs:14-25 (managed class, flattened by the shim at Mesh.Init time). */
typedef struct XnartTriangle {
    float v1[3], v2[3], v3[3];
    float uv1[2], uv2[2], uv3[2];
    float n1[3], n2[3], n3[3];
    float surfaceNormal[3];
    float color[4];
    int32_t id;
} XnartTriangle; /* 32 x 4 = 128 bytes */

/* Material.cs:28-39 + RayTracerTexture.cs:12-18 (ColorData: 0xAARRGGBB, top-down rows). */
typedef struct XnartMaterial {
    float reflectiveness;
    int32_t useTexture, interpolateNormals, transparent;
    float refractionIndex;
    int32_t texW, texH;
    int32_t _pad;
    const uint32_t* argb; /* texW*texH texels, borrowed; may be NULL when !useTexture */
} XnartMaterial;

/* SpotLight.cs:12-27, DirectionalLight.cs:11-14. */
typedef struct XnartLight {
    int32_t kind;
    float position[3], direction[3], color[3];
    float intensity, decayExponent, spotAngle;
} XnartLight;

/* RayTracer.cs:19-41 (properties) + :395-397 (view/proj/viewport captured per render). */
typedef struct XnartFrame {
    float view[16], proj[16];
    int32_t vpX, vpY, vpW, vpH;
    float minDepth, maxDepth;
    int32_t maxReflections; /* RayTracer.MaxReflections, RayTracer.cs:33 */
    int32_t filtering;      /* XNART_FILTER_*  */
    int32_t addressMode;    /* XNART_ADDRESS_* */
    int32_t semantics;      /* XNART_SEMANTICS_*: which framework's arithmetic the frame follows where XNA 4.0 and MonoGame
                               differ: the Color(Vector3) pack (round half to even / truncate) and
                               BoundingBox.Intersects(Ray) (reciprocal-multiply / divide form).  XNA 4.0 is what the reference
                               links (RayTraceProject.csproj:65) and what the fast kernels implement; MonoGame frames are
                               rendered by one generic kernel (a compatibility mode) */
    /* Row sharding for multi-GPU (RayTracer.cs:49-52 treats scanlines as independent work
     * items): this call renders only scanline bands b with (b % shardCount) == shardIndex,
     * a band being shardBandRows consecutive scanlines.  shardCount <= 1 renders everything. */
    int32_t shardIndex, shardCount, shardBandRows;
    int32_t _pad;
} XnartFrame;

/* Per-render statistics.  A "ray" is one call of the top-level GetRayIntersection
 * (RayTracer.cs:485,512).  Box/triangle counters are only filled when counting is enabled
 * (xnart_set_option(ctx, XNART_OPT_COUNTERS, 1)); ray counts are always filled. */
typedef struct XnartCounters {
    uint64_t raysPrimary, raysShadow, raysReflection, raysRefraction;
    uint64_t hits;        /* closest-hit queries that returned a hit (shaded fragments) */
    uint64_t nBox;        /* slab tests executed (mesh-level + scene-level + Mesh.RayIntersects) */
    uint64_t nTriRef;     /* leaf triangle references fetched */
    uint64_t nTriCulled;  /* triangle tests ended by the backface cull (RayExtensions.cs:50-51) */
    uint64_t nTriFull;    /* triangle tests run to the end */
    uint64_t pixels;
    double msTotal;       /* device time of the last render (CUDA events) */
    double msLevels;      /* ... of the trace+shade kernels */
    double msResolve;     /* ... of the bottom-up colour resolve */
    double msCopy;        /* ... of the device->host copy */
} XnartCounters;

enum {
    XNART_OPT_COUNTERS = 1,      /* 0/1: run the instrumented kernels (slower) */
    XNART_OPT_ARENA_BYTES = 2,   /* upper bound for the per-render path arena (default 8 GiB); larger frames are rendered in chunks.
                                    Scenes with transparent materials start from a small arena and grow it on overflow (blocking calls) */
    XNART_OPT_SEMANTICS = 3,     /* framework arithmetic of xnart_trace_rays: XNART_SEMANTICS_XNA40 (default) or _MONOGAME */
    XNART_OPT_TRAVERSAL = 4,     /* kernel schedule: 0 auto, 1 one thread per ray, 2 wavefront (persistent trace kernel) */
    XNART_OPT_FUSE_FRAME = 5,    /* 0/1 (default 0): shallow scenes render a frame in one cooperative launch */
    XNART_OPT_OVERLAP_LEVELS = 6, /* 0/1 (default 1): wavefront path runs a level's shadow batch beside the next level */
    XNART_OPT_CHAIN_RESOLVE = 8, /* 0/1 (default 1): opaque scenes fold a pixel's reflection chain in one launch */
    XNART_OPT_SPLIT_RAYS = 7,    /* wavefront path: batches of at most this many rays are traced as 8 segment queries per
                                    ray, which shortens the serial tail of deep levels (default 131072, 0 = never) */
    XNART_OPT_KEEP_COUNTERS = 10, /* 0/1 (default 0): 1 = render calls stop resetting the statistics, xnart_get_counters
                                    returns totals since the option was set (setting it, to 0 or 1, resets them) */
    XNART_OPT_SPLIT_HIT_PCT = 12, /* wavefront path: a batch is split only if the closest-hit batch it descends from hit at most this
                                    percentage of its rays (default 50): segments behind an early hit are wasted work in dense scenes */
    XNART_OPT_PATH_KERNEL = 9    /* 0/1 (default 1): small opaque scenes (single-leaf octrees that fit a kernel parameter,
                                    e.g. Free_crate) render a frame with ONE persistent launch: a thread follows a pixel's
                                    whole reflection chain, the scene is read from the constant bank (csrc/path.cu) */
};

typedef struct XnartCtx XnartCtx;

/* Context on one or more devices of this process (scene replicated on each).
 * Replaces: new RayTracer() (Game1.cs:119).
 * A device id may be listed several times: every entry is an independent in-flight state (stream,
 * path arena, scene copy) and xnart_render_frames deals whole frames round-robin over the entries,
 * so k entries for one GPU keep k frames in flight on it (their launch sequences fill each other's
 * gaps and tails: 26.9 vs 22.0 Grays/s on the 1080p crate flythrough with k = 2).  On deep scenes xnart_render_frame_device
 * also renders ONE frame as k sub-shards side by side.  xnart_render_frame splits a
 * single frame over distinct GPUs only. */
int32_t xnart_create(XnartCtx** out, int32_t nDevices, const int32_t* deviceIds);
int32_t xnart_destroy(XnartCtx* ctx);
const char* xnart_last_error(XnartCtx* ctx); /* ctx may be NULL (create failures) */
int32_t xnart_set_option(XnartCtx* ctx, int32_t option, int64_t value);

/* Mesh.Init() (Mesh.cs:27-32): builds the per-mesh octree natively, restating
 * MeshOctree.Build/BuildTree/SplitCuboid (MeshOctree.cs:56-96,204-236).  meshBoxMinMax is
 * Mesh.MeshBoundingBox (min xyz, max xyz); itemThreshold is MeshOctree.itemTreshold (50). */
int32_t xnart_add_mesh(XnartCtx* ctx, const XnartTriangle* tris, int64_t nTri,
                       const XnartMaterial* material, const float meshBoxMinMax[6],
                       int32_t itemThreshold, int32_t* meshIdOut);

/* new SceneObject + scene.Bodies.Add (Game1.cs:102-105): an instance of one or more meshes
 * with SceneObject.World / InverseWorld (SceneObject.cs:187-198) and the object's local
 * BoundingBox (SceneObject.cs:131). */
int32_t xnart_add_object(XnartCtx* ctx, const int32_t* meshIds, int32_t nMesh,
                         const float world[16], const float inverseWorld[16],
                         const float localBoxMinMax[6], int32_t* objIdOut);

/* scene.Bodies.Clear(): removes every object (the meshes stay, they are shared by all instances of a model).  A host
 * mirror calls it at the top of OctreeSpatialManager.Build(), which rebuilds from this.objects on every call
 * (OctreeSpatialManager.cs:64-99), so that calling Build() twice does not duplicate the scene. */
int32_t xnart_clear_objects(XnartCtx* ctx);

/* The reference reads SceneObject.World / InverseWorld per ray (OctreeSpatialManager.cs:358-364, 441-443): an object moved
 * after scene.Build() renders at its new pose, the scene octree keeps the membership it was built with.  Same here: the
 * matrices are replaced on every device of a built scene without rebuilding the octree. */
int32_t xnart_set_object_transform(XnartCtx* ctx, int32_t objId, const float world[16], const float inverseWorld[16]);

/* OctreeSpatialManager.Build (OctreeSpatialManager.cs:64-99); itemThreshold = 20.  Flattens
 * and uploads the scene to every device of the context. */
int32_t xnart_build_scene(XnartCtx* ctx, int32_t itemThreshold);

/* RayTracer.Lights (RayTracer.cs:36-39). */
int32_t xnart_set_lights(XnartCtx* ctx, const XnartLight* lights, int32_t n);

/* The body of RayTracer.RenderInternal (RayTracer.cs:103-126): fills renderTargetData
 * (vpW*vpH packed R | G<<8 | B<<16 | A<<24, row-major top-down) in host memory. */
int32_t xnart_render_frame(XnartCtx* ctx, const XnartFrame* frame, uint32_t* rgba8Out);

/* Same, but the output stays on the context's first device: devRgba8Out is a device pointer
 * to vpW*vpH uint32 and the work is enqueued on cudaStream (a cudaStream_t passed as void*,
 * NULL = the context's own stream, in which case the call returns after completion). */
int32_t xnart_render_frame_device(XnartCtx* ctx, const XnartFrame* frame,
                                  void* devRgba8Out, void* cudaStream);

/* A camera flythrough (Game1.cs:261-266 orbits the camera and renders frame after frame; the
 * reference would run RenderInternal + SetData n times): n frames of one viewport size into n
 * consecutive host frames of vpW*vpH pixels.  Frame f goes to device f % nDevices; on each device
 * the device->host copy of a frame overlaps the rendering of the next. */
int32_t xnart_render_frames(XnartCtx* ctx, const XnartFrame* frames, int32_t n, uint32_t* rgba8Out);

/* The ray casts of adaptive multisampling (RayTracer.GetColorForQuadrant, RayTracer.cs:215-311): for each
 * of n screen positions (x, y) — viewport pixel units, sub-pixel positions allowed — Unproject
 * (x, y, 0) and (x, y, 1), normalise, CastRay at iteration 0 (:220-276); rgba8Out[i] is the Color.
 * The refinement decision around the casts (:279-309) is host logic. */
int32_t xnart_cast_rays(XnartCtx* ctx, const XnartFrame* frame, const float* screenXY, int64_t n,
                        uint32_t* rgba8Out);

/* xnart_render_frames with the frames left in device memory (n consecutive vpW*vpH frames at
 * devRgba8Out), rendered by the in-flight states of the context's first GPU; ordered after the work
 * already enqueued on cudaStream, which is made to wait for all of it (NULL: the call returns after
 * completion).  For callers that gather the frames with a collective (one process per GPU). */
int32_t xnart_render_frames_device(XnartCtx* ctx, const XnartFrame* frames, int32_t n,
                                   void* devRgba8Out, void* cudaStream);

/* Device-memory utilities for one-process-per-GPU gathers of finished frames (no reference counterpart:
 * the reference is single-process).  Rank 0 allocates a receive buffer and exports a CUDA IPC handle; the
 * other ranks open it on their own device and copy their frames into it with xnart_copy_async on their
 * own streams (device-to-device over NVLink through the copy engines). */
void*   xnart_alloc_device(int32_t device, int64_t bytes);
void    xnart_free_device(int32_t device, void* devPtr);
int32_t xnart_ipc_export(int32_t device, const void* devPtr, uint8_t handle64[64]);
int32_t xnart_ipc_open(int32_t device, const uint8_t handle64[64], void** devPtrOut);
int32_t xnart_ipc_close(int32_t device, void* devPtr);
int32_t xnart_copy_async(int32_t device, void* dst, const void* src, int64_t bytes, void* cudaStream);

/* Batch form of ISpatialManager.GetRayIntersection (ISpatialManager.cs:15;
 * OctreeSpatialManager.cs:312-455) for parity testing: rays6 = n x (position xyz, direction
 * xyz); ignoreMeshTri2 = n x (meshId, triangleId) or NULL (-1,-1 = nothing ignored).
 * Misses report objId = meshId = triId = -1. worldPos3 may be NULL. */
int32_t xnart_trace_rays(XnartCtx* ctx, const float* rays6, int64_t n,
                         const int32_t* ignoreMeshTri2,
                         int32_t* objId, int32_t* meshId, int32_t* triId,
                         float* u, float* v, float* d, float* worldPos3);

int32_t xnart_get_counters(XnartCtx* ctx, XnartCounters* out);

/* RayTracer.Progress (RayTracer.cs:43-46: scanline / Height): fraction of the pixels of the blocking render call in
 * flight that are finished, reported by the kernels themselves (per pixel batch on small scenes, per recursion level and
 * arena chunk otherwise); 0 before the first render, 1 after a finished one.  Callable from any thread. */
float xnart_progress(XnartCtx* ctx);

/* Number of CUDA kernels launched by this context so far (for bench.py's gpu_launches). */
int64_t xnart_kernel_launches(XnartCtx* ctx);

/* VideoStream.AddFrame's bitmap preparation on the device (aviFileWrapper_src/VideoStream.cs:344-365: RotateNoneFlipY +
 * LockBits as 24bpp): nFrames consecutive width x height Color frames at devRgba8 -> nFrames DIBs at devDib, rows
 * bottom-up, 3 bytes B,G,R per pixel, every row padded to `stride` bytes (stride >= 3*width).  Enqueued on cudaStream. */
int32_t xnart_frames_to_dib24(int32_t device, const void* devRgba8, int64_t nFrames, int32_t width, int32_t height,
                              int32_t stride, void* devDib, void* cudaStream);

/* Measured ceilings of one GPU for the roofline bench.py reports (csrc/microbench.cu; no reference counterpart).
 * unitsPerSecond: operations/s for XNART_PEAK_FP32_NOFMA (FMUL + FADD without contraction — the arithmetic of this
 * library), bytes/s for the others; best of `reps` timed launches. */
enum { XNART_PEAK_FP32_NOFMA = 0, XNART_PEAK_L1_READ = 1, XNART_PEAK_L2_READ = 2, XNART_PEAK_HBM_READ = 3 };
int32_t xnart_measure_peak(int32_t device, int32_t which, int32_t reps, double* unitsPerSecond);

/* Version / build string (arch, flags). */
const char* xnart_version(void);

/* Page-locked host memory for renderTargetData (RayTracer.cs:16, the Color[] the render fills):
 * a frame buffer allocated here receives the device->host copy directly, without staging. */
void* xnart_alloc_host(int64_t bytes);
void xnart_free_host(void* p);

/* Structural export of the flattened octrees, for tests (host data only; also available in a
 * host-only context created with nDevices == 0).  which: 0 = MeshOctree of mesh `id`
 * (MeshOctree.cs:56-96), 1 = the scene octree (OctreeSpatialManager.cs:64-113).  Nodes are in
 * breadth order, the 8 children of a node consecutive; leaf: w0 = first reference, w1 = count;
 * internal: w0 = first child, w1 = -1 - index into consMax3; refs = leaf lists in depth-first
 * leaf order. */
int32_t xnart_debug_tree_info(XnartCtx* ctx, int32_t which, int32_t id, int32_t* nNodes,
                              int32_t* nRefs, int32_t* nCons, int32_t* depth);
int32_t xnart_debug_tree_dump(XnartCtx* ctx, int32_t which, int32_t id, float* lo3, float* hi3,
                              int32_t* w0, int32_t* w1, float* consMax3, int32_t* refs);
/* The 128-byte node expansion records of mesh `id` (32 floats per internal node: 12 child planes, consMax,
 * 8 child descriptors) and the global bases its record indices / leaf reference positions are relative to. */
int32_t xnart_debug_inner_dump(XnartCtx* ctx, int32_t id, float* records32, int32_t* innerBase, int32_t* refBase);

#ifdef __cplusplus
}
#endif
#endif /* XNART_H */
