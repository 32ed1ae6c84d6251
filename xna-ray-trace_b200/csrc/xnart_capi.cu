// xnart_capi.cu — the C ABI of include/xnart.h: context, scene upload, frame rendering.
//
// Replaces the worker fan-out of RayTracer.RenderInternal (RayTracer.cs:103-126) with one call.
// There is NO CPU fallback: every compute entry point needs a CUDA device and fails with
// XNART_ERR_NO_DEVICE / XNART_ERR_CUDA otherwise.  A context created with nDevices == 0 is
// "host only": it can build and export the flattened octrees (for structural tests) and nothing else.
#include <atomic>
#include <cmath>
#include <memory>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/xnart.h"
#include "device_types.h"
#include "kernels.h"
#include "scene_build.h"
#include "xna_math.cuh"

using namespace xn;

namespace {

struct DeviceState {
    int dev = -1;
    int smCount = 0;
    cudaStream_t stream = nullptr;
    // scene
    float4* nodes = nullptr; float4* consMax = nullptr; float4* inner = nullptr; int* refs = nullptr;
    float4* leafTri = nullptr; float4* triS = nullptr;
    DevMesh* meshes = nullptr; DevObj* objs = nullptr; int* objMeshIds = nullptr; uint32_t* texels = nullptr;
    DevLight* lights = nullptr; int nLightsUploaded = -1;
    DevScene scene{};
    // arena
    DevArena arena{};
    void* arenaBlock = nullptr; size_t arenaBytes = 0;
    unsigned arenaNodes = 0, arenaRays = 0;
    unsigned* levelWords = nullptr;  // levelCount | levelBase | overflow
    unsigned long long* counters = nullptr;
    unsigned long long* stamps = nullptr;   // k_frame: global timer at start / levels done / resolve done
    int frameBlocks = 0, frameBlocksCount = 0;  // resident grid of the cooperative frame kernel
    int traceBlocks = 0, traceBlocksCount = 0;  // resident grid of the persistent trace kernel
    int pathBlocks = 0, pathBlocksCount = 0;    // resident grid of the persistent path kernel (path.cu)
    bool lastPath = false;
    // RayTracer.Progress (RayTracer.cs:43-46): a page-locked word the kernels write through a mapped pointer
    unsigned char* countersHost = nullptr;      // page-locked snapshot of counters + level words, copied behind a synchronous frame (collectCounters)
    bool countersPrefetched = false;
    unsigned long long* progressHost = nullptr; // [0] = pixels of the call in flight finished on this state, as the kernels last reported
    unsigned long long progressEnqueued = 0;    // pixels of the frames enqueued on this state so far in the call in flight
    // scratch for the batch entry points (xnart_trace_rays / xnart_cast_rays): grows, never shrinks
    void* scratch = nullptr; size_t scratchBytes = 0;
    int2* pairs = nullptr;
    bool lastWave = false;
    cudaStream_t sideStream = nullptr;                 // shadow batch + shade B of a level (wavefront path)
    std::vector<cudaEvent_t> evA, evB;                  // per level: shade A done (main) / shade B done (side)
    bool lastFused = false;
    // output
    uint32_t* frameBuf = nullptr; size_t frameBufPixels = 0;
    // flythrough pipeline: second frame buffer, copy stream, per-buffer events
    uint32_t* frameBuf2 = nullptr; size_t frameBuf2Pixels = 0;
    cudaStream_t copyStream = nullptr;
    cudaEvent_t evRendered[2] = {nullptr, nullptr}, evCopied[2] = {nullptr, nullptr};
    uint32_t* pinned = nullptr; size_t pinnedPixels = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // start, levels done, resolve done, copy done
    cudaStream_t lastStream = nullptr;
    bool countersPending = false;
    bool singleChunk = true;
    long long lastPixels = 0;
    int lastLights = 0;
};

constexpr int kMaxLevels = 30;
// levelCount | levelBase | overflow(2) | shadowHits | fetch(2 per level)
constexpr int kLevelWords = 5 * (kMaxLevels + 2) + 2;
constexpr int kOvfWord = 2 * (kMaxLevels + 2);   // level word that flags a ray-queue overflow
constexpr size_t kCounterSnapBytes = sizeof(unsigned long long) * 16 + sizeof(unsigned) * (kOvfWord + 2);   // counters + level words up to the flag (+ 1: a multiple of 8 bytes)

inline float __int_as_float_host(int v) {
    float f;
    std::memcpy(&f, &v, 4);
    return f;
}

}  // namespace

struct XnartCtx {
    std::vector<DeviceState> devs;
    std::vector<HostMesh> meshes;
    std::vector<HostObj> objs;
    FlatTree sceneTree;
    std::vector<XnartLight> lights;
    bool built = false;
    bool lightsDirty = true;
    bool hasTransparent = false;
    std::mutex renderMutex;
    std::atomic<float> progress{0.0f};          // between calls: 0 before the first render, 1 after a finished one
    std::atomic<bool> rendering{false};         // a blocking render call is in flight: xnart_progress reads the devices' words
    std::atomic<unsigned long long> progressTotal{0};   // pixels of the call in flight
    std::vector<DevObj> hostObjs;               // DevObj records as uploaded (xnart_set_object_transform refreshes them)
    std::string err;
    bool countKernels = false;
    int64_t arenaLimit = (int64_t)8 << 30;
    int traceSemantics = XNART_SEMANTICS_XNA40;
    bool chainResolve = true;   // XNART_OPT_CHAIN_RESOLVE: opaque scenes resolve a pixel's chain in one launch
    int64_t splitHitPct = 50;   // XNART_OPT_SPLIT_HIT_PCT: split only batches whose parent batch hit at most this share of its rays
    int64_t splitRays = 131072;  // XNART_OPT_SPLIT_RAYS: batches up to this many rays are traced as segment queries (0 = never)
    bool overlapLevels = true;  // XNART_OPT_OVERLAP_LEVELS: side stream for the shadow batches of the wavefront path
    // XNART_OPT_FUSE_FRAME: one cooperative launch per frame on shallow scenes.  Off by default: measured
    // on B200 the grid-wide barriers cost more than the 17 launches they replace (0.63 vs 0.61 ms per 1080p
    // crate frame), and a cooperative launch cannot share the GPU with NCCL's gather kernels.
    bool fuseFrame = false;
    int traversalMode = 0;   // XNART_OPT_TRAVERSAL: 0 auto, 1 one thread per ray, 2 one octet per ray
    int maxMeshDepth = 0;
    int nPairs = 0;          // > 0: the scene octree is a single leaf (wavefront path available)
    SmallScene small{};      // the whole intersection data set as a kernel parameter (path.cu), when it fits
    bool smallOk = false;
    // Transparent materials can fork every hit into reflection + refraction (RayTracer.cs:586-702): the worst case is a full
    // binary tree per pixel, 2^(D+1)-1 path nodes — 511 at the reference's depth of 8, which cut a 1080p frame of the default
    // 4-sphere scene into 11 arena chunks of ~18 launches each although only 3 % of its pixels hit anything.  Blocking render
    // calls therefore size the arena for arenaScale x (D+1) nodes per pixel, detect an overflow on the device, and render the
    // frame again with four times the room (the scale sticks to the context); calls on a caller's stream cannot look at the
    // flag before they return and keep the worst-case size.
    double arenaScale = 2.0;
    bool optimisticArena = false;   // set by the blocking entry points around their enqueue
    bool overflowed = false;        // collectCounters saw the overflow flag
    bool keepCounters = false;  // XNART_OPT_KEEP_COUNTERS: statistics accumulate over render calls
    bool pathKernel = true;  // XNART_OPT_PATH_KERNEL: small opaque scenes render a frame with one persistent launch
    std::vector<float> hostInner;            // expansion records as uploaded (32 floats per internal node)
    std::vector<int> meshInnerBase, meshRefBase;
    int64_t totalRefs = 0;   // leaf references of all mesh octrees
    std::atomic<int64_t> launches{0};
    XnartCounters last{};
};

static thread_local std::string g_createError;

static int fail(XnartCtx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_createError = msg;
    return code;
}

#define CU_TRY(ctx, expr)                                                                                  \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess) {                                                                           \
            char _b[512];                                                                                  \
            snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return fail(ctx, _e == cudaErrorMemoryAllocation ? XNART_ERR_OUT_OF_MEMORY : XNART_ERR_CUDA, _b); \
        }                                                                                                  \
    } while (0)

static void freeScene(DeviceState& d) {
    cudaSetDevice(d.dev);
    cudaFree(d.nodes); cudaFree(d.consMax); cudaFree(d.inner); cudaFree(d.pairs); cudaFree(d.refs); cudaFree(d.leafTri); cudaFree(d.triS);
    cudaFree(d.meshes); cudaFree(d.objs); cudaFree(d.objMeshIds); cudaFree(d.texels);
    d.nodes = d.consMax = d.inner = d.leafTri = d.triS = nullptr; d.pairs = nullptr; d.refs = nullptr; d.meshes = nullptr; d.objs = nullptr;
    d.objMeshIds = nullptr; d.texels = nullptr;
}

template <class T>
static cudaError_t uploadVec(T** dst, const std::vector<T>& v) {
    *dst = nullptr;
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    cudaError_t e = cudaMalloc(dst, bytes);
    if (e != cudaSuccess) return e;
    if (!v.empty()) e = cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}

extern "C" {

const char* xnart_version(void) { return "libxnart 0.1 (sm_100a, -fmad=false, XNA 4.0 semantics)"; }

const char* xnart_last_error(XnartCtx* ctx) { return ctx ? ctx->err.c_str() : g_createError.c_str(); }

int32_t xnart_create(XnartCtx** out, int32_t nDevices, const int32_t* deviceIds) {
    if (!out || nDevices < 0) return fail(nullptr, XNART_ERR_INVALID_ARGUMENT, "xnart_create: bad arguments");
    *out = nullptr;
    XnartCtx* c = new XnartCtx();
    if (nDevices > 0) {
        int avail = 0;
        cudaError_t e = cudaGetDeviceCount(&avail);
        if (e != cudaSuccess || avail <= 0) {
            delete c;
            return fail(nullptr, XNART_ERR_NO_DEVICE,
                        std::string("no usable CUDA device (") + cudaGetErrorString(e) + "): libxnart has no CPU fallback");
        }
        for (int i = 0; i < nDevices; i++) {
            int id = deviceIds ? deviceIds[i] : i;
            if (id < 0 || id >= avail) { delete c; return fail(nullptr, XNART_ERR_INVALID_ARGUMENT, "xnart_create: device id out of range"); }
            DeviceState d;
            d.dev = id;
            if (cudaSetDevice(id) != cudaSuccess || cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) != cudaSuccess) {
                delete c;
                return fail(nullptr, XNART_ERR_CUDA, "xnart_create: cannot initialise device");
            }
            cudaDeviceGetAttribute(&d.smCount, cudaDevAttrMultiProcessorCount, id);
            for (auto& ev : d.ev) cudaEventCreate(&ev);
            cudaStreamCreateWithFlags(&d.copyStream, cudaStreamNonBlocking);
            cudaStreamCreateWithFlags(&d.sideStream, cudaStreamNonBlocking);
            d.evA.resize(kMaxLevels + 2); d.evB.resize(kMaxLevels + 2);
            for (auto& e : d.evA) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            for (auto& e : d.evB) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            for (int k = 0; k < 2; k++) {
                cudaEventCreateWithFlags(&d.evRendered[k], cudaEventDisableTiming);
                cudaEventCreateWithFlags(&d.evCopied[k], cudaEventDisableTiming);
            }
            // counters and level words share one block so one copy reads the counters and the overflow flag
            cudaMalloc(&d.counters, sizeof(unsigned long long) * kCounterSlots + sizeof(unsigned) * kLevelWords);
            d.levelWords = reinterpret_cast<unsigned*>(d.counters + kCounterSlots);
            cudaMalloc(&d.stamps, sizeof(unsigned long long) * 4);
            int coopOk = 0;
            cudaDeviceGetAttribute(&coopOk, cudaDevAttrCooperativeLaunch, id);
            if (coopOk) { d.frameBlocks = frameKernelMaxBlocks(false); d.frameBlocksCount = frameKernelMaxBlocks(true); }
            d.traceBlocks = traceKernelBlocksPerSm(false) * d.smCount; d.traceBlocksCount = traceKernelBlocksPerSm(true) * d.smCount;
            d.pathBlocks = pathKernelBlocksPerSm(false) * d.smCount; d.pathBlocksCount = pathKernelBlocksPerSm(true) * d.smCount;
            if (cudaHostAlloc((void**)&d.progressHost, 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); d.progressHost = nullptr; }
            if (d.progressHost) std::memset(d.progressHost, 0, 64);
            if (cudaHostAlloc((void**)&d.countersHost, kCounterSnapBytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); d.countersHost = nullptr; }
            cudaMemset(d.counters, 0, sizeof(unsigned long long) * kCounterSlots);
            c->devs.push_back(d);
        }
    }
    *out = c;
    return XNART_OK;
}

int32_t xnart_destroy(XnartCtx* c) {
    if (!c) return XNART_OK;
    for (auto& d : c->devs) {
        cudaSetDevice(d.dev);
        cudaStreamSynchronize(d.stream);
        freeScene(d);
        cudaFree(d.lights); cudaFree(d.arenaBlock); cudaFree(d.counters); cudaFree(d.stamps); cudaFree(d.frameBuf);
        if (d.pinned) cudaFreeHost(d.pinned);
        if (d.progressHost) cudaFreeHost(d.progressHost);
        if (d.countersHost) cudaFreeHost(d.countersHost);
        cudaFree(d.scratch);
        for (auto& ev : d.ev) if (ev) cudaEventDestroy(ev);
        for (int k = 0; k < 2; k++) { if (d.evRendered[k]) cudaEventDestroy(d.evRendered[k]); if (d.evCopied[k]) cudaEventDestroy(d.evCopied[k]); }
        if (d.copyStream) cudaStreamDestroy(d.copyStream);
        if (d.sideStream) { cudaStreamSynchronize(d.sideStream); cudaStreamDestroy(d.sideStream); }
        for (auto& e : d.evA) if (e) cudaEventDestroy(e);
        for (auto& e : d.evB) if (e) cudaEventDestroy(e);
        cudaFree(d.frameBuf2);
        cudaStreamDestroy(d.stream);
    }
    delete c;
    return XNART_OK;
}

int32_t xnart_set_option(XnartCtx* c, int32_t option, int64_t value) {
    if (!c) return XNART_ERR_INVALID_ARGUMENT;
    switch (option) {
        case XNART_OPT_COUNTERS: c->countKernels = value != 0; return XNART_OK;
        case XNART_OPT_ARENA_BYTES:
            if (value < (1 << 20)) return fail(c, XNART_ERR_INVALID_ARGUMENT, "arena limit too small");
            c->arenaLimit = value; return XNART_OK;
        case XNART_OPT_SEMANTICS:
            if (value != XNART_SEMANTICS_XNA40 && value != XNART_SEMANTICS_MONOGAME) return fail(c, XNART_ERR_INVALID_ARGUMENT, "unknown semantics");
            c->traceSemantics = (int)value; return XNART_OK;
        case XNART_OPT_FUSE_FRAME: c->fuseFrame = value != 0; return XNART_OK;
        case XNART_OPT_OVERLAP_LEVELS: c->overlapLevels = value != 0; return XNART_OK;
        case XNART_OPT_CHAIN_RESOLVE: c->chainResolve = value != 0; return XNART_OK;
        case XNART_OPT_SPLIT_RAYS:
            if (value < 0 || value > (1 << 24)) return fail(c, XNART_ERR_INVALID_ARGUMENT, "split threshold out of range");
            c->splitRays = value; return XNART_OK;
        case XNART_OPT_PATH_KERNEL: c->pathKernel = value != 0; return XNART_OK;
        case XNART_OPT_KEEP_COUNTERS: {
            std::lock_guard<std::mutex> lk(c->renderMutex);
            c->keepCounters = value != 0;
            for (DeviceState& d : c->devs) {
                if (cudaSetDevice(d.dev) != cudaSuccess) continue;
                cudaStreamSynchronize(d.stream);
                if (d.lastStream) cudaStreamSynchronize(d.lastStream);
                cudaMemset(d.counters, 0, sizeof(unsigned long long) * kCounterSlots);
                d.lastPixels = 0; d.countersPending = false;
            }
            c->last = XnartCounters{};
            return XNART_OK;
        }
        case XNART_OPT_SPLIT_HIT_PCT:
            if (value < 0 || value > 100) return fail(c, XNART_ERR_INVALID_ARGUMENT, "percentage out of range");
            c->splitHitPct = value; return XNART_OK;
        case XNART_OPT_TRAVERSAL:
            if (value < 0 || value > 2) return fail(c, XNART_ERR_INVALID_ARGUMENT, "traversal mode must be 0, 1 or 2");
            c->traversalMode = (int)value; return XNART_OK;
        default: return fail(c, XNART_ERR_INVALID_ARGUMENT, "unknown option");
    }
}

int32_t xnart_add_mesh(XnartCtx* c, const XnartTriangle* tris, int64_t nTri, const XnartMaterial* material,
                       const float box[6], int32_t itemThreshold, int32_t* meshIdOut) {
    if (!c || (!tris && nTri > 0) || nTri < 0 || !material || !box || !meshIdOut || itemThreshold < 0)
        return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_add_mesh: bad arguments");
    if (nTri > 0x7fffffffLL / 8) return fail(c, XNART_ERR_UNSUPPORTED, "xnart_add_mesh: too many triangles");
    if (material->useTexture && (!material->argb || material->texW <= 0 || material->texH <= 0))
        return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_add_mesh: useTexture without texels");
    // NaN coordinates are refused: the slab tests on the device use min / max instructions, which equal the
    // framework's compare-and-select forms only for operands that are not NaN (trace.cuh: slabAxis)
    for (int a = 0; a < 6; a++) if (box[a] != box[a]) return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_add_mesh: NaN in meshBoxMinMax");
    {
        bool nan = false;
#pragma omp parallel for reduction(|| : nan) if (nTri > 65536)
        for (long long i = 0; i < (long long)nTri; i++)
            for (int k = 0; k < 3; k++)
                if (tris[i].v1[k] != tris[i].v1[k] || tris[i].v2[k] != tris[i].v2[k] || tris[i].v3[k] != tris[i].v3[k]) nan = true;
        if (nan) return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_add_mesh: NaN vertex coordinate");
    }
    std::unique_lock<std::mutex> lk(c->renderMutex, std::try_to_lock);
    if (!lk.owns_lock()) return fail(nullptr, XNART_ERR_BUSY, "a render is already in flight");   // RayTracer.cs:62-63
    HostMesh m;
    m.nTri = nTri;
    std::memcpy(m.box.mn, box, 12);
    std::memcpy(m.box.mx, box + 3, 12);
    m.reflectiveness = material->reflectiveness;
    m.refractionIndex = material->refractionIndex;
    m.flags = (material->useTexture ? MESH_USE_TEXTURE : 0) | (material->interpolateNormals ? MESH_INTERP_NORMALS : 0) |
              (material->transparent ? MESH_TRANSPARENT : 0);
    if (material->useTexture) {
        m.texW = material->texW; m.texH = material->texH;
        m.texels.assign(material->argb, material->argb + (size_t)m.texW * (size_t)m.texH);
    }
    int rc = buildMeshOctree(tris, nTri, itemThreshold, m.tree, c->err);
    if (rc != XNART_OK) return rc;
    packTriangles(tris, nTri, m.triI, m.triS);
    *meshIdOut = (int32_t)c->meshes.size();
    c->meshes.push_back(std::move(m));
    c->built = false;
    return XNART_OK;
}

int32_t xnart_add_object(XnartCtx* c, const int32_t* meshIds, int32_t nMesh, const float world[16],
                         const float inverseWorld[16], const float localBox[6], int32_t* objIdOut) {
    if (!c || !meshIds || nMesh < 0 || !world || !inverseWorld || !localBox || !objIdOut)
        return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_add_object: bad arguments");
    for (int a = 0; a < 16; a++)
        if (world[a] != world[a] || inverseWorld[a] != inverseWorld[a]) return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_add_object: NaN in a matrix");
    for (int a = 0; a < 6; a++) if (localBox[a] != localBox[a]) return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_add_object: NaN in localBoxMinMax");
    std::unique_lock<std::mutex> lk(c->renderMutex, std::try_to_lock);
    if (!lk.owns_lock()) return fail(nullptr, XNART_ERR_BUSY, "a render is already in flight");   // RayTracer.cs:62-63
    HostObj o;
    for (int i = 0; i < nMesh; i++) {
        if (meshIds[i] < 0 || meshIds[i] >= (int)c->meshes.size()) return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_add_object: bad mesh id");
        o.meshIds.push_back(meshIds[i]);
    }
    std::memcpy(o.world, world, 64);
    std::memcpy(o.inv, inverseWorld, 64);
    std::memcpy(o.localBox.mn, localBox, 12);
    std::memcpy(o.localBox.mx, localBox + 3, 12);
    *objIdOut = (int32_t)c->objs.size();
    c->objs.push_back(o);
    c->built = false;
    return XNART_OK;
}

// scene.Bodies.Clear(): the meshes (Model.Tag["Meshes"], shared by all instances) stay, the objects go.  Lets a host
// mirror make OctreeSpatialManager.Build() idempotent — the reference rebuilds its tree from this.objects on every call
// (OctreeSpatialManager.cs:64-99) — instead of adding every body a second time.
int32_t xnart_clear_objects(XnartCtx* c) {
    if (!c) return XNART_ERR_INVALID_ARGUMENT;
    std::unique_lock<std::mutex> lk(c->renderMutex, std::try_to_lock);
    if (!lk.owns_lock()) return fail(nullptr, XNART_ERR_BUSY, "a render is already in flight");
    c->objs.clear();
    c->built = false;
    return XNART_OK;
}

// The reference reads SceneObject.World / InverseWorld per ray (OctreeSpatialManager.cs:358-364, 441-443), so an object
// moved after scene.Build() renders at its new pose while the scene octree keeps the leaf membership it was built with.
// Same here: the matrices of a built scene are replaced on every device, the octree is left alone.
int32_t xnart_set_object_transform(XnartCtx* c, int32_t objId, const float world[16], const float inverseWorld[16]) {
    if (!c || !world || !inverseWorld) return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_set_object_transform: bad arguments");
    for (int a = 0; a < 16; a++)
        if (world[a] != world[a] || inverseWorld[a] != inverseWorld[a]) return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_set_object_transform: NaN in a matrix");
    std::unique_lock<std::mutex> lk(c->renderMutex, std::try_to_lock);
    if (!lk.owns_lock()) return fail(nullptr, XNART_ERR_BUSY, "a render is already in flight");
    if (objId < 0 || objId >= (int)c->objs.size()) return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_set_object_transform: bad object id");
    HostObj& o = c->objs[(size_t)objId];
    std::memcpy(o.world, world, 64);
    std::memcpy(o.inv, inverseWorld, 64);
    if (!c->built || c->hostObjs.size() != c->objs.size()) return XNART_OK;   // picked up by the next xnart_build_scene
    std::vector<DevObj>& dobjs = c->hostObjs;
    for (size_t i = 0; i < c->objs.size(); i++) {
        DevObj& d = dobjs[i];
        std::memcpy(d.world, c->objs[i].world, 64);
        std::memcpy(d.inv, c->objs[i].inv, 64);
        d.xformClass = (int)i;
        for (size_t j = 0; j < i; j++)
            if (std::memcmp(c->objs[j].inv, c->objs[i].inv, 64) == 0) { d.xformClass = dobjs[j].xformClass; break; }
    }
    if (c->smallOk)
        for (int p = 0; p < c->small.nPairs; p++) {
            SmallPair& P = c->small.pair[p];
            const DevObj& dob = dobjs[(size_t)P.obj];
            for (int r = 0; r < 4; r++)
                for (int col = 0; col < 3; col++) { P.inv[3 * r + col] = dob.inv[4 * r + col]; P.world[3 * r + col] = dob.world[4 * r + col]; }
            P.cls = dob.xformClass;
        }
    for (DeviceState& d : c->devs) {
        CU_TRY(c, cudaSetDevice(d.dev));
        CU_TRY(c, cudaStreamSynchronize(d.stream));
        if (d.lastStream) CU_TRY(c, cudaStreamSynchronize(d.lastStream));
        CU_TRY(c, cudaMemcpy(d.objs, dobjs.data(), dobjs.size() * sizeof(DevObj), cudaMemcpyHostToDevice));
    }
    return XNART_OK;
}

int32_t xnart_build_scene(XnartCtx* c, int32_t itemThreshold) {
    if (!c || itemThreshold < 0) return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_build_scene: bad arguments");
    std::unique_lock<std::mutex> lk(c->renderMutex, std::try_to_lock);
    if (!lk.owns_lock()) return fail(nullptr, XNART_ERR_BUSY, "a render is already in flight");   // RayTracer.cs:62-63
    c->built = false;   // stays false if anything below fails: a half-uploaded scene must never be rendered
    int rc = buildSceneOctree(c->objs, itemThreshold, c->sceneTree, c->err);
    if (rc != XNART_OK) return rc;
    c->hasTransparent = false;
    c->maxMeshDepth = 0;
    for (const HostMesh& m : c->meshes) {
        c->maxMeshDepth = std::max(c->maxMeshDepth, m.tree.depth);
        if (m.flags & MESH_TRANSPARENT) c->hasTransparent = true;
        if (m.tree.depth > kMaxMeshDepth) return fail(c, XNART_ERR_UNSUPPORTED, "mesh octree deeper than the device traversal stack allows");
    }
    if (c->sceneTree.depth > kMaxSceneDepth) return fail(c, XNART_ERR_UNSUPPORTED, "scene octree deeper than the device traversal stack allows");

    // ---- flatten: all mesh trees, then the scene tree, in one node array -------------------------
    size_t nNodes = c->sceneTree.nodeCount(), nCons = c->sceneTree.consMax.size() / 3, nRefs = c->sceneTree.refs.size();
    size_t nTri = 0, nTex = 0, nObjMesh = 0;
    for (const HostMesh& m : c->meshes) {
        nNodes += m.tree.nodeCount(); nCons += m.tree.consMax.size() / 3; nRefs += m.tree.refs.size();
        nTri += (size_t)m.nTri; nTex += m.texels.size();
    }
    for (const HostObj& o : c->objs) nObjMesh += o.meshIds.size();
    if (nNodes > 0x3fffffffULL || nRefs > 0x7fffffffULL || nTri > 0x7fffffffULL / 8)
        return fail(c, XNART_ERR_UNSUPPORTED, "scene too large for 32-bit node/reference indices");
    size_t nMeshRefs = nRefs - c->sceneTree.refs.size();
    std::vector<float4> nodes(2 * nNodes), cons(nCons), inner(8 * nCons), leafTri(4 * nMeshRefs), triS(5 * nTri);
    bool planesShared = true;
    std::vector<int> refs(nRefs), objMeshIds(nObjMesh);
    std::vector<uint32_t> texels(nTex);
    std::vector<DevMesh> dmeshes(c->meshes.size());
    std::vector<DevObj> dobjs(c->objs.size());
    size_t nodeBase = 0, consBase = 0, refBase = 0, triBase = 0, texBase = 0;
    auto emitTree = [&](const FlatTree& t) {
        for (size_t n = 0; n < t.nodeCount(); n++) {
            int w0 = t.w0[n], w1 = t.w1[n];
            if (w1 >= 0) w0 += (int)refBase;                                   // leaf: global ref position
            else { w0 += (int)nodeBase; w1 = -1 - ((-1 - w1) + (int)consBase); }  // internal: global child / cons index
            nodes[2 * (nodeBase + n)] = make_float4(t.lo[3 * n], t.lo[3 * n + 1], t.lo[3 * n + 2], __int_as_float_host(w0));
            nodes[2 * (nodeBase + n) + 1] = make_float4(t.hi[3 * n], t.hi[3 * n + 1], t.hi[3 * n + 2], __int_as_float_host(w1));
        }
        for (size_t k = 0; k < t.consMax.size() / 3; k++)
            cons[consBase + k] = make_float4(t.consMax[3 * k], t.consMax[3 * k + 1], t.consMax[3 * k + 2], 0.0f);
        // expansion records: the 12 planes of the 8 children + consMax + child descriptors
        for (size_t n = 0; n < t.nodeCount(); n++) {
            if (t.w1[n] >= 0) continue;
            const size_t k = (size_t)(-1 - t.w1[n]), f = (size_t)t.w0[n];
            float4* R = &inner[8 * (consBase + k)];
            const size_t cx = f + 4, cy = f + 2, cz = f + 1;
            R[0] = make_float4(t.lo[3 * f + 0], t.hi[3 * f + 0], t.lo[3 * cx + 0], t.hi[3 * cx + 0]);
            R[1] = make_float4(t.lo[3 * f + 1], t.hi[3 * f + 1], t.lo[3 * cy + 1], t.hi[3 * cy + 1]);
            R[2] = make_float4(t.lo[3 * f + 2], t.hi[3 * f + 2], t.lo[3 * cz + 2], t.hi[3 * cz + 2]);
            R[3] = make_float4(t.consMax[3 * k], t.consMax[3 * k + 1], t.consMax[3 * k + 2], 0.0f);
            const float pl[3][4] = {{R[0].x, R[0].y, R[0].z, R[0].w}, {R[1].x, R[1].y, R[1].z, R[1].w}, {R[2].x, R[2].y, R[2].z, R[2].w}};
            int desc[16];
            for (int ch = 0; ch < 8; ch++) {
                const size_t cn = f + (size_t)ch;
                const int h3[3] = {ch >> 2, (ch >> 1) & 1, ch & 1};
                for (int a = 0; a < 3; a++)
                    if (t.lo[3 * cn + a] != pl[a][2 * h3[a]] || t.hi[3 * cn + a] != pl[a][2 * h3[a] + 1]) planesShared = false;
                if (t.w1[cn] >= 0) { desc[2 * ch] = t.w0[cn] + (int)refBase; desc[2 * ch + 1] = t.w1[cn]; }
                else { desc[2 * ch] = (-1 - t.w1[cn]) + (int)consBase; desc[2 * ch + 1] = -1; }
            }
            // the node's own Min must be child 0's Min (Min + size*0): the conservative-box test relies on it
            for (int a = 0; a < 3; a++) if (t.lo[3 * n + a] != pl[a][0]) planesShared = false;
            std::memcpy(&R[4], desc, 64);
        }
        std::copy(t.refs.begin(), t.refs.end(), refs.begin() + (long)refBase);
        nodeBase += t.nodeCount(); consBase += t.consMax.size() / 3; refBase += t.refs.size();
    };
    for (size_t i = 0; i < c->meshes.size(); i++) {
        const HostMesh& m = c->meshes[i];
        DevMesh& dm = dmeshes[i];
        std::memcpy(dm.box, m.box.mn, 12);
        std::memcpy(dm.box + 3, m.box.mx, 12);
        dm.rootNode = (int)nodeBase;
        {
            float ov = 0.0f;
            for (size_t n = 0; n < m.tree.nodeCount(); n++) {
                if (m.tree.w1[n] >= 0) continue;
                const size_t k = (size_t)(-1 - m.tree.w1[n]);
                for (int a = 0; a < 3; a++) ov = std::max(ov, m.tree.consMax[3 * k + a] - m.tree.hi[3 * n + a]);
            }
            dm.overshoot = ov;
        }
        std::memcpy(dm.rootBox, &m.tree.lo[0], 12);
        std::memcpy(dm.rootBox + 3, &m.tree.hi[0], 12);
        if (m.tree.w1[0] >= 0) { dm.rootA = m.tree.w0[0] + (int)refBase; dm.rootB = m.tree.w1[0]; }
        else { dm.rootA = (-1 - m.tree.w1[0]) + (int)consBase; dm.rootB = -1; }
        dm.triBase = (int)triBase;
        dm.reflectiveness = m.reflectiveness;
        dm.oneMinusReflectiveness = 1.0f - m.reflectiveness;
        dm.refractionIndex = m.refractionIndex;
        dm.flags = m.flags;
        dm.texOffset = (int)texBase; dm.texW = m.texW; dm.texH = m.texH;
        std::memset(dm._pad, 0, sizeof dm._pad);
        const size_t firstRef = refBase;
        emitTree(m.tree);
        // one intersection record per leaf reference, at the reference's own position
        {
            const long long nr = (long long)m.tree.refs.size();
            const float* src = m.triI.data();
#pragma omp parallel for schedule(static) if (nr > 65536)
            for (long long k = 0; k < nr; k++) {
                const int ti = m.tree.refs[(size_t)k];
                float4* dst = &leafTri[4 * (firstRef + (size_t)k)];
                std::memcpy(dst, src + 16 * (size_t)ti, 64);
                dst[0].w = __int_as_float_host(ti);
            }
        }
        std::memcpy(&triS[5 * triBase], m.triS.data(), m.triS.size() * sizeof(float));
        std::copy(m.texels.begin(), m.texels.end(), texels.begin() + (long)texBase);
        triBase += (size_t)m.nTri; texBase += m.texels.size();
    }
    const int sceneRoot = (int)nodeBase;
    emitTree(c->sceneTree);
    if (!planesShared) return fail(c, XNART_ERR_UNSUPPORTED, "internal: octree children do not share their split planes");
    size_t om = 0;
    for (size_t i = 0; i < c->objs.size(); i++) {
        const HostObj& o = c->objs[i];
        DevObj& d = dobjs[i];
        std::memcpy(d.world, o.world, 64);
        std::memcpy(d.inv, o.inv, 64);
        d.meshStart = (int)om; d.meshCount = (int)o.meshIds.size(); d._pad = 0;
        d.xformClass = (int)i;
        for (size_t j = 0; j < i; j++)
            if (std::memcmp(c->objs[j].inv, o.inv, 64) == 0) { d.xformClass = dobjs[j].xformClass; break; }
        for (int id : o.meshIds) objMeshIds[om++] = id;
    }
    // (object, mesh) pairs in visiting order when the scene octree is one leaf (wavefront path)
    std::vector<int2> pairs;
    if (c->sceneTree.nodeCount() == 1 && c->sceneTree.w1[0] >= 0) {
        for (int i = 0; i < c->sceneTree.w1[0]; i++) {
            const int oi = c->sceneTree.refs[(size_t)c->sceneTree.w0[0] + (size_t)i];
            for (int id : c->objs[(size_t)oi].meshIds) pairs.push_back(make_int2(oi, id));
        }
    }
    c->nPairs = (int)pairs.size();
    c->totalRefs = (int64_t)nMeshRefs;
    c->hostObjs = dobjs;
    // Small opaque scene: pairs + leaf records as one kernel parameter (device_types.h: SmallScene, path.cu)
    c->smallOk = false;
    if (!pairs.empty() && (int)pairs.size() <= kSmallPairs && c->maxMeshDepth == 0 && !c->hasTransparent) {
        SmallScene& SS = c->small;
        std::memset(&SS, 0, sizeof SS);
        std::vector<int> firstOf(c->meshes.size(), -1);
        int used = 0;
        bool fits = true;
        for (size_t p = 0; p < pairs.size() && fits; p++) {
            const int oi = pairs[p].x, mi = pairs[p].y;
            const DevMesh& dm = dmeshes[(size_t)mi];
            const DevObj& dob = dobjs[(size_t)oi];
            SmallPair& P = SS.pair[p];
            if (dm.rootB < 0) { fits = false; break; }   // (cannot happen with maxMeshDepth == 0)
            if (firstOf[(size_t)mi] < 0) {
                if (used + dm.rootB > kSmallTris) { fits = false; break; }
                firstOf[(size_t)mi] = used;
                std::memcpy(&SS.tri[4 * used], &leafTri[4 * (size_t)dm.rootA], (size_t)dm.rootB * 64);
                used += dm.rootB;
            }
            for (int r = 0; r < 4; r++)
                for (int col = 0; col < 3; col++) { P.inv[3 * r + col] = dob.inv[4 * r + col]; P.world[3 * r + col] = dob.world[4 * r + col]; }
            std::memcpy(P.box, dm.box, 24);
            std::memcpy(P.rootBox, dm.rootBox, 24);
            P.obj = oi; P.mesh = mi; P.cls = dob.xformClass;
            P.first = firstOf[(size_t)mi]; P.count = dm.rootB; P.refBase = dm.rootA;
            P.identity = 1;
            for (int k = 0; k < P.count; k++) {
                int id; std::memcpy(&id, &SS.tri[4 * (P.first + k)].w, sizeof id);
                if (id != k) P.identity = 0;
            }
            P.boxCoversRoot = 1;
            for (int a = 0; a < 3; a++)
                if (!(dm.box[a] <= dm.rootBox[a] && dm.box[3 + a] >= dm.rootBox[3 + a])) P.boxCoversRoot = 0;
        }
        if (fits) {
            SS.nPairs = (int)pairs.size();
            const size_t sr = (size_t)sceneRoot;
            SS.sceneBox[0] = nodes[2 * sr].x; SS.sceneBox[1] = nodes[2 * sr].y; SS.sceneBox[2] = nodes[2 * sr].z;
            SS.sceneBox[3] = nodes[2 * sr + 1].x; SS.sceneBox[4] = nodes[2 * sr + 1].y; SS.sceneBox[5] = nodes[2 * sr + 1].z;
            c->smallOk = true;
        }
    }
    // kept on the host for xnart_debug_inner_dump (expansion records and per-mesh bases)
    c->hostInner.assign(reinterpret_cast<const float*>(inner.data()), reinterpret_cast<const float*>(inner.data()) + 4 * inner.size());
    c->meshInnerBase.clear(); c->meshRefBase.clear();
    {
        size_t cb = 0, rb = 0;
        for (const HostMesh& m : c->meshes) {
            c->meshInnerBase.push_back((int)cb); c->meshRefBase.push_back((int)rb);
            cb += m.tree.consMax.size() / 3; rb += m.tree.refs.size();
        }
    }
    if (c->devs.empty()) { c->built = true; return XNART_OK; }  // host-only context: trees and records can be exported
    for (DeviceState& d : c->devs) {
        CU_TRY(c, cudaSetDevice(d.dev));
        CU_TRY(c, cudaStreamSynchronize(d.stream));
        freeScene(d);
        if (!pairs.empty()) CU_TRY(c, uploadVec(&d.pairs, pairs));
        d.scene.pairs = d.pairs; d.scene.nPairs = (int)pairs.size();
        CU_TRY(c, uploadVec(&d.nodes, nodes));
        CU_TRY(c, uploadVec(&d.consMax, cons));
        CU_TRY(c, uploadVec(&d.inner, inner));
        CU_TRY(c, uploadVec(&d.refs, refs));
        CU_TRY(c, uploadVec(&d.leafTri, leafTri));
        CU_TRY(c, uploadVec(&d.triS, triS));
        CU_TRY(c, uploadVec(&d.meshes, dmeshes));
        CU_TRY(c, uploadVec(&d.objs, dobjs));
        CU_TRY(c, uploadVec(&d.objMeshIds, objMeshIds));
        CU_TRY(c, uploadVec(&d.texels, texels));
        d.scene.nodes = d.nodes; d.scene.consMax = d.consMax; d.scene.inner = d.inner; d.scene.refs = d.refs;
        d.scene.leafTri = d.leafTri; d.scene.triS = d.triS; d.scene.meshes = d.meshes; d.scene.objs = d.objs;
        d.scene.objMeshIds = d.objMeshIds; d.scene.texels = d.texels;
        d.scene.sceneRoot = sceneRoot; d.scene.nMeshes = (int)dmeshes.size(); d.scene.nObjs = (int)dobjs.size();
    }
    c->built = true;
    return XNART_OK;
}

int32_t xnart_set_lights(XnartCtx* c, const XnartLight* lights, int32_t n) {
    if (!c || n < 0 || (n > 0 && !lights)) return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_set_lights: bad arguments");
    if (n > kMaxLights) return fail(c, XNART_ERR_UNSUPPORTED, "too many lights");
    for (int i = 0; i < n; i++)
        if (lights[i].kind != XNART_LIGHT_SPOT && lights[i].kind != XNART_LIGHT_DIRECTIONAL)
            return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_set_lights: unknown light kind");
    std::unique_lock<std::mutex> lk(c->renderMutex, std::try_to_lock);
    if (!lk.owns_lock()) return fail(nullptr, XNART_ERR_BUSY, "a render is already in flight");   // RayTracer.cs:62-63
    c->lights.assign(lights, lights + n);
    c->lightsDirty = true;
    return XNART_OK;
}

static int uploadLights(XnartCtx* c) {
    std::vector<DevLight> dl(std::max<size_t>(c->lights.size(), 1));
    for (size_t i = 0; i < c->lights.size(); i++) {
        const XnartLight& s = c->lights[i];
        DevLight& L = dl[i];
        L.kind = s.kind;
        L.px = s.position[0]; L.py = s.position[1]; L.pz = s.position[2];
        L.dx = s.direction[0]; L.dy = s.direction[1]; L.dz = s.direction[2];
        L.cr = s.color[0]; L.cg = s.color[1]; L.cb = s.color[2];
        L.intensity = s.intensity;
        L.angleCosine = (float)std::cos((double)(s.spotAngle * 0.5f));                       // SpotLight.cs:25
        L.powDen = std::pow((double)(1.0f - L.angleCosine), (double)s.decayExponent);       // SpotLight.cs:54
    }
    for (DeviceState& d : c->devs) {
        CU_TRY(c, cudaSetDevice(d.dev));
        CU_TRY(c, cudaStreamSynchronize(d.stream));
        cudaFree(d.lights);
        d.lights = nullptr;
        CU_TRY(c, cudaMalloc(&d.lights, dl.size() * sizeof(DevLight)));
        CU_TRY(c, cudaMemcpy(d.lights, dl.data(), dl.size() * sizeof(DevLight), cudaMemcpyHostToDevice));
    }
    c->lightsDirty = false;
    return XNART_OK;
}

// rows y of [0,H) owned by a shard
static int localRowsOf(int H, int shardIndex, int shardCount, int bandRows) {
    if (shardCount <= 1) return H;
    int rows = 0;
    for (int b = shardIndex; b * bandRows < H; b += shardCount) rows += std::min(bandRows, H - b * bandRows);
    return rows;
}

static int ensureArena(XnartCtx* c, DeviceState& d, unsigned nodes, unsigned rays, bool wave = false, int nLights = 0) {
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const unsigned splitMax = (wave && c->splitRays > 0 && c->nPairs == 1 && !c->countKernels) ? (unsigned)c->splitRays : 0u;
    const size_t splitBytes = al((size_t)splitMax * kSplit * 32);
    const size_t shRays = wave ? (size_t)rays * (size_t)std::max(nLights, 1) : 0;
    size_t need = al((size_t)nodes * 16) * 2 + al((size_t)nodes * 8) + al((size_t)nodes * 4) * 2 +
                  2 * (al((size_t)rays * 16) * 2 + al((size_t)rays * 8));
    if (wave) need += 2 * (al((size_t)rays * 32) + al(shRays * 16) * 2 + al(shRays * 8) + al(shRays * 32)) + 4 * splitBytes;
    if (need > d.arenaBytes) {
        CU_TRY(c, cudaStreamSynchronize(d.stream));
        cudaFree(d.arenaBlock);
        d.arenaBlock = nullptr; d.arenaBytes = 0;
        CU_TRY(c, cudaMalloc(&d.arenaBlock, need));
        d.arenaBytes = need;
    }
    char* p = (char*)d.arenaBlock;
    auto take = [&](size_t bytes) { char* r = p; p += al(bytes); return r; };
    DevArena& A = d.arena;
    A.nodeA = (float4*)take((size_t)nodes * 16);
    A.nodeB = (float4*)take((size_t)nodes * 16);
    A.nodeC = (uint2*)take((size_t)nodes * 8);
    A.color = (uint32_t*)take((size_t)nodes * 4);
    A.flags = (uint32_t*)take((size_t)nodes * 4);
    for (int q = 0; q < 2; q++) {
        A.rayO[q] = (float4*)take((size_t)rays * 16);
        A.rayD[q] = (float4*)take((size_t)rays * 16);
        A.rayIgn[q] = (int2*)take((size_t)rays * 8);
    }
    for (int q = 0; q < 2; q++) {
        A.hit[q] = nullptr; A.shO[q] = A.shD[q] = nullptr; A.shIgn[q] = nullptr; A.shHit[q] = nullptr;
        if (wave) {
            A.hit[q] = (int4*)take((size_t)rays * 32);
            A.shO[q] = (float4*)take(shRays * 16);
            A.shD[q] = (float4*)take(shRays * 16);
            A.shIgn[q] = (int2*)take(shRays * 8);
            A.shHit[q] = (int4*)take(shRays * 32);
        }
        A.hitSplit[q] = (wave && splitMax) ? (int4*)take((size_t)splitMax * kSplit * 32) : nullptr;
        A.shSplit[q] = (wave && splitMax) ? (int4*)take((size_t)splitMax * kSplit * 32) : nullptr;
    }
    A.splitMax = splitMax;
    A.splitHitPct = (unsigned)c->splitHitPct;
    A.levelCount = d.levelWords;
    A.levelBase = d.levelWords + (kMaxLevels + 2);
    A.overflow = d.levelWords + 2 * (kMaxLevels + 2);
    A.shadowHits = d.levelWords + 2 * (kMaxLevels + 2) + 2;
    A.fetch = d.levelWords + 3 * (kMaxLevels + 2) + 2;
    A.counters = d.counters;
    A.nodeCapacity = nodes; A.rayCapacity = rays;
    d.arenaNodes = nodes; d.arenaRays = rays;
    return XNART_OK;
}

// Deep octrees: the per-thread kernel is launched in its high-occupancy build.
static bool deepScene(const XnartCtx* c) { return c->maxMeshDepth >= 2; }
// Wavefront schedule (persistent trace kernel, wavefront.cu): the cost of a query on a deep
// MeshOctree is heavy-tailed and a whole CastRay per thread leaves most lanes of a warp idle.
// Needs the scene octree to be a single leaf; mode 2 forces it wherever it is available.
static bool useWave(const XnartCtx* c, const DeviceState& d) {
    if (c->nPairs <= 0 || d.traceBlocks <= 0 || d.traceBlocksCount <= 0) return false;
    if (c->traversalMode == 1) return false;
    if (c->traversalMode == 2) return true;
    // Any octree of depth >= 2.  Even the 960-triangle spheres of the reference's default scene (Game1.cs:98-107) belong
    // here: their deep recursion levels hold a few thousand rays that start ON a sphere and cross its whole octree, one
    // thread per ray leaves a handful of warps walking them serially (ncu launch list: levels 1-5 take 300-490 us each for
    // 12 K hits), while the persistent kernel serves those leaves with whole warps and splits the queries — 512x512:
    // 1.93 -> 0.81 ms, 1920x1080: 2.48 -> 1.18 ms.
    return deepScene(c);
}

#ifndef XN_PATH_SLICES
#define XN_PATH_SLICES 4   // slices of a host-bound path-kernel frame (each copied while the next ones render)
#endif
// One persistent launch per frame (path.cu) for small opaque scenes.
static bool usePath(const XnartCtx* c, const DeviceState& d) {
    return c->pathKernel && c->smallOk && c->traversalMode == 0 && !c->fuseFrame && d.pathBlocks > 0 && d.pathBlocksCount > 0;
}

static int validateFrame(XnartCtx* c, const XnartFrame* f) {
    if (!f) return fail(c, XNART_ERR_INVALID_ARGUMENT, "null frame");
    if (f->vpW <= 0 || f->vpH <= 0) return fail(c, XNART_ERR_INVALID_ARGUMENT, "empty viewport");
    // Material.LookupUV throws ArgumentException for unknown enum values (Material.cs:85,97)
    if (f->filtering != XNART_FILTER_POINT && f->filtering != XNART_FILTER_BILINEAR)
        return fail(c, XNART_ERR_INVALID_ARGUMENT, "Value does not fall within the expected range: filtering");
    if (f->addressMode < XNART_ADDRESS_CLAMP || f->addressMode > XNART_ADDRESS_MIRROR)
        return fail(c, XNART_ERR_INVALID_ARGUMENT, "Value does not fall within the expected range: addressMode");
    if (f->semantics != XNART_SEMANTICS_XNA40 && f->semantics != XNART_SEMANTICS_MONOGAME)
        return fail(c, XNART_ERR_INVALID_ARGUMENT, "unknown semantics");
    if (f->maxReflections < 0 || f->maxReflections + 1 > kMaxLevels) return fail(c, XNART_ERR_UNSUPPORTED, "maxReflections out of range");
    if (f->shardCount > 1 && (f->shardIndex < 0 || f->shardIndex >= f->shardCount || f->shardBandRows <= 0))
        return fail(c, XNART_ERR_INVALID_ARGUMENT, "bad shard parameters");
    return XNART_OK;
}

// Enqueue the whole trace loop of one frame shard on `st`, writing full-frame-indexed RGBA8 to devOut.
// hostDst != NULL (page-locked, full-frame indexed like devOut): a frame rendered by the path kernel is cut into a few
// slices of whole tile rows, and each slice is copied to the host on the state's copy stream while the next one renders
// (the caller waits for copyStream); *sliced reports whether that happened — otherwise the caller copies the frame itself.
static int enqueueFrame(XnartCtx* c, DeviceState& d, const XnartFrame& f, uint32_t* devOut, cudaStream_t st, bool accumulate = false,
                        const float2* devSamples = nullptr, long long nSamples = 0, uint32_t* hostDst = nullptr, bool* sliced = nullptr) {
    if (sliced) *sliced = false;
    const bool timeIt = !accumulate;          // a single render: per-stage events
    accumulate = accumulate || c->keepCounters;
    DevFrame F;
    {
        // Viewport.Unproject: Invert(Multiply(Multiply(world = Identity, view), proj)) — once per frame
        float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}, wv[16], wvp[16];
        multiply44(ident, f.view, wv);
        multiply44(wv, f.proj, wvp);
        invert44(wvp, F.inv);
    }
    F.vpX = f.vpX; F.vpY = f.vpY; F.vpW = f.vpW; F.vpH = f.vpH;
    F.minDepth = f.minDepth; F.maxDepth = f.maxDepth;
    {   // Viewport.Unproject's source.Z for the near and far points (shade.cuh: unproject); volatile keeps the two IEEE
        // fp32 operations as written whatever the host compiler would like to fold
        volatile float lo = f.minDepth, range = f.maxDepth - f.minDepth;
        F.ndcZ[0] = (0.0f - lo) / range;
        F.ndcZ[1] = (1.0f - lo) / range;
    }
    F.maxReflections = f.maxReflections; F.filtering = f.filtering; F.addressMode = f.addressMode; F.semantics = f.semantics;
    F.shardIndex = f.shardCount > 1 ? f.shardIndex : 0;
    F.shardCount = f.shardCount > 1 ? f.shardCount : 1;
    F.shardBandRows = f.shardCount > 1 ? f.shardBandRows : 1;
    F.nLights = (int)c->lights.size();
    F.tiled = 0;
    F.samples = devSamples;
    if (devSamples) { F.shardIndex = 0; F.shardCount = 1; F.shardBandRows = 1; }
    F.progress = c->rendering.load() ? d.progressHost : nullptr;
    F.progressBase = d.progressEnqueued;
    const int D = f.maxReflections;
    const long long localPixels = devSamples ? nSamples : (long long)localRowsOf(f.vpH, F.shardIndex, F.shardCount, F.shardBandRows) * f.vpW;
    d.lastPixels = accumulate ? d.lastPixels + localPixels : localPixels;
    d.lastLights = F.nLights;
    d.progressEnqueued += (unsigned long long)localPixels;
    const bool mono = f.semantics == XNART_SEMANTICS_MONOGAME;   // compatibility mode: the generic level kernel, MonoGame slab form
    // the path kernel keeps a pixel's position in the frame as a 32-bit index (path.cu: primaryRay)
    const bool fitsPath = devSamples ? nSamples <= 0x7fffffffLL : (long long)f.vpW * (long long)f.vpH <= 0x7fffffffLL;
    if (!mono && fitsPath && usePath(c, d)) {
        if (localPixels > 0x7fffffffLL) return fail(c, XNART_ERR_UNSUPPORTED, "too many pixels in one shard");
        // no arena, no chunks: one persistent launch renders the shard (path.cu)
        F.tiled = (!devSamples && (f.vpW % 8 == 0) && (localPixels % (4LL * f.vpW) == 0)) ? 1 : 0;
        d.singleChunk = true; d.lastFused = false; d.lastWave = false; d.lastPath = true;
        if (!accumulate) CU_TRY(c, cudaMemsetAsync(d.counters, 0, sizeof(unsigned long long) * kCounterSlots, st));
        if (timeIt) CU_TRY(c, cudaEventRecord(d.ev[0], st));
        unsigned* fetch = d.levelWords + 3 * (kMaxLevels + 2) + 2;
        // slices of whole tile rows for the copy overlap (unsharded frames only: a slice is then a run of scanlines)
        const long long tileRow = 4LL * f.vpW;
        const long long tileRows = F.tiled ? localPixels / tileRow : 0;
        int nSlices = 1;
        if (hostDst && F.tiled && F.shardCount == 1 && localPixels >= (1 << 19) && tileRows >= 2 * XN_PATH_SLICES) nSlices = XN_PATH_SLICES;
        // Slices alternate between the caller's stream and the state's side stream, each with its own work counter: a
        // slice's grid fills the GPU, so the next slice's blocks move in exactly as the previous slice's last warps
        // finish their chains — the slices still complete in order, but their tails overlap (four slices on one
        // stream, each behind a counter memset: 0.26 ms of rendering for a frame that takes 0.175 ms unsliced).
        CU_TRY(c, cudaMemsetAsync(d.levelWords, 0, sizeof(unsigned) * kLevelWords, st));
        if (nSlices > 1) {
            CU_TRY(c, cudaEventRecord(d.evB[0], st));
            CU_TRY(c, cudaStreamWaitEvent(d.sideStream, d.evB[0], 0));
        }
        for (int sIdx = 0; sIdx < nSlices; sIdx++) {
            const long long r0 = tileRows * sIdx / nSlices, r1 = tileRows * (sIdx + 1) / nSlices;
            cudaStream_t ss = (sIdx & 1) ? d.sideStream : st;
            F.pixelBase = nSlices > 1 ? r0 * tileRow : 0;
            F.chunkPixels = (int)(nSlices > 1 ? (r1 - r0) * tileRow : localPixels);
            CU_TRY(c, launchPath(c->small, d.scene, F, d.lights, devOut, fetch + sIdx, d.counters,
                                 c->countKernels ? d.pathBlocksCount : d.pathBlocks, c->countKernels, ss));
            c->launches++;
            if (nSlices > 1) {
                CU_TRY(c, cudaEventRecord(d.evA[sIdx], ss));
                CU_TRY(c, cudaStreamWaitEvent(d.copyStream, d.evA[sIdx], 0));
                const size_t off = (size_t)(r0 * tileRow), cnt = (size_t)((r1 - r0) * tileRow);
                CU_TRY(c, cudaMemcpyAsync(hostDst + off, devOut + off, cnt * 4, cudaMemcpyDeviceToHost, d.copyStream));
            }
        }
        if (nSlices > 1) {
            const int lastOdd = ((nSlices - 1) & 1) ? nSlices - 1 : nSlices - 2;   // the side stream's last slice
            CU_TRY(c, cudaStreamWaitEvent(st, d.evA[lastOdd], 0));
        }
        if (nSlices > 1 && sliced) *sliced = true;
        if (timeIt) CU_TRY(c, cudaEventRecord(d.ev[1], st));
        CU_TRY(c, cudaEventRecord(d.ev[2], st));
        d.lastStream = st;
        d.countersPending = true;
        return XNART_OK;
    }
    d.lastPath = false;
    // worst-case arena per pixel: a chain of D+1 path nodes, or a full binary tree when a
    // transparent material can fork every hit into reflection + refraction (RayTracer.cs:586-702)
    double nodesPerPixel = c->hasTransparent ? (std::ldexp(1.0, D + 1) - 1.0) : (double)(D + 1);
    double raysPerPixel = c->hasTransparent ? std::ldexp(1.0, D) : 1.0;
    if (c->hasTransparent && c->optimisticArena) {   // see XnartCtx::arenaScale
        nodesPerPixel = std::min(nodesPerPixel, c->arenaScale * (double)(D + 1));
        raysPerPixel = std::min(raysPerPixel, c->arenaScale);
    }
    const bool wave = !mono && useWave(c, d);
    const double bytesPerPixel = nodesPerPixel * 48.0 + raysPerPixel * (80.0 + (wave ? 2.0 * (32.0 + 72.0 * std::max(F.nLights, 1)) : 0.0)) + 64.0;
    long long chunk = (long long)std::floor((double)c->arenaLimit / bytesPerPixel);
    chunk = std::min<long long>(chunk, (long long)(4.0e9 / nodesPerPixel));
    chunk = std::min(chunk, localPixels);
    if (chunk < 1) return fail(c, XNART_ERR_OUT_OF_MEMORY, "arena limit too small for one pixel at this recursion depth");
    if (chunk < localPixels && chunk >= 4096) chunk -= chunk % 4096;
    // 8x4 pixel tiles per warp when the shard is made of whole tiles (and so is every chunk)
    {
        const long long tileRow = 4LL * f.vpW;
        const bool whole = !devSamples && (f.vpW % 8 == 0) && (localPixels % tileRow == 0);
        if (whole && chunk < localPixels && chunk >= tileRow) chunk -= chunk % tileRow;
        F.tiled = (whole && (chunk >= localPixels || chunk % tileRow == 0)) ? 1 : 0;
    }
    const unsigned nodeCap = (unsigned)std::min(4.0e9, std::ceil((double)chunk * nodesPerPixel));
    const unsigned rayCap = (unsigned)std::min(4.0e9, std::ceil((double)chunk * raysPerPixel));
    int rc = ensureArena(c, d, nodeCap, rayCap, wave, F.nLights);
    if (rc != XNART_OK) return rc;
    const int maxBlocks = d.smCount * 8;
    d.singleChunk = chunk >= localPixels;
    // Shallow scenes (every MeshOctree a leaf or one split): launch overhead of the deep, nearly
    // empty levels matters -> one cooperative launch per chunk.  Deep scenes keep one launch per
    // level in the high-occupancy build of the kernel.
    const bool fused = !mono && !wave && c->fuseFrame && !deepScene(c) && d.frameBlocks > 0 && d.frameBlocksCount > 0;
    d.lastFused = fused;
    d.lastWave = wave;

    if (!accumulate) CU_TRY(c, cudaMemsetAsync(d.counters, 0, sizeof(unsigned long long) * kCounterSlots, st));
    if (timeIt) CU_TRY(c, cudaEventRecord(d.ev[0], st));
    for (long long p0 = 0; p0 < localPixels; p0 += chunk) {
        F.pixelBase = p0;
        F.chunkPixels = (int)std::min(chunk, localPixels - p0);
        CU_TRY(c, cudaMemsetAsync(d.levelWords, 0, sizeof(unsigned) * kLevelWords, st));
        if (fused) {
            // whole chunk in one cooperative launch (levels + resolve, grid barriers in between)
            const int blocks = c->countKernels ? d.frameBlocksCount : d.frameBlocks;
            CU_TRY(c, launchFrame(d.scene, F, d.arena, d.lights, devOut, d.stamps, blocks, c->countKernels, st));
            c->launches++;
            continue;
        }
        for (int L = 0; L <= D; L++) {
            double ub = c->hasTransparent ? std::min((double)rayCap, (double)F.chunkPixels * std::ldexp(1.0, L)) : (double)F.chunkPixels;
            if (!wave) {
                // deep levels usually hold a small fraction of the rays: a grid of resident blocks only (the
                // kernels are grid-stride loops) keeps the cost of a nearly empty launch down
                const int cap = (L >= 2 && !c->hasTransparent) ? d.smCount * 3 : maxBlocks;
                CU_TRY(c, launchLevel(d.scene, F, d.arena, d.lights, L, (long long)ub, cap, c->countKernels, mono ? 3 : (deepScene(c) ? 2 : (c->maxMeshDepth >= 1 ? 1 : 0)), st));
                c->launches++;
                continue;
            }
            // wavefront, main stream: [primary rays] -> trace -> shade A (shadow rays + next level's rays);
            // side stream: shadow trace -> shade B, overlapping the next level on the main stream
            const DevArena& A = d.arena;
            const int tb = c->countKernels ? d.traceBlocksCount : d.traceBlocks;
            if (L == 0) { CU_TRY(c, launchGenPrimary(F, A, maxBlocks, st)); c->launches++; }
            const int in = L & 1;
            // hit / shadow buffers of this parity were last used by level L-2: its shade B must be done
            if (L >= 2) CU_TRY(c, cudaStreamWaitEvent(st, d.evB[L - 2], 0));
            TraceJob J;
            J.rayO = A.rayO[in]; J.rayD = A.rayD[in]; J.rayIgn = A.rayIgn[in]; J.out = A.hit[in];
            J.segCount = &A.levelCount[L]; J.fetch = &A.fetch[2 * L]; J.nSeg = 1; J.segStride = 0;
            J.outSplit = A.hitSplit[in]; J.splitMax = A.splitMax; J.splitHitPct = A.splitHitPct;
            J.hintHits = L > 0 ? &A.shadowHits[L - 1] : nullptr; J.hintRays = L > 0 ? &A.levelCount[L - 1] : nullptr;
            CU_TRY(c, launchTrace(d.scene, J, d.counters, tb, c->countKernels, st));
            CU_TRY(c, launchShadeA(d.scene, F, A, d.lights, L, (long long)ub, maxBlocks, st));
            c->launches += 2;
            CU_TRY(c, cudaEventRecord(d.evA[L], st));
            cudaStream_t ss = c->overlapLevels ? d.sideStream : st;
            if (ss != st) CU_TRY(c, cudaStreamWaitEvent(ss, d.evA[L], 0));
            if (F.nLights > 0) {
                TraceJob Sj;
                Sj.rayO = A.shO[in]; Sj.rayD = A.shD[in]; Sj.rayIgn = A.shIgn[in]; Sj.out = A.shHit[in];
                Sj.segCount = &A.shadowHits[L]; Sj.fetch = &A.fetch[2 * L + 1]; Sj.nSeg = (unsigned)F.nLights; Sj.segStride = A.rayCapacity;
                Sj.outSplit = A.shSplit[in]; Sj.splitMax = A.splitMax; Sj.splitHitPct = A.splitHitPct;
                Sj.hintHits = &A.shadowHits[L]; Sj.hintRays = &A.levelCount[L];
                CU_TRY(c, launchTrace(d.scene, Sj, d.counters, tb, c->countKernels, ss));
                c->launches++;
            }
            CU_TRY(c, launchShadeB(d.scene, F, A, d.lights, L, (long long)ub, maxBlocks, ss));
            c->launches++;
            CU_TRY(c, cudaEventRecord(d.evB[L], ss));
            if (L == D) {
                // the resolve needs every shade B (the side stream runs them in order)
                CU_TRY(c, cudaStreamWaitEvent(st, d.evB[L], 0));
                if (L >= 1) CU_TRY(c, cudaStreamWaitEvent(st, d.evB[L - 1], 0));
            }
        }
        if (p0 == 0) CU_TRY(c, cudaEventRecord(d.ev[1], st));
        if (!c->hasTransparent && c->chainResolve) {
            CU_TRY(c, launchResolveChain(F, d.arena, maxBlocks, devOut, st));
            c->launches++;
            continue;
        }
        for (int L = D - 1; L >= 0; L--) {
            double ub = c->hasTransparent ? std::min((double)rayCap, (double)F.chunkPixels * std::ldexp(1.0, L)) : (double)F.chunkPixels;
            CU_TRY(c, launchResolve(F, d.arena, L, (long long)ub, (L >= 2 && !c->hasTransparent) ? d.smCount * 2 : maxBlocks, devOut, st));
            c->launches++;
        }
        if (D == 0) {
            CU_TRY(c, launchResolve(F, d.arena, 0, F.chunkPixels, maxBlocks, devOut, st));
            c->launches++;
        }
    }
    CU_TRY(c, cudaEventRecord(d.ev[2], st));
    d.lastStream = st;
    d.countersPending = true;
    return XNART_OK;
}

// After an arena overflow in a blocking call: four times the room, unless the worst case was already provided.
static bool growArena(XnartCtx* c) {
    if (!c->overflowed || !c->hasTransparent || c->arenaScale >= 1.0e9) return false;
    c->overflowed = false;
    c->arenaScale *= 4.0;
    return true;
}

// Brackets of a blocking render call for xnart_progress.
static void beginProgress(XnartCtx* c, unsigned long long totalPixels) {
    for (DeviceState& d : c->devs) {
        if (d.progressHost) *reinterpret_cast<volatile unsigned long long*>(d.progressHost) = 0ULL;
        d.progressEnqueued = 0;
    }
    c->progressTotal = totalPixels;
    c->progress = 0.0f;
    c->rendering = true;
}
static void endProgress(XnartCtx* c, bool ok) {
    c->rendering = false;
    if (ok) c->progress = 1.0f;
}
struct ProgressScope {
    XnartCtx* c; bool ok = false;
    ProgressScope(XnartCtx* ctx, unsigned long long total) : c(ctx) { beginProgress(c, total); }
    ~ProgressScope() { endProgress(c, ok); }
};

static int collectCounters(XnartCtx* c) {
    XnartCounters t{};
    bool any = false;
    for (DeviceState& d : c->devs) {
        if (!d.countersPending && !c->keepCounters) continue;
        any = true;
        CU_TRY(c, cudaSetDevice(d.dev));
        if (d.lastStream) CU_TRY(c, cudaStreamSynchronize(d.lastStream));
        struct { unsigned long long h[kCounterSlots]; unsigned words[kOvfWord + 2]; } got;
        static_assert(sizeof got == kCounterSnapBytes, "counter snapshot layout");
        if (d.countersPrefetched) std::memcpy(&got, d.countersHost, sizeof got);   // copied behind the frame on the stream just synchronised
        else CU_TRY(c, cudaMemcpy(&got, d.counters, sizeof got, cudaMemcpyDeviceToHost));
        d.countersPrefetched = false;
        const unsigned long long* h = got.h;
        const unsigned ovf = got.words[kOvfWord];
        if (ovf) {
            c->overflowed = true;
            d.countersPending = false;
            return fail(c, XNART_ERR_OUT_OF_MEMORY, "ray queue overflow: the path arena was too small for this frame");
        }
        t.pixels += (uint64_t)d.lastPixels;
        t.raysPrimary += (uint64_t)d.lastPixels;
        t.hits += h[CTR_HITS];
        t.raysShadow += h[CTR_HITS] * (uint64_t)d.lastLights;
        t.raysReflection += h[CTR_RAYS_REFLECTION];
        t.raysRefraction += h[CTR_RAYS_REFRACTION];
        t.nBox += h[CTR_NBOX]; t.nTriRef += h[CTR_NTRIREF]; t.nTriCulled += h[CTR_NTRICULLED]; t.nTriFull += h[CTR_NTRIFULL];
        float ms = 0, ms2 = 0;
        if (d.lastFused && d.singleChunk) {
            unsigned long long st3[3] = {0, 0, 0};
            CU_TRY(c, cudaMemcpy(st3, d.stamps, sizeof st3, cudaMemcpyDeviceToHost));
            t.msLevels = std::max(t.msLevels, (double)(st3[1] - st3[0]) * 1e-6);
            t.msResolve = std::max(t.msResolve, (double)(st3[2] - st3[1]) * 1e-6);
        } else if (d.singleChunk) {
            if (cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]) == cudaSuccess) t.msLevels = std::max(t.msLevels, (double)ms);
            if (cudaEventElapsedTime(&ms2, d.ev[1], d.ev[2]) == cudaSuccess) t.msResolve = std::max(t.msResolve, (double)ms2);
        } else if (cudaEventElapsedTime(&ms, d.ev[0], d.ev[2]) == cudaSuccess) {
            t.msLevels = std::max(t.msLevels, (double)ms);  // chunks interleave level and resolve kernels
        }
        d.countersPending = false;
    }
    if (any) {
        t.msCopy = c->last.msCopy; t.msTotal = c->last.msTotal;
        c->last = t;
    }
    return XNART_OK;
}

static int renderFrameDeviceBody(XnartCtx* c, const XnartFrame* frame, void* devOut, void* cudaStream) {
    if (!c || !devOut) return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_render_frame_device: bad arguments");
    if (c->devs.empty()) return fail(c, XNART_ERR_NO_DEVICE, "host-only context: libxnart has no CPU fallback");
    if (!c->built) return fail(c, XNART_ERR_NOT_BUILT, "xnart_build_scene has not been called");
    int rc = validateFrame(c, frame);
    if (rc != XNART_OK) return rc;
    if (c->lightsDirty && (rc = uploadLights(c)) != XNART_OK) return rc;
    DeviceState& d = c->devs[0];
    CU_TRY(c, cudaSetDevice(d.dev));
    cudaStream_t st = cudaStream ? (cudaStream_t)cudaStream : d.stream;
    // Deep scenes (wavefront path): a frame is ~30 launches, each ending in the serial tail of its longest rays.  With k
    // in-flight states on this GPU the shard is rendered as k sub-shards of scanline bands side by side — sub-shard j of
    // shard (i of n) is shard (i + j*n of k*n), a subset of the same bands — so the tails of one overlap the bulk of the
    // others.  What a rank of a tile split gains with k = 2 is in DESIGN.md section 7.
    std::vector<int> states;
    for (int i = 0; i < (int)c->devs.size(); i++)
        if (c->devs[i].dev == d.dev) states.push_back(i);
    const int sc = frame->shardCount > 1 ? frame->shardCount : 1;
    const int si = sc > 1 ? frame->shardIndex : 0, br = sc > 1 ? frame->shardBandRows : 8;
    const int k = (int)states.size();
    const bool sub = k > 1 && frame->semantics != XNART_SEMANTICS_MONOGAME && useWave(c, d) && (long long)br * sc * k <= (long long)frame->vpH;
    auto enqueueAll = [&]() -> int {
        if (!sub) return enqueueFrame(c, d, *frame, (uint32_t*)devOut, st);
        CU_TRY(c, cudaEventRecord(d.evRendered[0], st));
        for (int j = 0; j < k; j++) {
            DeviceState& dj = c->devs[states[(size_t)j]];
            cudaStream_t sj = (j == 0) ? st : dj.stream;
            if (sj != st) CU_TRY(c, cudaStreamWaitEvent(sj, d.evRendered[0], 0));
            XnartFrame f = *frame;
            f.shardCount = sc * k; f.shardIndex = si + j * sc; f.shardBandRows = br;
            int r = enqueueFrame(c, dj, f, (uint32_t*)devOut, sj);
            if (r != XNART_OK) return r;
            if (sj != st) {
                CU_TRY(c, cudaEventRecord(dj.evRendered[1], sj));
                CU_TRY(c, cudaStreamWaitEvent(st, dj.evRendered[1], 0));
            }
        }
        return XNART_OK;
    };
    if (cudaStream) {   // asynchronous: the caller owns completion, nothing to poll
        c->progress = 0.0f;
        rc = enqueueAll();
        if (rc == XNART_OK) c->progress = 1.0f;
        return rc;
    }
    ProgressScope ps(c, (unsigned long long)localRowsOf(frame->vpH, si, sc, br) * (unsigned long long)frame->vpW);
    rc = enqueueAll();
    if (rc != XNART_OK) return rc;
    if (!sub && d.countersHost && cudaMemcpyAsync(d.countersHost, d.counters, kCounterSnapBytes, cudaMemcpyDeviceToHost, st) == cudaSuccess)
        d.countersPrefetched = true;   // read by collectCounters after the synchronisation below
    CU_TRY(c, cudaStreamSynchronize(st));
    c->last.msCopy = 0; c->last.msTotal = 0;
    rc = collectCounters(c);
    c->last.msTotal = c->last.msLevels + c->last.msResolve;
    ps.ok = rc == XNART_OK;
    return rc;
}

static int renderFrameBody(XnartCtx* c, const XnartFrame* frame, uint32_t* out) {
    if (!c || !out) return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_render_frame: bad arguments");
    if (c->devs.empty()) return fail(c, XNART_ERR_NO_DEVICE, "host-only context: libxnart has no CPU fallback");
    if (!c->built) return fail(c, XNART_ERR_NOT_BUILT, "xnart_build_scene has not been called");
    int rc = validateFrame(c, frame);
    if (rc != XNART_OK) return rc;
    if (c->lightsDirty && (rc = uploadLights(c)) != XNART_OK) return rc;
    const size_t pixels = (size_t)frame->vpW * (size_t)frame->vpH;
    const int psc = frame->shardCount > 1 ? frame->shardCount : 1;
    ProgressScope ps(c, (unsigned long long)localRowsOf(frame->vpH, psc > 1 ? frame->shardIndex : 0, psc, psc > 1 ? frame->shardBandRows : 1) * (unsigned long long)frame->vpW);
    // One frame is split over DISTINCT GPUs only: a device id listed several times gives several in-flight
    // states for xnart_render_frames (whole frames side by side); splitting one frame over them would only
    // repeat the launch sequence on the same GPU.
    std::vector<int> prim;
    for (int i = 0; i < (int)c->devs.size(); i++) {
        bool seen = false;
        for (int j : prim) seen = seen || c->devs[j].dev == c->devs[i].dev;
        if (!seen) prim.push_back(i);
    }
    const int nDev = (int)prim.size();
    // Is the caller's buffer page-locked?  Then the device->host copy can land in it directly.
    cudaPointerAttributes pa{};
    bool outPinned = (cudaPointerGetAttributes(&pa, out) == cudaSuccess) && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    // In-process multi-GPU: scanline bands are dealt round-robin to the devices of the context
    // (RayTracer.cs:49-52 treats scanlines as independent work items); a caller-supplied shard is
    // subdivided the same way.
    const bool callerShard = frame->shardCount > 1;
    if (callerShard && nDev > 1) return fail(c, XNART_ERR_UNSUPPORTED, "caller sharding with a multi-device context");
    for (int i = 0; i < nDev; i++) {
        DeviceState& d = c->devs[prim[i]];
        CU_TRY(c, cudaSetDevice(d.dev));
        if (d.frameBufPixels < pixels) {
            CU_TRY(c, cudaStreamSynchronize(d.stream));
            cudaFree(d.frameBuf); d.frameBuf = nullptr; d.frameBufPixels = 0;
            CU_TRY(c, cudaMalloc(&d.frameBuf, pixels * 4));
            d.frameBufPixels = pixels;
        }
        if (!outPinned && d.pinnedPixels < pixels) {
            if (d.pinned) cudaFreeHost(d.pinned);
            d.pinned = nullptr; d.pinnedPixels = 0;
            CU_TRY(c, cudaMallocHost(&d.pinned, pixels * 4));
            d.pinnedPixels = pixels;
        }
        XnartFrame f = *frame;
        if (nDev > 1) { f.shardIndex = i; f.shardCount = nDev; f.shardBandRows = 8; }
        uint32_t* hostDst = outPinned ? out : d.pinned;
        bool sliced = false;
        // (small scenes: the frame is rendered in a few slices and every slice is copied while the next one renders)
        rc = enqueueFrame(c, d, f, d.frameBuf, d.stream, false, nullptr, 0, hostDst, &sliced);
        if (rc != XNART_OK) return rc;
        if (sliced) {
            CU_TRY(c, cudaEventRecord(d.evCopied[0], d.copyStream));
            CU_TRY(c, cudaStreamWaitEvent(d.stream, d.evCopied[0], 0));   // the caller waits on d.stream below
            if (d.countersHost && cudaMemcpyAsync(d.countersHost, d.counters, kCounterSnapBytes, cudaMemcpyDeviceToHost, d.stream) == cudaSuccess) d.countersPrefetched = true;
            CU_TRY(c, cudaEventRecord(d.ev[3], d.stream));
            continue;
        }
        // device -> host: only the rows this device owns
        const int sc = f.shardCount > 1 ? f.shardCount : 1, br = f.shardCount > 1 ? f.shardBandRows : f.vpH;
        for (int b = (sc > 1 ? f.shardIndex : 0); (long long)b * br < f.vpH; b += sc) {
            int y0 = b * br, rows = std::min(br, f.vpH - y0);
            size_t off = (size_t)y0 * (size_t)f.vpW;
            CU_TRY(c, cudaMemcpyAsync(hostDst + off, d.frameBuf + off, (size_t)rows * f.vpW * 4, cudaMemcpyDeviceToHost, d.stream));
        }
        if (d.countersHost && cudaMemcpyAsync(d.countersHost, d.counters, kCounterSnapBytes, cudaMemcpyDeviceToHost, d.stream) == cudaSuccess) d.countersPrefetched = true;
        CU_TRY(c, cudaEventRecord(d.ev[3], d.stream));
    }
    double msTotal = 0, msCopy = 0;
    for (int i = 0; i < nDev; i++) {
        DeviceState& d = c->devs[prim[i]];
        CU_TRY(c, cudaSetDevice(d.dev));
        CU_TRY(c, cudaStreamSynchronize(d.stream));
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, d.ev[0], d.ev[3]);
        cudaEventElapsedTime(&b, d.ev[2], d.ev[3]);
        msTotal = std::max(msTotal, (double)a); msCopy = std::max(msCopy, (double)b);
        if (!outPinned) {
            const int sc = nDev > 1 ? nDev : (callerShard ? frame->shardCount : 1);
            const int br = nDev > 1 ? 8 : (callerShard ? frame->shardBandRows : frame->vpH);
            const int si = nDev > 1 ? i : (callerShard ? frame->shardIndex : 0);
            for (int bnd = si; (long long)bnd * br < frame->vpH; bnd += sc) {
                int y0 = bnd * br, rows = std::min(br, frame->vpH - y0);
                size_t off = (size_t)y0 * (size_t)frame->vpW;
                std::memcpy(out + off, d.pinned + off, (size_t)rows * frame->vpW * 4);
            }
        }
    }
    c->last.msCopy = msCopy; c->last.msTotal = msTotal;
    rc = collectCounters(c);
    c->last.msCopy = msCopy; c->last.msTotal = msTotal;
    ps.ok = rc == XNART_OK;
    return rc;
}

// A camera flythrough (config C4; Game1.cs:261-266 moves the camera on a circle and renders frame
// after frame): n frames of one viewport size into n consecutive host frames.  Frame f is rendered by
// device f % nDevices; on each device the device->host copy of a frame overlaps the rendering of
// the next one (two frame buffers, a copy stream).  Replaces n x RenderInternal + SetData.
static int renderFramesBody(XnartCtx* c, const XnartFrame* frames, int32_t n, uint32_t* out) {
    if (!c || n < 0 || (n > 0 && (!frames || !out))) return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_render_frames: bad arguments");
    if (c->devs.empty()) return fail(c, XNART_ERR_NO_DEVICE, "host-only context: libxnart has no CPU fallback");
    if (!c->built) return fail(c, XNART_ERR_NOT_BUILT, "xnart_build_scene has not been called");
    if (n == 0) return XNART_OK;
    int rc;
    for (int i = 0; i < n; i++) {
        if ((rc = validateFrame(c, &frames[i])) != XNART_OK) return rc;
        if (frames[i].vpW != frames[0].vpW || frames[i].vpH != frames[0].vpH || frames[i].shardCount > 1)
            return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_render_frames: all frames must share one viewport size and be unsharded");
    }
    if (c->lightsDirty && (rc = uploadLights(c)) != XNART_OK) return rc;
    const size_t pixels = (size_t)frames[0].vpW * (size_t)frames[0].vpH;
    ProgressScope ps(c, (unsigned long long)n * (unsigned long long)pixels);
    const int nDev = (int)c->devs.size();
    cudaPointerAttributes pa{};
    const bool outPinned = (cudaPointerGetAttributes(&pa, out) == cudaSuccess) && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    for (DeviceState& d : c->devs) {
        CU_TRY(c, cudaSetDevice(d.dev));
        CU_TRY(c, cudaStreamSynchronize(d.stream));
        CU_TRY(c, cudaStreamSynchronize(d.copyStream));
        if (d.frameBufPixels < pixels) {
            cudaFree(d.frameBuf); d.frameBuf = nullptr; d.frameBufPixels = 0;
            CU_TRY(c, cudaMalloc(&d.frameBuf, pixels * 4));
            d.frameBufPixels = pixels;
        }
        if (d.frameBuf2Pixels < pixels) {
            cudaFree(d.frameBuf2); d.frameBuf2 = nullptr; d.frameBuf2Pixels = 0;
            CU_TRY(c, cudaMalloc(&d.frameBuf2, pixels * 4));
            d.frameBuf2Pixels = pixels;
        }
        if (!outPinned && d.pinnedPixels < 2 * pixels) {
            if (d.pinned) cudaFreeHost(d.pinned);
            d.pinned = nullptr; d.pinnedPixels = 0;
            CU_TRY(c, cudaMallocHost(&d.pinned, 2 * pixels * 4));
            d.pinnedPixels = 2 * pixels;
        }
        if (!c->keepCounters) {
            d.lastPixels = 0;
            CU_TRY(c, cudaMemsetAsync(d.counters, 0, sizeof(unsigned long long) * kCounterSlots, d.stream));
        }
        CU_TRY(c, cudaEventRecord(d.ev[0], d.stream));
    }
    std::vector<int> perDev(nDev, 0);
    // staged (pageable destination) frames waiting for their host memcpy: (device, slot, frame index)
    struct Pending { int dev, slot, frame; };
    std::vector<Pending> pending;
    auto drain = [&](size_t keep) -> int {
        while (pending.size() > keep) {
            Pending p = pending.front();
            pending.erase(pending.begin());
            DeviceState& d = c->devs[p.dev];
            CU_TRY(c, cudaSetDevice(d.dev));
            CU_TRY(c, cudaEventSynchronize(d.evCopied[p.slot]));
            std::memcpy(out + (size_t)p.frame * pixels, d.pinned + (size_t)p.slot * pixels, pixels * 4);
        }
        return XNART_OK;
    };
    for (int f = 0; f < n; f++) {
        const int di = f % nDev;
        DeviceState& d = c->devs[di];
        CU_TRY(c, cudaSetDevice(d.dev));
        const int slot = perDev[di] & 1;
        uint32_t* buf = slot ? d.frameBuf2 : d.frameBuf;
        if (perDev[di] >= 2) {
            if (!outPinned && (rc = drain((size_t)nDev)) != XNART_OK) return rc;  // the staging slot must be free again
            CU_TRY(c, cudaSetDevice(d.dev));                                       // (drain visits other devices)
            CU_TRY(c, cudaStreamWaitEvent(d.stream, d.evCopied[slot], 0));        // buffer free once its copy is done
        }
        rc = enqueueFrame(c, d, frames[f], buf, d.stream, /*accumulate=*/true);
        if (rc != XNART_OK) return rc;
        CU_TRY(c, cudaEventRecord(d.evRendered[slot], d.stream));
        CU_TRY(c, cudaStreamWaitEvent(d.copyStream, d.evRendered[slot], 0));
        uint32_t* dst = outPinned ? out + (size_t)f * pixels : d.pinned + (size_t)slot * pixels;
        CU_TRY(c, cudaMemcpyAsync(dst, buf, pixels * 4, cudaMemcpyDeviceToHost, d.copyStream));
        CU_TRY(c, cudaEventRecord(d.evCopied[slot], d.copyStream));
        if (!outPinned) pending.push_back(Pending{di, slot, f});
        perDev[di]++;
    }
    if (!outPinned && (rc = drain(0)) != XNART_OK) return rc;
    double msTotal = 0;
    for (DeviceState& d : c->devs) {
        CU_TRY(c, cudaSetDevice(d.dev));
        CU_TRY(c, cudaEventRecord(d.ev[3], d.copyStream));
        CU_TRY(c, cudaStreamSynchronize(d.stream));
        CU_TRY(c, cudaStreamSynchronize(d.copyStream));
        float a = 0;
        if (cudaEventElapsedTime(&a, d.ev[0], d.ev[3]) == cudaSuccess) msTotal = std::max(msTotal, (double)a);
        d.singleChunk = false;  // per-stage split is not meaningful for a pipelined batch
        d.lastFused = false;
    }
    c->last.msCopy = 0; c->last.msTotal = msTotal;
    rc = collectCounters(c);
    c->last.msCopy = 0; c->last.msTotal = msTotal;
    ps.ok = rc == XNART_OK;
    return rc;
}

// xnart_render_frames with the output left on the device: n frames of one viewport size into n consecutive
// device frames, dealt round-robin over the in-flight states of the context's first GPU (a device id listed k
// times in xnart_create keeps k frames in flight).  The work is ordered after what is already enqueued on
// `cudaStream` and the stream is made to wait for all of it, so the caller can chain a collective or a copy;
// with a NULL stream the call returns after completion.
static int renderFramesDeviceBody(XnartCtx* c, const XnartFrame* frames, int32_t n, void* devOut, void* cudaStream) {
    if (!c || n < 0 || (n > 0 && (!frames || !devOut))) return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_render_frames_device: bad arguments");
    if (c->devs.empty()) return fail(c, XNART_ERR_NO_DEVICE, "host-only context: libxnart has no CPU fallback");
    if (!c->built) return fail(c, XNART_ERR_NOT_BUILT, "xnart_build_scene has not been called");
    if (n == 0) return XNART_OK;
    int rc;
    for (int i = 0; i < n; i++) {
        if ((rc = validateFrame(c, &frames[i])) != XNART_OK) return rc;
        if (frames[i].vpW != frames[0].vpW || frames[i].vpH != frames[0].vpH || frames[i].shardCount > 1)
            return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_render_frames_device: all frames must share one viewport size and be unsharded");
    }
    if (c->lightsDirty && (rc = uploadLights(c)) != XNART_OK) return rc;
    const size_t pixels = (size_t)frames[0].vpW * (size_t)frames[0].vpH;
    // (a NULL stream makes the call blocking: progress can be polled; with a caller's stream nothing is reported)
    std::unique_ptr<ProgressScope> ps;
    if (!cudaStream) ps.reset(new ProgressScope(c, (unsigned long long)n * (unsigned long long)pixels));
    else c->progress = 0.0f;
    std::vector<int> states;   // in-flight states on the GPU of the first entry
    for (int i = 0; i < (int)c->devs.size(); i++)
        if (c->devs[i].dev == c->devs[0].dev) states.push_back(i);
    DeviceState& d0 = c->devs[0];
    CU_TRY(c, cudaSetDevice(d0.dev));
    cudaStream_t user = cudaStream ? (cudaStream_t)cudaStream : d0.stream;
    // order the states' streams after the caller's stream (buffer reuse), then fan out
    CU_TRY(c, cudaEventRecord(d0.evRendered[0], user));
    for (int si : states) {
        DeviceState& d = c->devs[si];
        if (d.stream != user) CU_TRY(c, cudaStreamWaitEvent(d.stream, d0.evRendered[0], 0));
        if (!c->keepCounters) {
            d.lastPixels = 0;
            CU_TRY(c, cudaMemsetAsync(d.counters, 0, sizeof(unsigned long long) * kCounterSlots, d.stream));
        }
        CU_TRY(c, cudaEventRecord(d.ev[0], d.stream));
    }
    for (int f = 0; f < n; f++) {
        DeviceState& d = c->devs[states[(size_t)f % states.size()]];
        rc = enqueueFrame(c, d, frames[f], (uint32_t*)devOut + (size_t)f * pixels, d.stream, /*accumulate=*/true);
        if (rc != XNART_OK) return rc;
    }
    for (int si : states) {
        DeviceState& d = c->devs[si];
        d.singleChunk = false; d.lastFused = false;
        if (d.stream == user) continue;
        CU_TRY(c, cudaEventRecord(d.evRendered[1], d.stream));
        CU_TRY(c, cudaStreamWaitEvent(user, d.evRendered[1], 0));
    }
    if (!cudaStream) {
        CU_TRY(c, cudaStreamSynchronize(user));
        c->last.msCopy = 0; c->last.msTotal = 0;
        rc = collectCounters(c);
        c->last.msTotal = c->last.msLevels;
        ps->ok = rc == XNART_OK;
    } else {
        c->progress = 1.0f;
    }
    return rc;
}

// Scratch of the batch entry points: one block per in-flight state that only grows (no cudaMalloc / cudaFree per call).
static cudaError_t ensureScratch(DeviceState& d, size_t bytes) {
    if (bytes <= d.scratchBytes) return cudaSuccess;
    cudaError_t e = cudaStreamSynchronize(d.stream);
    if (e != cudaSuccess) return e;
    cudaFree(d.scratch);
    d.scratch = nullptr; d.scratchBytes = 0;
    bytes += bytes / 4;   // head room: batches of an adaptive-multisampling pass vary in size
    e = cudaMalloc(&d.scratch, bytes);
    if (e == cudaSuccess) d.scratchBytes = bytes;
    return e;
}

// CastRay for arbitrary screen positions: what GetColorForQuadrant does four times per quadrant
// (RayTracer.cs:220-276) — Unproject (x, y, 0) and (x, y, 1), normalise, CastRay at iteration 0.
// screenXY = n x (x, y) in viewport pixel units (sub-pixel positions allowed), rgba8Out = n colours.
// The adaptive refinement around it (RayTracer.cs:279-309) is host logic (host.py).
static int castRaysBody(XnartCtx* c, const XnartFrame* frame, const float* screenXY, int64_t n, uint32_t* out) {
    if (!c || n < 0 || (n > 0 && (!screenXY || !out))) return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_cast_rays: bad arguments");
    if (c->devs.empty()) return fail(c, XNART_ERR_NO_DEVICE, "host-only context: libxnart has no CPU fallback");
    if (!c->built) return fail(c, XNART_ERR_NOT_BUILT, "xnart_build_scene has not been called");
    int rc = validateFrame(c, frame);
    if (rc != XNART_OK) return rc;
    if (n > 0x7fffffffLL) return fail(c, XNART_ERR_UNSUPPORTED, "xnart_cast_rays: too many samples in one batch");
    if (n == 0) return XNART_OK;
    if (c->lightsDirty && (rc = uploadLights(c)) != XNART_OK) return rc;
    DeviceState& d = c->devs[0];
    CU_TRY(c, cudaSetDevice(d.dev));
    const size_t xyBytes = (((size_t)n * 8) + 255) & ~(size_t)255;
    cudaError_t e = ensureScratch(d, xyBytes + (size_t)n * 4);
    float2* ds = (float2*)d.scratch;
    uint32_t* dout = (uint32_t*)((char*)d.scratch + xyBytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ds, screenXY, (size_t)n * 8, cudaMemcpyHostToDevice, d.stream);
    if (e != cudaSuccess) return fail(c, e == cudaErrorMemoryAllocation ? XNART_ERR_OUT_OF_MEMORY : XNART_ERR_CUDA, std::string("xnart_cast_rays: ") + cudaGetErrorString(e));
    rc = enqueueFrame(c, d, *frame, dout, d.stream, false, ds, n);
    if (rc == XNART_OK) {
        e = cudaMemcpyAsync(out, dout, (size_t)n * 4, cudaMemcpyDeviceToHost, d.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
        if (e != cudaSuccess) rc = fail(c, XNART_ERR_CUDA, std::string("xnart_cast_rays: ") + cudaGetErrorString(e));
    } else {
        cudaStreamSynchronize(d.stream);
    }
    if (rc != XNART_OK) return rc;
    c->last.msCopy = 0; c->last.msTotal = 0;
    rc = collectCounters(c);
    c->last.msTotal = c->last.msLevels + c->last.msResolve;
    return rc;
}

// ---- public render entry points: one render in flight per context (RayTracer.cs:62-63), and — for calls that return
// after completion — the arena-overflow retry described at XnartCtx::arenaScale.
#define XNART_RENDER_ENTRY(blocking, call)                                                                          \
    if (!c) return fail(nullptr, XNART_ERR_INVALID_ARGUMENT, "null context");                                       \
    std::unique_lock<std::mutex> lk(c->renderMutex, std::try_to_lock);                                              \
    if (!lk.owns_lock()) return fail(nullptr, XNART_ERR_BUSY, "a render is already in flight");                     \
    int rc;                                                                                                         \
    do {                                                                                                            \
        c->optimisticArena = (blocking);                                                                            \
        c->overflowed = false;                                                                                      \
        rc = (call);                                                                                                \
        c->optimisticArena = false;                                                                                 \
    } while ((blocking) && rc == XNART_ERR_OUT_OF_MEMORY && growArena(c));                                          \
    return rc;

int32_t xnart_render_frame_device(XnartCtx* c, const XnartFrame* frame, void* devOut, void* cudaStream) {
    XNART_RENDER_ENTRY(cudaStream == nullptr, renderFrameDeviceBody(c, frame, devOut, cudaStream))
}
int32_t xnart_render_frame(XnartCtx* c, const XnartFrame* frame, uint32_t* out) {
    XNART_RENDER_ENTRY(true, renderFrameBody(c, frame, out))
}
int32_t xnart_render_frames(XnartCtx* c, const XnartFrame* frames, int32_t n, uint32_t* out) {
    XNART_RENDER_ENTRY(true, renderFramesBody(c, frames, n, out))
}
int32_t xnart_render_frames_device(XnartCtx* c, const XnartFrame* frames, int32_t n, void* devOut, void* cudaStream) {
    XNART_RENDER_ENTRY(cudaStream == nullptr, renderFramesDeviceBody(c, frames, n, devOut, cudaStream))
}
int32_t xnart_cast_rays(XnartCtx* c, const XnartFrame* frame, const float* screenXY, int64_t n, uint32_t* out) {
    XNART_RENDER_ENTRY(true, castRaysBody(c, frame, screenXY, n, out))
}
#undef XNART_RENDER_ENTRY

int32_t xnart_trace_rays(XnartCtx* c, const float* rays6, int64_t n, const int32_t* ignore2, int32_t* objId, int32_t* meshId,
                         int32_t* triId, float* u, float* v, float* dOut, float* worldPos3) {
    if (!c || n < 0 || (n > 0 && (!rays6 || !objId || !meshId || !triId || !u || !v || !dOut)))
        return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_trace_rays: bad arguments");
    std::unique_lock<std::mutex> lk(c->renderMutex, std::try_to_lock);
    if (!lk.owns_lock()) return fail(nullptr, XNART_ERR_BUSY, "a render is already in flight");
    if (c->devs.empty()) return fail(c, XNART_ERR_NO_DEVICE, "host-only context: libxnart has no CPU fallback");
    if (!c->built) return fail(c, XNART_ERR_NOT_BUILT, "xnart_build_scene has not been called");
    if (n == 0) return XNART_OK;
    DeviceState& d = c->devs[0];
    CU_TRY(c, cudaSetDevice(d.dev));
    size_t N = (size_t)n;
    if (N > 0x7fffffffULL) return fail(c, XNART_ERR_UNSUPPORTED, "xnart_trace_rays: too many rays in one batch");
    const bool monoT = c->traceSemantics == XNART_SEMANTICS_MONOGAME;
    const bool wave = !monoT && useWave(c, d);
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t need = al(N * 24) + 6 * al(N * 4) + al(N * 12) + al(N * 8);
    if (wave) need += 2 * al(N * 16) + al(N * 8) + al(N * 32) + 256;
    {
        cudaError_t e = ensureScratch(d, need);
        if (e != cudaSuccess) return fail(c, e == cudaErrorMemoryAllocation ? XNART_ERR_OUT_OF_MEMORY : XNART_ERR_CUDA, std::string("xnart_trace_rays: ") + cudaGetErrorString(e));
    }
    char* sp = (char*)d.scratch;
    auto take = [&](size_t bytes) { char* r = sp; sp += al(bytes); return r; };
    float* dr = (float*)take(N * 24);
    float *du = (float*)take(N * 4), *dv = (float*)take(N * 4), *dd = (float*)take(N * 4), *dp = (float*)take(N * 12);
    int *dobj = (int*)take(N * 4), *dmesh = (int*)take(N * 4), *dtri = (int*)take(N * 4);
    int* di = ignore2 ? (int*)take(N * 8) : nullptr;
    float4 *qo = nullptr, *qd = nullptr; int2* qi = nullptr; int4* qh = nullptr; unsigned* qw = nullptr;
    if (wave) { qo = (float4*)take(N * 16); qd = (float4*)take(N * 16); qi = (int2*)take(N * 8); qh = (int4*)take(N * 32); qw = (unsigned*)take(8); }
#define TR_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cudaStreamSynchronize(d.stream); return fail(c, XNART_ERR_CUDA, std::string(#expr " failed: ") + cudaGetErrorString(_e)); } } while (0)
    if (ignore2) TR_TRY(cudaMemcpyAsync(di, ignore2, N * 8, cudaMemcpyHostToDevice, d.stream));
    TR_TRY(cudaMemcpyAsync(dr, rays6, N * 24, cudaMemcpyHostToDevice, d.stream));
    TR_TRY(cudaMemsetAsync(d.counters, 0, sizeof(unsigned long long) * kCounterSlots, d.stream));
    if (wave) {
        // same persistent kernel as the frame path: pack -> trace -> unpack
        const unsigned words[2] = {(unsigned)N, 0u};
        TR_TRY(cudaMemcpyAsync(qw, words, 8, cudaMemcpyHostToDevice, d.stream));
        TR_TRY(launchPackRays(dr, di, n, qo, qd, qi, d.smCount * 8, d.stream));
        TraceJob J;
        J.rayO = qo; J.rayD = qd; J.rayIgn = qi; J.out = qh; J.segCount = qw; J.fetch = qw + 1; J.nSeg = 1; J.segStride = 0;
        J.outSplit = nullptr; J.splitMax = 0; J.hintHits = J.hintRays = nullptr; J.splitHitPct = 100;
        TR_TRY(launchTrace(d.scene, J, d.counters, c->countKernels ? d.traceBlocksCount : d.traceBlocks, c->countKernels, d.stream));
        TR_TRY(launchUnpackHits(d.scene, qh, n, dobj, dmesh, dtri, du, dv, dd, dp, d.counters, d.smCount * 8, d.stream));
        c->launches += 3;
    } else {
        TR_TRY(launchTraceRays(d.scene, dr, n, di, dobj, dmesh, dtri, du, dv, dd, dp, d.counters, d.smCount * 8, c->countKernels, monoT, d.stream));
        c->launches++;
    }
    TR_TRY(cudaMemcpyAsync(objId, dobj, N * 4, cudaMemcpyDeviceToHost, d.stream));
    TR_TRY(cudaMemcpyAsync(meshId, dmesh, N * 4, cudaMemcpyDeviceToHost, d.stream));
    TR_TRY(cudaMemcpyAsync(triId, dtri, N * 4, cudaMemcpyDeviceToHost, d.stream));
    TR_TRY(cudaMemcpyAsync(u, du, N * 4, cudaMemcpyDeviceToHost, d.stream));
    TR_TRY(cudaMemcpyAsync(v, dv, N * 4, cudaMemcpyDeviceToHost, d.stream));
    TR_TRY(cudaMemcpyAsync(dOut, dd, N * 4, cudaMemcpyDeviceToHost, d.stream));
    if (worldPos3) TR_TRY(cudaMemcpyAsync(worldPos3, dp, N * 12, cudaMemcpyDeviceToHost, d.stream));
    unsigned long long h[kCounterSlots];
    TR_TRY(cudaMemcpyAsync(h, d.counters, sizeof h, cudaMemcpyDeviceToHost, d.stream));
    TR_TRY(cudaStreamSynchronize(d.stream));
#undef TR_TRY
    XnartCounters t{};
    t.raysPrimary = (uint64_t)n; t.hits = h[CTR_HITS];
    t.nBox = h[CTR_NBOX]; t.nTriRef = h[CTR_NTRIREF]; t.nTriCulled = h[CTR_NTRICULLED]; t.nTriFull = h[CTR_NTRIFULL];
    c->last = t;
    return XNART_OK;
}

int32_t xnart_get_counters(XnartCtx* c, XnartCounters* out) {
    if (!c || !out) return XNART_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(c->renderMutex);
    int rc = collectCounters(c);
    *out = c->last;
    return rc;
}

// RayTracer.Progress (RayTracer.cs:43-46: scanline / Height, polled from the game thread every 3 s, Game1.cs:331-341).
// While a blocking render call is in flight the kernels report the pixels they have finished into one page-locked word
// per in-flight state; may be called from any thread.
float xnart_progress(XnartCtx* c) {
    if (!c) return 0.0f;
    if (!c->rendering.load()) return c->progress.load();
    const unsigned long long total = c->progressTotal.load();
    if (total == 0) return 0.0f;
    unsigned long long done = 0;
    for (const DeviceState& d : c->devs)
        if (d.progressHost) done += *reinterpret_cast<volatile unsigned long long*>(d.progressHost);
    const double f = (double)done / (double)total;
    return (float)(f < 0.0 ? 0.0 : (f > 1.0 ? 1.0 : f));
}

int64_t xnart_kernel_launches(XnartCtx* c) { return c ? c->launches.load() : 0; }

// ---- device-memory utilities for one-process-per-GPU gathers ------------------------------------
// Finished frames of the other ranks can be written straight into a buffer of rank 0 with
// device-to-device copies over NVLink (copy engines, no SMs) instead of a collective whose receive
// kernels compete with rank 0's own rendering: rank 0 allocates the buffer and exports a CUDA IPC
// handle, the other ranks open it and copy into it on their own streams.
void* xnart_alloc_device(int32_t device, int64_t bytes) {
    void* p = nullptr;
    if (bytes <= 0 || cudaSetDevice(device) != cudaSuccess || cudaMalloc(&p, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    cudaMemset(p, 0, (size_t)bytes);
    return p;
}
void xnart_free_device(int32_t device, void* p) { if (p && cudaSetDevice(device) == cudaSuccess) cudaFree(p); }
int32_t xnart_ipc_export(int32_t device, const void* devPtr, uint8_t handle64[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    cudaIpcMemHandle_t h;
    if (!devPtr || !handle64 || cudaSetDevice(device) != cudaSuccess || cudaIpcGetMemHandle(&h, const_cast<void*>(devPtr)) != cudaSuccess) { cudaGetLastError(); return XNART_ERR_CUDA; }
    std::memcpy(handle64, &h, 64);
    return XNART_OK;
}
int32_t xnart_ipc_open(int32_t device, const uint8_t handle64[64], void** devPtrOut) {
    cudaIpcMemHandle_t h;
    if (!handle64 || !devPtrOut) return XNART_ERR_INVALID_ARGUMENT;
    std::memcpy(&h, handle64, 64);
    if (cudaSetDevice(device) != cudaSuccess || cudaIpcOpenMemHandle(devPtrOut, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return XNART_ERR_CUDA; }
    return XNART_OK;
}
int32_t xnart_ipc_close(int32_t device, void* devPtr) {
    if (!devPtr || cudaSetDevice(device) != cudaSuccess || cudaIpcCloseMemHandle(devPtr) != cudaSuccess) { cudaGetLastError(); return XNART_ERR_CUDA; }
    return XNART_OK;
}
int32_t xnart_copy_async(int32_t device, void* dst, const void* src, int64_t bytes, void* cudaStream) {
    if (!dst || !src || bytes < 0) return XNART_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(device) != cudaSuccess || cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)cudaStream) != cudaSuccess) { cudaGetLastError(); return XNART_ERR_CUDA; }
    return XNART_OK;
}

// Page-locked host memory for the frame (so the device->host copy needs no staging).
void* xnart_alloc_host(int64_t bytes) {
    void* p = nullptr;
    if (bytes <= 0 || cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void xnart_free_host(void* p) { if (p) cudaFreeHost(p); }

// Structural export of the flattened octrees (host data; works in a host-only context).
// which: 0 = mesh tree `id`, 1 = scene tree.
int32_t xnart_debug_tree_info(XnartCtx* c, int32_t which, int32_t id, int32_t* nNodes, int32_t* nRefs, int32_t* nCons, int32_t* depth) {
    if (!c) return XNART_ERR_INVALID_ARGUMENT;
    const FlatTree* t = which == 1 ? &c->sceneTree : ((id >= 0 && id < (int)c->meshes.size()) ? &c->meshes[id].tree : nullptr);
    if (!t) return fail(c, XNART_ERR_INVALID_ARGUMENT, "bad tree id");
    *nNodes = (int32_t)t->nodeCount(); *nRefs = (int32_t)t->refs.size(); *nCons = (int32_t)(t->consMax.size() / 3); *depth = t->depth;
    return XNART_OK;
}
int32_t xnart_debug_tree_dump(XnartCtx* c, int32_t which, int32_t id, float* lo3, float* hi3, int32_t* w0, int32_t* w1,
                              float* consMax3, int32_t* refs) {
    if (!c) return XNART_ERR_INVALID_ARGUMENT;
    const FlatTree* t = which == 1 ? &c->sceneTree : ((id >= 0 && id < (int)c->meshes.size()) ? &c->meshes[id].tree : nullptr);
    if (!t) return fail(c, XNART_ERR_INVALID_ARGUMENT, "bad tree id");
    std::memcpy(lo3, t->lo.data(), t->lo.size() * 4);
    std::memcpy(hi3, t->hi.data(), t->hi.size() * 4);
    std::memcpy(w0, t->w0.data(), t->w0.size() * 4);
    std::memcpy(w1, t->w1.data(), t->w1.size() * 4);
    if (consMax3 && !t->consMax.empty()) std::memcpy(consMax3, t->consMax.data(), t->consMax.size() * 4);
    if (refs && !t->refs.empty()) std::memcpy(refs, t->refs.data(), t->refs.size() * 4);
    return XNART_OK;
}

// The 128-byte expansion records of mesh `id` as uploaded (32 floats per internal node, device_types.h), plus the
// global bases that its record indices and leaf reference positions are relative to.  Works in a host-only context.
int32_t xnart_debug_inner_dump(XnartCtx* c, int32_t id, float* records32, int32_t* innerBase, int32_t* refBase) {
    if (!c || !c->built || id < 0 || id >= (int)c->meshes.size() || !innerBase || !refBase)
        return fail(c, XNART_ERR_INVALID_ARGUMENT, "xnart_debug_inner_dump: bad arguments or scene not built");
    const size_t n = c->meshes[(size_t)id].tree.consMax.size() / 3;
    *innerBase = c->meshInnerBase[(size_t)id];
    *refBase = c->meshRefBase[(size_t)id];
    if (records32 && n) std::memcpy(records32, c->hostInner.data() + 32 * (size_t)*innerBase, n * 128);
    return XNART_OK;
}

}  // extern "C"
