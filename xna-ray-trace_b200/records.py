"""Flat records of include/xnart.h as numpy dtypes and ctypes structures.

The reference's `Triangle` (RayTracerTypeLibrary/Triangle.cs:14-25) and `Material`
(RayTracerTypeLibrary/Material.cs:28-39) are managed classes; the shim copies them once into
these blittable layouts.  Field order and sizes must match include/xnart.h exactly
(tests/test_abi.py checks the sizes against the compiled library).
"""
import ctypes as C

import numpy as np

# XnartTriangle: 32 x 4 bytes = 128 bytes
TRIANGLE_DTYPE = np.dtype(
    [
        ("v1", "<f4", (3,)), ("v2", "<f4", (3,)), ("v3", "<f4", (3,)),
        ("uv1", "<f4", (2,)), ("uv2", "<f4", (2,)), ("uv3", "<f4", (2,)),
        ("n1", "<f4", (3,)), ("n2", "<f4", (3,)), ("n3", "<f4", (3,)),
        ("surfaceNormal", "<f4", (3,)),
        ("color", "<f4", (4,)),
        ("id", "<i4"),
    ]
)
assert TRIANGLE_DTYPE.itemsize == 128


class XnartMaterial(C.Structure):
    _fields_ = [
        ("reflectiveness", C.c_float),
        ("useTexture", C.c_int32),
        ("interpolateNormals", C.c_int32),
        ("transparent", C.c_int32),
        ("refractionIndex", C.c_float),
        ("texW", C.c_int32),
        ("texH", C.c_int32),
        ("_pad", C.c_int32),
        ("argb", C.c_void_p),
    ]


class XnartLight(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("position", C.c_float * 3),
        ("direction", C.c_float * 3),
        ("color", C.c_float * 3),
        ("intensity", C.c_float),
        ("decayExponent", C.c_float),
        ("spotAngle", C.c_float),
    ]


class XnartFrame(C.Structure):
    _fields_ = [
        ("view", C.c_float * 16),
        ("proj", C.c_float * 16),
        ("vpX", C.c_int32), ("vpY", C.c_int32), ("vpW", C.c_int32), ("vpH", C.c_int32),
        ("minDepth", C.c_float), ("maxDepth", C.c_float),
        ("maxReflections", C.c_int32),
        ("filtering", C.c_int32),
        ("addressMode", C.c_int32),
        ("semantics", C.c_int32),
        ("shardIndex", C.c_int32), ("shardCount", C.c_int32), ("shardBandRows", C.c_int32),
        ("_pad", C.c_int32),
    ]


class XnartCounters(C.Structure):
    _fields_ = [
        ("raysPrimary", C.c_uint64), ("raysShadow", C.c_uint64),
        ("raysReflection", C.c_uint64), ("raysRefraction", C.c_uint64),
        ("hits", C.c_uint64), ("nBox", C.c_uint64), ("nTriRef", C.c_uint64),
        ("nTriCulled", C.c_uint64), ("nTriFull", C.c_uint64), ("pixels", C.c_uint64),
        ("msTotal", C.c_double), ("msLevels", C.c_double), ("msResolve", C.c_double), ("msCopy", C.c_double),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}

    @property
    def rays(self):
        return self.raysPrimary + self.raysShadow + self.raysReflection + self.raysRefraction


assert C.sizeof(XnartMaterial) == 40
assert C.sizeof(XnartLight) == 52
assert C.sizeof(XnartFrame) == 184
assert C.sizeof(XnartCounters) == 112

# status codes / enums (include/xnart.h)
XNART_OK = 0
XNART_ERR_INVALID_ARGUMENT = 1
XNART_ERR_BUSY = 2
XNART_ERR_CUDA = 3
XNART_ERR_NOT_BUILT = 4
XNART_ERR_OCTREE_DEPTH = 5
XNART_ERR_NO_DEVICE = 6
XNART_ERR_OUT_OF_MEMORY = 7
XNART_ERR_UNSUPPORTED = 8

FILTER_POINT, FILTER_BILINEAR = 0, 1
ADDRESS_CLAMP, ADDRESS_WRAP, ADDRESS_MIRROR = 0, 1, 2
LIGHT_SPOT, LIGHT_DIRECTIONAL = 0, 1
SEMANTICS_XNA40, SEMANTICS_MONOGAME = 0, 1

OPT_COUNTERS, OPT_ARENA_BYTES, OPT_SEMANTICS, OPT_TRAVERSAL, OPT_FUSE_FRAME, OPT_OVERLAP_LEVELS, OPT_SPLIT_RAYS, OPT_CHAIN_RESOLVE = 1, 2, 3, 4, 5, 6, 7, 8
OPT_PATH_KERNEL = 9
OPT_KEEP_COUNTERS = 10
OPT_SPLIT_HIT_PCT = 12


def f32_array(values, n):
    arr = (C.c_float * n)()
    flat = np.asarray(values, dtype=np.float32).reshape(-1)
    assert flat.size == n, (flat.size, n)
    for i in range(n):
        arr[i] = float(flat[i])
    return arr


def make_frame(view, proj, width, height, max_reflections=8, filtering=FILTER_POINT,
               address_mode=ADDRESS_WRAP, semantics=SEMANTICS_XNA40, vp_x=0, vp_y=0,
               min_depth=0.0, max_depth=1.0, shard_index=0, shard_count=1, shard_band_rows=1):
    """XnartFrame from the state RayTracer.Render captures (RayTracer.cs:395-397)."""
    f = XnartFrame()
    f.view = f32_array(view, 16)
    f.proj = f32_array(proj, 16)
    f.vpX, f.vpY, f.vpW, f.vpH = vp_x, vp_y, width, height
    f.minDepth, f.maxDepth = min_depth, max_depth
    f.maxReflections = max_reflections
    f.filtering = filtering
    f.addressMode = address_mode
    f.semantics = semantics
    f.shardIndex, f.shardCount, f.shardBandRows = shard_index, shard_count, shard_band_rows
    return f


def make_light_array(lights):
    arr = (XnartLight * max(1, len(lights)))()
    for i, L in enumerate(lights):
        arr[i].kind = L["kind"]
        arr[i].position = f32_array(L.get("position", (0, 0, 0)), 3)
        arr[i].direction = f32_array(L.get("direction", (0, 0, 0)), 3)
        arr[i].color = f32_array(L.get("color", (1, 1, 1)), 3)
        arr[i].intensity = L.get("intensity", 1.0)
        arr[i].decayExponent = L.get("decayExponent", 1.3)
        arr[i].spotAngle = L.get("spotAngle", 0.0)
    return arr


def make_material(reflectiveness=0.0, use_texture=False, interpolate_normals=False, transparent=False,
                  refraction_index=0.0, argb=None):
    """Returns (XnartMaterial, keepalive) — keep `keepalive` referenced while the struct is in use."""
    m = XnartMaterial()
    m.reflectiveness = reflectiveness
    m.useTexture = int(use_texture)
    m.interpolateNormals = int(interpolate_normals)
    m.transparent = int(transparent)
    m.refractionIndex = refraction_index
    keep = None
    if argb is not None:
        keep = np.ascontiguousarray(argb, dtype=np.uint32)
        m.texH, m.texW = keep.shape
        m.argb = keep.ctypes.data
    return m, keep
