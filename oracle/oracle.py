"""ctypes binding of oracle/libxnart_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (xna-ray-trace_b200/) never does.
The API object mirrors the method names of the libxnart binding so scenes.upload() can feed
either one.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from xna_ray_trace_b200.records import (TRIANGLE_DTYPE, XnartCounters, XnartFrame, XnartLight, XnartMaterial,
                                        f32_array, make_light_array, make_material)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libxnart_oracle.so")
_lib = None


def build(force=False):
    """Compile the oracle with the committed Makefile (g++, strict fp32 flags)."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("xnart_oracle.cpp", "xna_math.h", "Makefile")
    ):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.xo_last_error.restype = C.c_char_p
        _lib.xo_pack_color.restype = C.c_uint32
        _lib.xo_round_channel_premultiplied.restype = C.c_uint32
        _lib.xo_round_channel_premultiplied.argtypes = [C.c_float]
        _lib.xo_create_perspective_fov.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("oracle error %d: %s" % (code, msg))
        self.code = code


class Oracle:
    """Same surface as xna_ray_trace_b200.capi.Xnart, backed by the CPU restatement."""

    MODE_LITERAL, MODE_PRUNED = 0, 1

    def __init__(self, mode=0, semantics=0):
        self._l = lib()
        self._h = C.c_void_p()
        self._l.xo_create(C.byref(self._h))
        self._keep = []
        self.set_mode(mode)
        self.set_semantics(semantics)

    def close(self):
        if self._h:
            self._l.xo_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise OracleError(rc, self._l.xo_last_error(self._h).decode())

    def set_mode(self, mode):
        self._l.xo_set_mode(self._h, int(mode))

    def set_semantics(self, s):
        self._l.xo_set_semantics(self._h, int(s))

    def set_slab_override(self, s):
        """-1: the slab test follows XnartFrame.semantics; 0 / 1: XNA 4.0 / MonoGame form whatever the frame says."""
        self._l.xo_set_slab_override(self._h, int(s))

    def clear_objects(self):
        self._check(self._l.xo_clear_objects(self._h))

    def set_object_transform(self, obj_id, world, inverse_world):
        self._check(self._l.xo_set_object_transform(self._h, C.c_int32(obj_id), f32_array(world, 16), f32_array(inverse_world, 16)))

    def add_mesh(self, tris, reflectiveness=0.0, use_texture=False, interpolate_normals=False, transparent=False,
                 refraction_index=0.0, argb=None, box=None, item_threshold=50):
        tris = np.ascontiguousarray(tris, dtype=TRIANGLE_DTYPE)
        mat, keep = make_material(reflectiveness, use_texture, interpolate_normals, transparent, refraction_index, argb)
        b = f32_array(box, 6)
        out = C.c_int32(-1)
        self._check(self._l.xo_add_mesh(self._h, _p(tris), C.c_int64(len(tris)), C.byref(mat), b,
                                        C.c_int32(item_threshold), C.byref(out)))
        return out.value

    def add_object(self, mesh_ids, world, inverse_world, local_box):
        ids = (C.c_int32 * len(mesh_ids))(*mesh_ids)
        out = C.c_int32(-1)
        self._check(self._l.xo_add_object(self._h, ids, C.c_int32(len(mesh_ids)), f32_array(world, 16),
                                          f32_array(inverse_world, 16), f32_array(local_box, 6), C.byref(out)))
        return out.value

    def build_scene(self, item_threshold=20):
        self._check(self._l.xo_build_scene(self._h, C.c_int32(item_threshold)))

    def set_lights(self, lights):
        arr = make_light_array(lights)
        self._check(self._l.xo_set_lights(self._h, arr, C.c_int32(len(lights))))

    def render_frame(self, frame, n_threads=1, rows=None, out=None):
        if out is None:
            out = np.zeros((frame.vpH, frame.vpW), dtype=np.uint32)
        y0, y1 = rows if rows is not None else (0, frame.vpH)
        self._check(self._l.xo_render_rows(self._h, C.byref(frame), _p(out), C.c_int32(n_threads), C.c_int32(y0), C.c_int32(y1)))
        return out

    def render_frame_multisample(self, frame, quality, n_threads=1):
        """RayTracer.RenderInternalWithMultisampling (RayTracer.cs:128-311)."""
        out = np.zeros((frame.vpH, frame.vpW), dtype=np.uint32)
        self._check(self._l.xo_render_frame_multisample(self._h, C.byref(frame), C.c_int32(quality), _p(out), C.c_int32(n_threads)))
        return out

    def cast_samples(self, frame, xy):
        xy = np.ascontiguousarray(xy, dtype=np.float32).reshape(-1, 2)
        out = np.zeros(len(xy), dtype=np.uint32)
        self._check(self._l.xo_cast_samples(self._h, C.byref(frame), _p(xy), C.c_int64(len(xy)), _p(out)))
        return out

    cast_rays = cast_samples   # same name as capi.Xnart.cast_rays: the oracle can stand in as a host-mirror backend

    def trace_rays(self, rays6, ignore=None):
        rays6 = np.ascontiguousarray(rays6, dtype=np.float32).reshape(-1, 6)
        n = len(rays6)
        ign = None if ignore is None else np.ascontiguousarray(ignore, dtype=np.int32).reshape(-1, 2)
        res = dict(obj=np.zeros(n, np.int32), mesh=np.zeros(n, np.int32), tri=np.zeros(n, np.int32),
                   u=np.zeros(n, np.float32), v=np.zeros(n, np.float32), d=np.zeros(n, np.float32),
                   pos=np.zeros((n, 3), np.float32))
        self._check(self._l.xo_trace_rays(self._h, _p(rays6), C.c_int64(n), _p(ign) if ign is not None else None,
                                          _p(res["obj"]), _p(res["mesh"]), _p(res["tri"]), _p(res["u"]), _p(res["v"]),
                                          _p(res["d"]), _p(res["pos"])))
        return res

    def counters(self):
        c = XnartCounters()
        self._l.xo_get_counters(self._h, C.byref(c))
        return c

    def tree(self, mesh_id=None):
        """Flattened octree of a mesh (or of the scene when mesh_id is None)."""
        which, mid = (1, 0) if mesh_id is None else (0, mesh_id)
        nn, nr, dp = C.c_int32(), C.c_int32(), C.c_int32()
        self._check(self._l.xo_tree_info(self._h, which, mid, C.byref(nn), C.byref(nr), C.byref(dp)))
        n = nn.value
        out = dict(boxes=np.zeros((n, 6), np.float32), cons=np.zeros((n, 6), np.float32), first_child=np.zeros(n, np.int32),
                   ref_start=np.zeros(n, np.int32), ref_count=np.zeros(n, np.int32), refs=np.zeros(max(1, nr.value), np.int32),
                   depth=dp.value)
        self._check(self._l.xo_tree_dump(self._h, which, mid, _p(out["boxes"]), _p(out["cons"]), _p(out["first_child"]),
                                         _p(out["ref_start"]), _p(out["ref_count"]), _p(out["refs"])))
        out["refs"] = out["refs"][: nr.value]
        return out


def primary_rays(frame):
    out = np.zeros((frame.vpH * frame.vpW, 6), dtype=np.float32)
    lib().xo_primary_rays(C.byref(frame), _p(out))
    return out


# ---- unit helpers for the known-answer tests --------------------------------------------------
def slab(box6, ray6, semantics=0):
    t = C.c_float()
    hit = lib().xo_slab(f32_array(box6, 6), f32_array(ray6, 6), C.c_int32(semantics), C.byref(t))
    return (t.value if hit else None)


def tri_test(tri_record, ray6):
    tr = np.ascontiguousarray(tri_record, dtype=TRIANGLE_DTYPE).reshape(1)
    u, v, d = C.c_float(), C.c_float(), C.c_float()
    r = lib().xo_tri_test(_p(tr), f32_array(ray6, 6), C.byref(u), C.byref(v), C.byref(d))
    return r, u.value, v.value, d.value


def pack_color(rgb, semantics=0):
    return lib().xo_pack_color(f32_array(rgb, 3), C.c_int32(semantics))


def round_channel_premultiplied(t):
    return lib().xo_round_channel_premultiplied(C.c_float(t))


def unpack_color(c):
    out = (C.c_float * 3)()
    lib().xo_unpack_color(C.c_uint32(c), out)
    return np.array(list(out), dtype=np.float32)


def _m16(fn, *args):
    out = (C.c_float * 16)()
    fn(*args, out)
    return np.array(list(out), dtype=np.float32).reshape(4, 4)


def matrix_invert(m):
    return _m16(lib().xo_matrix_invert, f32_array(m, 16))


def matrix_multiply(a, b):
    return _m16(lib().xo_matrix_multiply, f32_array(a, 16), f32_array(b, 16))


def create_look_at(eye, target, up):
    return _m16(lib().xo_create_look_at, f32_array(eye, 3), f32_array(target, 3), f32_array(up, 3))


def create_perspective_fov(fov, aspect, zn, zf):
    out = (C.c_float * 16)()
    lib().xo_create_perspective_fov(float(fov), float(aspect), float(zn), float(zf), out)
    return np.array(list(out), dtype=np.float32).reshape(4, 4)


def lookup_uv(argb, uv, address_mode=1, filtering=0):
    mat, keep = make_material(use_texture=True, argb=argb)
    out = (C.c_float * 3)()
    lib().xo_lookup_uv(C.byref(mat), f32_array(uv, 2), C.c_int32(address_mode), C.c_int32(filtering), out)
    return np.array(list(out), dtype=np.float32)


def light_for_fragment(light, pos, normal):
    arr = make_light_array([light])
    out = (C.c_float * 3)()
    lib().xo_light_for_fragment(C.byref(arr[0]), f32_array(pos, 3), f32_array(normal, 3), out)
    return np.array(list(out), dtype=np.float32)


def _v3(fn, *args):
    out = (C.c_float * 3)()
    fn(*args, out)
    return np.array(list(out), dtype=np.float32)


def normalize(v):
    return _v3(lib().xo_normalize, f32_array(v, 3))


def reflect(v, n):
    return _v3(lib().xo_reflect, f32_array(v, 3), f32_array(n, 3))


def transform(p, m):
    return _v3(lib().xo_transform, f32_array(p, 3), f32_array(m, 16))
