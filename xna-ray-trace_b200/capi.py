"""ctypes binding of libxnart.so (include/xnart.h).

This is the same ABI a C# P/Invoke shim binds.  There is no CPU fallback: if the
library has not been built the import of the *library* (not of this module) raises, and every
compute call on a box without a CUDA device raises `XnartError(XNART_ERR_NO_DEVICE)`.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .records import (TRIANGLE_DTYPE, XNART_ERR_BUSY, XNART_ERR_INVALID_ARGUMENT, XnartCounters, XnartFrame,
                      f32_array, make_light_array, make_material)

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("XNART_LIB") or os.path.join(_PKG, "libxnart.so")  # XNART_LIB: development override
_lib = None


class XnartError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("xnart error %d: %s" % (code, msg))
        self.code = code


class BusyError(XnartError):
    """RayTracer.RenderAsync throws InvalidOperationException while IsBusy (RayTracer.cs:62-63)."""


def build(force=False, verbose=False):
    """Compile libxnart.so in-tree with nvcc for sm_100a (csrc/Makefile)."""
    cmd = ["make", "-C", os.path.join(_PKG, "csrc")] + (["-B"] if force else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libxnart.so failed")
    return LIB_PATH


def lib():
    """Load libxnart.so; raises if it is missing (the product path never falls back to the CPU)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.xnart_last_error.restype = C.c_char_p
        L.xnart_last_error.argtypes = [C.c_void_p]
        L.xnart_version.restype = C.c_char_p
        L.xnart_progress.restype = C.c_float
        L.xnart_progress.argtypes = [C.c_void_p]
        L.xnart_kernel_launches.restype = C.c_int64
        L.xnart_kernel_launches.argtypes = [C.c_void_p]
        L.xnart_alloc_host.restype = C.c_void_p
        L.xnart_alloc_host.argtypes = [C.c_int64]
        L.xnart_free_host.argtypes = [C.c_void_p]
        L.xnart_set_option.argtypes = [C.c_void_p, C.c_int32, C.c_int64]
        L.xnart_render_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.xnart_render_frame_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


EXPORTS = [
    "xnart_create", "xnart_destroy", "xnart_last_error", "xnart_set_option", "xnart_add_mesh", "xnart_add_object",
    "xnart_build_scene", "xnart_clear_objects", "xnart_set_object_transform", "xnart_set_lights", "xnart_render_frame", "xnart_render_frame_device", "xnart_render_frames",
    "xnart_trace_rays", "xnart_cast_rays", "xnart_render_frames_device",
    "xnart_alloc_device", "xnart_free_device", "xnart_ipc_export", "xnart_ipc_open", "xnart_ipc_close", "xnart_copy_async",
    "xnart_get_counters", "xnart_progress", "xnart_kernel_launches", "xnart_version", "xnart_alloc_host",
    "xnart_free_host", "xnart_measure_peak", "xnart_frames_to_dib24", "xnart_debug_tree_info", "xnart_debug_tree_dump", "xnart_debug_inner_dump",
]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def alloc_device(device, nbytes):
    L = lib()
    L.xnart_alloc_device.restype = C.c_void_p
    L.xnart_alloc_device.argtypes = [C.c_int32, C.c_int64]
    p = L.xnart_alloc_device(device, nbytes)
    if not p:
        raise XnartError(7, "xnart_alloc_device failed")
    return p


def free_device(device, ptr):
    L = lib()
    L.xnart_free_device.argtypes = [C.c_int32, C.c_void_p]
    L.xnart_free_device(device, ptr)


def ipc_export(device, ptr):
    """64-byte CUDA IPC handle of a buffer from alloc_device (bytes object; send it to the other ranks)."""
    L = lib()
    h = (C.c_uint8 * 64)()
    L.xnart_ipc_export.argtypes = [C.c_int32, C.c_void_p, C.c_void_p]
    rc = L.xnart_ipc_export(device, ptr, h)
    if rc != 0:
        raise XnartError(rc, "xnart_ipc_export failed")
    return bytes(h)


def ipc_open(device, handle):
    L = lib()
    out = C.c_void_p()
    buf = (C.c_uint8 * 64).from_buffer_copy(handle)
    L.xnart_ipc_open.argtypes = [C.c_int32, C.c_void_p, C.c_void_p]
    rc = L.xnart_ipc_open(device, buf, C.byref(out))
    if rc != 0:
        raise XnartError(rc, "xnart_ipc_open failed")
    return out.value


def ipc_close(device, ptr):
    L = lib()
    L.xnart_ipc_close.argtypes = [C.c_int32, C.c_void_p]
    L.xnart_ipc_close(device, ptr)


def copy_async(device, dst, src, nbytes, stream=None):
    L = lib()
    L.xnart_copy_async.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    rc = L.xnart_copy_async(device, dst, src, nbytes, C.c_void_p(stream) if stream else None)
    if rc != 0:
        raise XnartError(rc, "xnart_copy_async failed")


class PinnedFrame:
    """A (H, W) — or (N, H, W) with frames=N — uint32 frame buffer in page-locked memory (xnart_alloc_host)."""

    def __init__(self, height, width, frames=None):
        self._l = lib()
        self.nbytes = height * width * 4 * (frames or 1)
        self.ptr = self._l.xnart_alloc_host(self.nbytes)
        if not self.ptr:
            raise XnartError(7, "xnart_alloc_host failed")
        buf = (C.c_uint32 * (self.nbytes // 4)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=np.uint32).reshape((height, width) if frames is None else (frames, height, width))

    def free(self):
        if self.ptr:
            self.array = None
            self._l.xnart_free_host(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Xnart:
    """One libxnart context (`new RayTracer()` + scene, Game1.cs:98-138)."""

    def __init__(self, devices=(0,), frames_in_flight=1):
        """devices: CUDA device ids; frames_in_flight k > 1 lists every device k times, i.e. k independent
        in-flight states per GPU for render_frames (see xnart_create in include/xnart.h)."""
        self._l = lib()
        self._h = C.c_void_p()
        devices = tuple(devices) * max(1, int(frames_in_flight))
        n = len(devices)
        ids = (C.c_int32 * max(1, n))(*devices)
        rc = self._l.xnart_create(C.byref(self._h), C.c_int32(n), ids)
        if rc != 0:
            raise XnartError(rc, self._l.xnart_last_error(None).decode())
        self.n_devices = n

    def close(self):
        if self._h:
            self._l.xnart_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            if rc == XNART_ERR_BUSY:
                raise BusyError(rc, self._l.xnart_last_error(None).decode())
            raise XnartError(rc, self._l.xnart_last_error(self._h).decode())

    def set_option(self, option, value):
        self._check(self._l.xnart_set_option(self._h, C.c_int32(option), C.c_int64(value)))

    def add_mesh(self, tris, reflectiveness=0.0, use_texture=False, interpolate_normals=False, transparent=False,
                 refraction_index=0.0, argb=None, box=None, item_threshold=50):
        tris = np.ascontiguousarray(tris, dtype=TRIANGLE_DTYPE)
        mat, _keep = make_material(reflectiveness, use_texture, interpolate_normals, transparent, refraction_index, argb)
        out = C.c_int32(-1)
        self._check(self._l.xnart_add_mesh(self._h, _p(tris), C.c_int64(len(tris)), C.byref(mat), f32_array(box, 6),
                                           C.c_int32(item_threshold), C.byref(out)))
        return out.value

    def add_object(self, mesh_ids, world, inverse_world, local_box):
        ids = (C.c_int32 * max(1, len(mesh_ids)))(*mesh_ids)
        out = C.c_int32(-1)
        self._check(self._l.xnart_add_object(self._h, ids, C.c_int32(len(mesh_ids)), f32_array(world, 16),
                                             f32_array(inverse_world, 16), f32_array(local_box, 6), C.byref(out)))
        return out.value

    def clear_objects(self):
        self._check(self._l.xnart_clear_objects(self._h))

    def set_object_transform(self, obj_id, world, inverse_world):
        self._check(self._l.xnart_set_object_transform(self._h, C.c_int32(obj_id), f32_array(world, 16), f32_array(inverse_world, 16)))

    def build_scene(self, item_threshold=20):
        self._check(self._l.xnart_build_scene(self._h, C.c_int32(item_threshold)))

    def set_lights(self, lights):
        arr = make_light_array(lights)
        self._check(self._l.xnart_set_lights(self._h, arr, C.c_int32(len(lights))))

    def render_frame(self, frame, out=None):
        """xnart_render_frame: host buffer in, host buffer out (copies inside the call)."""
        if out is None:
            out = np.zeros((frame.vpH, frame.vpW), dtype=np.uint32)
        arr = out.array if isinstance(out, PinnedFrame) else out
        assert arr.dtype == np.uint32 and arr.size == frame.vpH * frame.vpW and arr.flags["C_CONTIGUOUS"]
        self._check(self._l.xnart_render_frame(self._h, C.addressof(frame), arr.ctypes.data))
        return arr

    def render_frames(self, frames, out=None):
        """xnart_render_frames: a flythrough of n frames into an (n, H, W) uint32 host array (pipelined copies)."""
        n = len(frames)
        arr_t = XnartFrame * max(1, n)
        arr = arr_t(*frames)
        h, w = frames[0].vpH, frames[0].vpW
        if out is None:
            out = np.zeros((n, h, w), dtype=np.uint32)
        assert out.dtype == np.uint32 and out.size == n * h * w and out.flags["C_CONTIGUOUS"]
        self._l.xnart_render_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        self._check(self._l.xnart_render_frames(self._h, C.cast(arr, C.c_void_p), C.c_int32(n), out.ctypes.data))
        return out

    def render_frames_device(self, frames, dev_ptr, stream=None):
        """xnart_render_frames_device: n frames into n consecutive device frames at `dev_ptr`, all in-flight
        states of the context working side by side; ordered on / joined into `stream`."""
        n = len(frames)
        arr = (XnartFrame * max(1, n))(*frames)
        self._l.xnart_render_frames_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        self._check(self._l.xnart_render_frames_device(self._h, C.cast(arr, C.c_void_p), C.c_int32(n), C.c_void_p(dev_ptr),
                                                       C.c_void_p(stream) if stream else None))

    def inner_records(self, mesh_id=0):
        """xnart_debug_inner_dump: (records (n_inner, 32) float32, inner base, ref base) of a mesh octree."""
        n = int((self.tree(mesh_id)["w1"] < 0).sum())
        rec = np.zeros((max(1, n), 32), dtype=np.float32)[:n]
        ib, rb = C.c_int32(), C.c_int32()
        self._l.xnart_debug_inner_dump.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        self._check(self._l.xnart_debug_inner_dump(self._h, mesh_id, C.c_void_p(rec.ctypes.data) if n else None, C.byref(ib), C.byref(rb)))
        return rec, ib.value, rb.value

    def cast_rays(self, frame, screen_xy):
        """xnart_cast_rays: CastRay through n screen positions (x, y); returns n packed colours."""
        xy = np.ascontiguousarray(screen_xy, dtype=np.float32).reshape(-1, 2)
        out = np.zeros(len(xy), dtype=np.uint32)
        self._l.xnart_cast_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        self._check(self._l.xnart_cast_rays(self._h, C.addressof(frame), xy.ctypes.data, C.c_int64(len(xy)), out.ctypes.data))
        return out

    def render_frame_device(self, frame, dev_ptr, stream=None):
        """xnart_render_frame_device: output stays in device memory at `dev_ptr` (e.g. a torch tensor's data_ptr())."""
        self._check(self._l.xnart_render_frame_device(self._h, C.addressof(frame), C.c_void_p(dev_ptr),
                                                      C.c_void_p(stream) if stream else None))

    def trace_rays(self, rays6, ignore=None):
        rays6 = np.ascontiguousarray(rays6, dtype=np.float32).reshape(-1, 6)
        n = len(rays6)
        ign = None if ignore is None else np.ascontiguousarray(ignore, dtype=np.int32).reshape(-1, 2)
        res = dict(obj=np.zeros(n, np.int32), mesh=np.zeros(n, np.int32), tri=np.zeros(n, np.int32),
                   u=np.zeros(n, np.float32), v=np.zeros(n, np.float32), d=np.zeros(n, np.float32),
                   pos=np.zeros((n, 3), np.float32))
        self._check(self._l.xnart_trace_rays(self._h, _p(rays6), C.c_int64(n), _p(ign) if ign is not None else None,
                                             _p(res["obj"]), _p(res["mesh"]), _p(res["tri"]), _p(res["u"]),
                                             _p(res["v"]), _p(res["d"]), _p(res["pos"])))
        return res

    def counters(self):
        c = XnartCounters()
        self._check(self._l.xnart_get_counters(self._h, C.byref(c)))
        return c

    def progress(self):
        return float(self._l.xnart_progress(self._h))

    def kernel_launches(self):
        return int(self._l.xnart_kernel_launches(self._h))

    def tree(self, mesh_id=None):
        """Flattened octree (breadth order) of a mesh, or of the scene when mesh_id is None."""
        which, mid = (1, 0) if mesh_id is None else (0, mesh_id)
        nn, nr, nc, dp = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        self._check(self._l.xnart_debug_tree_info(self._h, which, mid, C.byref(nn), C.byref(nr), C.byref(nc), C.byref(dp)))
        n = nn.value
        out = dict(lo=np.zeros((n, 3), np.float32), hi=np.zeros((n, 3), np.float32), w0=np.zeros(n, np.int32),
                   w1=np.zeros(n, np.int32), cons_max=np.zeros((max(1, nc.value), 3), np.float32),
                   refs=np.zeros(max(1, nr.value), np.int32), depth=dp.value)
        self._check(self._l.xnart_debug_tree_dump(self._h, which, mid, _p(out["lo"]), _p(out["hi"]), _p(out["w0"]),
                                                  _p(out["w1"]), _p(out["cons_max"]), _p(out["refs"])))
        out["refs"] = out["refs"][: nr.value]
        out["cons_max"] = out["cons_max"][: nc.value]
        return out


def frames_to_dib24(device, src_ptr, n_frames, width, height, stride, dst_ptr, stream=None):
    """xnart_frames_to_dib24: n Color frames on the device -> n bottom-up 24bpp DIBs on the device (VideoStream.AddFrame)."""
    L = lib()
    L.xnart_frames_to_dib24.argtypes = [C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    rc = L.xnart_frames_to_dib24(device, C.c_void_p(src_ptr), n_frames, width, height, stride, C.c_void_p(dst_ptr),
                                 C.c_void_p(stream) if stream else None)
    if rc != 0:
        raise XnartError(rc, "xnart_frames_to_dib24 failed")


PEAK_FP32_NOFMA, PEAK_L1_READ, PEAK_L2_READ, PEAK_HBM_READ = 0, 1, 2, 3


def measure_peak(device, which, reps=5):
    """xnart_measure_peak: operations/s (PEAK_FP32_NOFMA) or bytes/s of one GPU, best of `reps` launches."""
    L = lib()
    out = C.c_double(0.0)
    L.xnart_measure_peak.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    rc = L.xnart_measure_peak(device, which, reps, C.byref(out))
    if rc != 0:
        raise XnartError(rc, "xnart_measure_peak failed")
    return out.value


def version():
    return lib().xnart_version().decode()
