"""Camera flythrough (BASELINE.json configs[3]): frames sharded across the GPUs of one box, gathered
to rank 0 and handed to the AviFile-style frame sink.

The reference orbits the camera with `(sin(rot) * 60, 4, cos(rot) * 60)` (Game1.cs:261,266) and renders
frame after frame; `compileVideo` (Game1.cs:192-210) then feeds the frames to
`AviManager.AddVideoStream` / `VideoStream.AddFrame`.  Here frame f is rendered by rank f % world
(`parallel.frame_owner`) through `xnart_render_frames` (device->host copy of a frame overlaps the
rendering of the next), rounds of `world` frames are gathered over NCCL (gloo on CPU) and rank 0
appends them to the sink in frame order.  No collective touches the traversal itself.
"""
import math
import time

import numpy as np

from . import scenes as _scenes
from .parallel import frame_owner


def orbit_cameras(n_frames, radius=90.0, height_y=30.0, target=(0.0, 0.0, 0.0)):
    """eye_f = (sin(rot) * r, h, cos(rot) * r), rot = 2*pi*f / n_frames."""
    cams = []
    for f in range(n_frames):
        rot = 2.0 * math.pi * f / n_frames
        cams.append(_scenes.CameraData((math.sin(rot) * radius, height_y, math.cos(rot) * radius), target))
    return cams


def dib_bits_on_device(frames, bpp, stride):
    """(..., H, W) int32 frames in the XNA Color layout on the device -> (..., H * stride) uint8 DIB bits
    (rows bottom-up, BGR, rows padded to `stride`): VideoStream.AddFrame's RotateNoneFlipY + LockBits (VideoStream.cs:
    345-351), done by libxnart's xnart_frames_to_dib24 kernel on the GPU (24 bpp only)."""
    import torch

    from . import capi

    assert bpp == 24, "the device conversion writes 24bpp DIBs"
    h, w = frames.shape[-2], frames.shape[-1]
    src = frames.contiguous()
    n = src.numel() // (h * w)
    out = torch.empty(tuple(frames.shape[:-2]) + (h * stride,), dtype=torch.uint8, device=frames.device)
    capi.frames_to_dib24(frames.device.index or 0, src.data_ptr(), n, w, h, stride, out.data_ptr(),
                         torch.cuda.current_stream(frames.device).cuda_stream)
    return out


_STAGING = {}


def _staging_buffers(shape, dtype, count=3):
    """Page-locked staging buffers of rank 0, kept between calls: pinning 1.2 GB (8 ranks x 8 rounds of 1080p DIB
    frames, three buffers) costs about half a second, which belongs to set-up, not to every flythrough."""
    key = (tuple(shape), dtype)
    if key not in _STAGING:
        import torch

        _STAGING.clear()      # one shape at a time: do not hoard pinned memory
        _STAGING[key] = [torch.empty(shape, dtype=dtype, pin_memory=True) for _ in range(count)]
    return _STAGING[key]


class NullSink:
    """A frame sink that discards what it is given (warm-up runs: allocates the staging buffers, writes nothing)."""

    def __init__(self, width, height, bits_per_pixel=24):
        self.Width, self.Height, self.CountBitsPerPixel = width, height, bits_per_pixel
        self._stride = (width * (bits_per_pixel // 8) + 3) & ~3
        self.CountFrames = 0

    def AddFrameBits(self, bits):
        self.CountFrames += 1

    def AddFrame(self, bmp):
        self.CountFrames += 1


def render_flythrough(x, scene, cameras, width, height, sink=None, world=1, rank=0, batch=8, device=None, dib24=None):
    """Render `cameras` (one frame each) on this rank's share and deliver the frames in order to `sink`
    (an avi.VideoStream, or any object with AddFrame / AddFrameBits) on rank 0.  Returns (frames rendered here,
    rays traced here — exact, from the device counters —, seconds until the last frame was delivered).
    dib24: ship 24bpp DIB rows instead of Color frames; every rank must pass the same value (default: rank 0's
    sink decides when world == 1, False otherwise).

    Frames stay on the device: every rank renders `batch` rounds into one half of a ring and, for a sink that takes
    raw DIB bits, converts them to 24bpp bottom-up rows there (xnart_frames_to_dib24: 3 bytes per pixel cross NVLink
    and PCIe instead of 4); ONE async collective (on NCCL's stream) ships the group to rank 0 while the next group
    is being rendered; rank 0 copies a gathered group into page-locked memory and a writer thread appends it to the
    sink while rendering goes on."""
    import queue
    import threading

    import torch

    from . import records

    n = len(cameras)
    mine = [f for f in range(n) if frame_owner(f, world) == rank]
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    dist = None
    if world > 1:
        import torch.distributed as dist
    stream = torch.cuda.current_stream()
    rounds = (n + world - 1) // world
    if dib24 is None:
        dib24 = world == 1 and sink is not None and hasattr(sink, "AddFrameBits") and getattr(sink, "CountBitsPerPixel", 0) == 24
    raw = bool(dib24)
    ring = torch.zeros((2, batch, height, width), dtype=torch.int32, device=device)
    if raw:
        stride = (width * 3 + 3) & ~3
        assert sink is None or sink._stride == stride
        unit_shape, unit_dtype = (batch, height * stride), torch.uint8
    else:
        unit_shape, unit_dtype = (batch, height, width), torch.int32
    ship = [torch.zeros(unit_shape, dtype=unit_dtype, device=device) for _ in range(2)] if raw else [ring[0], ring[1]]
    if world > 1 and rank == 0:
        recv = [[torch.zeros(unit_shape, dtype=unit_dtype, device=device) for _ in range(world)] for _ in range(2)]
    else:
        recv = [None, None]
    pending = [None, None]   # (work or None, first round, rounds in the group)

    # rank 0: page-locked staging buffers and the writer thread
    writer = None
    if rank == 0 and sink is not None:
        free_q, work_q = queue.Queue(), queue.Queue()
        for b in _staging_buffers((world,) + tuple(unit_shape), unit_dtype):
            free_q.put(b)
        failure, busy = [], [0.0]

        def write_loop():
            while True:
                item = work_q.get()
                if item is None:
                    return
                buf, ev, r0, cnt = item
                try:
                    ev.synchronize()
                    tw = time.perf_counter()
                    host = buf.numpy()
                    for k in range(cnt):
                        for q in range(world):
                            if (r0 + k) * world + q < n and not failure:
                                if raw:
                                    sink.AddFrameBits(host[q, k])
                                else:
                                    sink.AddFrame(host[q, k].view(np.uint32))
                    if hasattr(sink, "Flush"):
                        sink.Flush()   # payload writers (AviManager io_threads) are done with this staging buffer
                    busy[0] += time.perf_counter() - tw
                except Exception as e:   # e.g. disk full: keep draining so the render loop never blocks
                    failure.append(e)
                free_q.put(buf)

        writer = threading.Thread(target=write_loop, daemon=True)
        writer.start()

    def deliver(half):
        work, r0, cnt = pending[half]
        if work is not None:
            work.wait()
        pending[half] = None
        if writer is None:
            return
        buf = free_q.get()
        if world > 1:
            for q in range(world):
                buf[q].copy_(recv[half][q], non_blocking=True)
        else:
            buf[0].copy_(ship[half], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(stream)
        work_q.put((buf, ev, r0, cnt))

    x.set_option(records.OPT_KEEP_COUNTERS, 1)   # exact ray totals over all batches of this rank
    t0 = time.perf_counter()
    for g0 in range(0, rounds, batch):
        half = (g0 // batch) & 1
        if pending[half] is not None:
            deliver(half)
        cnt = min(batch, rounds - g0)
        fl = [(g0 + k) * world + rank for k in range(cnt)]
        fl = [f for f in fl if f < n]      # a rank's frames of a group are consecutive slots 0..len-1
        if fl:
            x.render_frames_device([scene.frame(width, height, camera=cameras[f]) for f in fl], ring[half].data_ptr(), stream.cuda_stream)
            if raw:
                from . import capi

                capi.frames_to_dib24(device.index or 0, ring[half].data_ptr(), len(fl), width, height, stride,
                                     ship[half].data_ptr(), stream.cuda_stream)
        work = dist.gather(ship[half], recv[half], dst=0, async_op=True) if world > 1 else None
        pending[half] = (work, g0, cnt)
    for h in sorted([h for h in range(2) if pending[h] is not None], key=lambda h: pending[h][1]):
        deliver(h)
    torch.cuda.synchronize()
    if writer is not None:
        work_q.put(None)
        writer.join()
        if failure:
            x.set_option(records.OPT_KEEP_COUNTERS, 0)
            raise failure[0]
    secs = time.perf_counter() - t0
    render_flythrough.last_sink_busy_s = busy[0] if writer is not None else None   # seconds the sink spent in AddFrame*
    rays = x.counters().rays
    x.set_option(records.OPT_KEEP_COUNTERS, 0)
    return len(mine), rays, secs
