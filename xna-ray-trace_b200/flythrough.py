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
    (rows bottom-up, BGR(A), rows padded to `stride`): VideoStream.AddFrame's RotateNoneFlipY + LockBits, done
    with three torch ops on the GPU instead of numpy on the host."""
    import torch

    nb = bpp // 8
    h, w = frames.shape[-2], frames.shape[-1]
    px = frames.contiguous().view(torch.uint8).reshape(*frames.shape[:-2], h, w, 4)
    idx = torch.tensor([2, 1, 0, 3][:nb], device=frames.device)
    bgr = px.index_select(-1, idx).flip(-3).reshape(*frames.shape[:-2], h, w * nb)
    if stride != w * nb:
        bgr = torch.nn.functional.pad(bgr, (0, stride - w * nb))
    return bgr.reshape(*frames.shape[:-2], h * stride)


def render_flythrough(x, scene, cameras, width, height, sink=None, world=1, rank=0, batch=8, device=None):
    """Render `cameras` (one frame each) on this rank's share and deliver the frames in order to `sink`
    (an avi.VideoStream, or any object with AddFrame) on rank 0.  Returns (frames rendered here, rays
    traced here, seconds spent here).

    Frames stay on the device: every rank renders `batch` rounds into one half of a ring, ONE async
    collective (on NCCL's stream) ships the group to rank 0 while the next group is being rendered, and
    rank 0 converts a gathered group to DIB bits on the GPU and brings it to the host only to feed the sink."""
    import torch

    n = len(cameras)
    mine = [f for f in range(n) if frame_owner(f, world) == rank]
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    dist = None
    if world > 1:
        import torch.distributed as dist
    stream = torch.cuda.current_stream()
    rounds = (n + world - 1) // world
    ring = torch.zeros((2, batch, height, width), dtype=torch.int32, device=device)
    if world > 1 and rank == 0:
        recv = [[torch.zeros((batch, height, width), dtype=torch.int32, device=device) for _ in range(world)] for _ in range(2)]
    else:
        recv = [None, None]
    pending = [None, None]   # (work or None, first round, rounds in the group)
    raw = sink is not None and hasattr(sink, "AddFrameBits")

    def deliver(half):
        work, r0, cnt = pending[half]
        if work is not None:
            work.wait()
        pending[half] = None
        if rank != 0 or sink is None:
            return
        grp = torch.stack(recv[half]) if world > 1 else ring[half].unsqueeze(0)      # (world, batch, H, W)
        if raw:
            host = dib_bits_on_device(grp[:, :cnt], sink.CountBitsPerPixel, sink._stride).cpu().numpy()
        else:
            host = grp[:, :cnt].cpu().numpy().view(np.uint32)
        for k in range(cnt):
            for q in range(world):
                if (r0 + k) * world + q < n:
                    if raw:
                        sink.AddFrameBits(host[q, k].tobytes())
                    else:
                        sink.AddFrame(host[q, k])

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
        work = dist.gather(ring[half], recv[half], dst=0, async_op=True) if world > 1 else None
        pending[half] = (work, g0, cnt)
    for h in sorted([h for h in range(2) if pending[h] is not None], key=lambda h: pending[h][1]):
        deliver(h)
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    # rays of this rank: the device counters are per render; sample the middle frame of this rank's share
    rays = 0
    if mine:
        x.render_frame_device(scene.frame(width, height, camera=cameras[mine[len(mine) // 2]]), ring[0, 0].data_ptr(), None)
        rays = x.counters().rays * len(mine)
    return len(mine), rays, secs
