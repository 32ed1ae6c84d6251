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


def render_flythrough(x, scene, cameras, width, height, sink=None, world=1, rank=0, batch=8, device=None):
    """Render `cameras` (one frame each) on this rank's share and deliver the frames in order to `sink`
    (an avi.VideoStream or any object with AddFrame) on rank 0.  Returns (frames rendered here, rays
    traced here, seconds spent rendering here)."""
    n = len(cameras)
    mine = [f for f in range(n) if frame_owner(f, world) == rank]
    rays, secs = 0, 0.0
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
    rounds = (n + world - 1) // world
    for r0 in range(0, rounds, batch):
        rs = list(range(r0, min(rounds, r0 + batch)))
        idx = [r * world + rank for r in rs if r * world + rank < n]
        out = None
        if idx:
            frames = [scene.frame(width, height, camera=cameras[f]) for f in idx]
            t0 = time.perf_counter()
            out = x.render_frames(frames)
            secs += time.perf_counter() - t0
            rays += x.counters().rays
        k = 0
        for r in rs:
            f_mine = r * world + rank
            have = f_mine < n
            frame = out[k] if have else np.zeros((height, width), dtype=np.uint32)
            k += 1 if have else 0
            if world == 1:
                if sink is not None:
                    sink.AddFrame(frame)
                continue
            t = torch.from_numpy(frame.view(np.int32).copy())
            if device is not None:
                t = t.to(device)
            if rank == 0:
                bufs = [torch.empty_like(t) for _ in range(world)]
                dist.gather(t, bufs, dst=0)
                for q in range(world):
                    if r * world + q < n and sink is not None:
                        sink.AddFrame(bufs[q].cpu().numpy().view(np.uint32))
            else:
                dist.gather(t, None, dst=0)
    return len(mine), rays, secs
