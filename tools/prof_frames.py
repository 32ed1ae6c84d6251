#!/usr/bin/env python
"""Render a few frames of one scene on cuda:0, one frame per call — the command ncu wraps.
  python tools/prof_frames.py crate 1920 1080 [frames]      python tools/prof_frames.py grid 1920 1080 3"""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xna_ray_trace_b200 import capi, records, scenes

name, W, H = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 4
s = {"crate": scenes.crate_scene, "grid": lambda: scenes.grid_scene(707), "spheres": scenes.spheres_scene,
     "soup": lambda: scenes.soup_scene(10_000_000)}[name]()
x = capi.Xnart(devices=(0,))
for kv in sys.argv[5:]:
    k, v = kv.split("=")
    x.set_option(int(k), int(v))
scenes.upload(x, s)
buf = torch.zeros((H, W), dtype=torch.int32, device="cuda:0")
for i in range(n):
    if name == "crate":
        rot = 2.0 * math.pi * i / 64
        cam = scenes.CameraData((math.sin(rot) * 72.0, 35.0, math.cos(rot) * 72.0), (0.0, 11.8, 0.0))
    else:
        cam = None
    f = s.frame(W, H, camera=cam) if cam is not None else s.frame(W, H)
    x.render_frame_device(f, buf.data_ptr(), None)
    c = x.counters()
    print("frame %d: %.3f ms, %d rays" % (i, c.msLevels + c.msResolve, c.rays), flush=True)
x.close()
