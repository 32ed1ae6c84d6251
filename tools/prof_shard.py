#!/usr/bin/env python
"""Render shard 0 of N of a 3840x2160 frame of the 1M-triangle grid a few times (the command ncu wraps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xna_ray_trace_b200 import capi, scenes
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
s = scenes.grid_scene(707)
x = capi.Xnart(devices=(0,))
scenes.upload(x, s)
buf = torch.zeros((2160, 3840), dtype=torch.int32, device="cuda:0")
f = s.frame(3840, 2160, shard_index=0, shard_count=N, shard_band_rows=8)
for i in range(3):
    x.render_frame_device(f, buf.data_ptr(), None)
    c = x.counters()
    print("frame %d: %.3f ms levels, %.3f resolve, %d rays" % (i, c.msLevels, c.msResolve, c.rays), flush=True)
x.close()
