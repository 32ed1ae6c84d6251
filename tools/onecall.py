#!/usr/bin/env python
"""Where a synchronous xnart_render_frame call spends its time (1080p crate, page-locked output)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xna_ray_trace_b200 import capi, scenes
s = scenes.crate_scene()
x = capi.Xnart(devices=(0,))
scenes.upload(x, s)
pin = capi.PinnedFrame(1080, 1920)
f = s.frame(1920, 1080)
for _ in range(5):
    x.render_frame(f, out=pin)
t0 = time.perf_counter()
n = 200
for _ in range(n):
    x.render_frame(f, out=pin)
wall = (time.perf_counter() - t0) / n * 1e3
c = x.counters()
print("wall %.3f ms per call; device: total %.3f (levels %.3f, resolve %.3f, copy-after-render %.3f)" % (wall, c.msTotal, c.msLevels, c.msResolve, c.msCopy))
import torch
buf = torch.zeros((1080, 1920), dtype=torch.int32, device="cuda:0")
t0 = time.perf_counter()
for _ in range(n):
    x.render_frame_device(f, buf.data_ptr(), None)
print("device-output call: wall %.3f ms" % ((time.perf_counter() - t0) / n * 1e3))
x.close()
