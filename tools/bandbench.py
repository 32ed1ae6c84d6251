#!/usr/bin/env python
"""Band height of a tile split, evaluated on ONE GPU: every shard i of N of a 3840x2160 frame of the 1M-triangle grid is
rendered alone and timed; the split's frame time is that of the slowest shard."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xna_ray_trace_b200 import capi, records, scenes

name = sys.argv[1] if len(sys.argv) > 1 else "grid"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8
if name == "grid":
    s, W, H = scenes.grid_scene(707), 3840, 2160
else:
    s, W, H = scenes.soup_scene(int(sys.argv[3]) if len(sys.argv) > 3 else 2_000_000), 1920, 1080
x = capi.Xnart(devices=(0,))
scenes.upload(x, s)
buf = torch.zeros((H, W), dtype=torch.int32, device="cuda:0")
st = torch.cuda.current_stream()
for rows in (4, 8, 16, 32, 64, 128, H // N):
    times = []
    for i in range(N):
        f = s.frame(W, H, shard_index=i, shard_count=N, shard_band_rows=rows)
        x.render_frame_device(f, buf.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(3):
            x.render_frame_device(f, buf.data_ptr(), st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / 3)
    print(json.dumps({"scene": s.name, "ranks": N, "band_rows": rows, "slowest_ms": round(max(times), 3), "fastest_ms": round(min(times), 3),
                      "mean_ms": round(float(np.mean(times)), 3)}), flush=True)
x.close()
