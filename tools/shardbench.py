#!/usr/bin/env python
"""One rank's share of a tile split, timed on ONE GPU: shard 0 of N of a 3840x2160 frame of the 1M-triangle grid (or of a
1080p frame of the soup), with 1, 2 and 3 in-flight states (the shard is rendered as sub-shards side by side)."""
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
ref = None
for k in (1, 2, 3):
    x = capi.Xnart(devices=(0,), frames_in_flight=k)
    scenes.upload(x, s)
    f = s.frame(W, H, shard_index=0, shard_count=N, shard_band_rows=8)
    buf = torch.zeros((H, W), dtype=torch.int32, device="cuda:0")
    st = torch.cuda.current_stream()
    for _ in range(2):
        x.render_frame_device(f, buf.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5):
        x.render_frame_device(f, buf.data_ptr(), st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    chk = int(buf.to(torch.int64).sum().item())
    ref = chk if ref is None else ref
    print(json.dumps({"scene": s.name, "shard": "0 of %d" % N, "states": k, "ms": round(e0.elapsed_time(e1) / 5, 3), "same": chk == ref}), flush=True)
    x.close()
