#!/usr/bin/env python
"""XNART_OPT_SPLIT_RAYS on one rank's share of a tile split (shard 0 of N), soup or grid, on ONE GPU."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xna_ray_trace_b200 import capi, records, scenes
name = sys.argv[1]; N = int(sys.argv[2]); n = int(sys.argv[3]) if len(sys.argv) > 3 else 2_000_000
if name == "grid":
    s, W, H = scenes.grid_scene(707), 3840, 2160
else:
    s, W, H = scenes.soup_scene(n), 1920, 1080
x = capi.Xnart(devices=(0,))
scenes.upload(x, s)
buf = torch.zeros((H, W), dtype=torch.int32, device="cuda:0")
st = torch.cuda.current_stream()
f = s.frame(W, H, shard_index=0, shard_count=N, shard_band_rows=8) if N > 1 else s.frame(W, H)
ref = None
for split, pct in ((131072, 50), (131072, 100), (131072, 25), (131072, 75), (0, 50), (400000, 50)):
    x.set_option(records.OPT_SPLIT_RAYS, split)
    x.set_option(records.OPT_SPLIT_HIT_PCT, pct)
    for _ in range(2):
        x.render_frame_device(f, buf.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(3):
        x.render_frame_device(f, buf.data_ptr(), st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    chk = int(buf.to(torch.int64).sum().item()); ref = chk if ref is None else ref
    print(json.dumps({"scene": s.name, "shard": "0 of %d" % N, "split_rays": split, "hit_pct": pct, "ms": round(e0.elapsed_time(e1) / 3, 3), "same": chk == ref}), flush=True)
x.close()
