#!/usr/bin/env python
"""Device time of a 1080p frame of the 10M-triangle soup (configs[4], depth 8), one frame in flight (tuning aid)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xna_ray_trace_b200 import capi, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
s = scenes.soup_scene(n)
x = capi.Xnart(devices=(0,))
scenes.upload(x, s)
buf = torch.zeros((1080, 1920), dtype=torch.int32, device="cuda:0")
st = torch.cuda.current_stream()
f = s.frame(1920, 1080)
x.render_frame_device(f, buf.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(3):
    x.render_frame_device(f, buf.data_ptr(), st.cuda_stream)
e1.record(st)
torch.cuda.synchronize()
print(json.dumps({"label": os.path.basename(os.environ.get("XNART_LIB", "default")), "scene": s.name, "ms_per_frame": round(e0.elapsed_time(e1) / 3, 2),
                  "checksum": int(buf.to(torch.int64).sum().item())}), flush=True)
x.close()
