#!/usr/bin/env python
"""Device time of a 1080p frame of the 1M-triangle grid (depth 4, 2 lights), 1 and 2 frames in flight (tuning aid)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xna_ray_trace_b200 import capi, records, scenes

W, H = 1920, 1080
s = scenes.grid_scene(707)
tag = os.path.basename(os.environ.get("XNART_LIB", "default"))
for inflight in (1, 2):
    x = capi.Xnart(devices=(0,), frames_in_flight=inflight)
    scenes.upload(x, s)
    frames = [s.frame(W, H, camera=scenes.CameraData((np.sin(0.13 * i) * 80.0, 40.0, np.cos(0.13 * i) * 80.0), (0.0, 0.0, 0.0))) for i in range(12)]
    buf = torch.zeros((len(frames), H, W), dtype=torch.int32, device="cuda:0")
    st = torch.cuda.current_stream()
    x.render_frames_device(frames[:3], buf.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    x.render_frames_device(frames, buf.data_ptr(), st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / len(frames)
    x.render_frames_device(frames, buf.data_ptr(), None)
    rays = x.counters().rays / len(frames)
    print(json.dumps({"label": tag, "inflight": inflight, "ms_per_frame": round(ms, 3), "Mrays_s": round(rays / ms / 1e3, 1),
                      "checksum": int(buf.to(torch.int64).sum().item())}))
    x.close()
