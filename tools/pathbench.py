#!/usr/bin/env python
"""Device time of a 1080p crate frame (one frame in flight, CUDA events) with the path kernel on and off."""
import os, sys, math, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from xna_ray_trace_b200 import capi, records, scenes

W, H = 1920, 1080
def cam(i, n=64):
    rot = 2.0 * math.pi * (i % n) / n
    return scenes.CameraData((math.sin(rot) * 72.0, 35.0, math.cos(rot) * 72.0), (0.0, 11.8, 0.0))

def run(label, opts, inflight=1, n=64):
    s = scenes.crate_scene()
    x = capi.Xnart(devices=(0,), frames_in_flight=inflight)
    for k, v in opts.items():
        x.set_option(k, v)
    scenes.upload(x, s)
    frames = [s.frame(W, H, camera=cam(i)) for i in range(n)]
    buf = torch.zeros((n, H, W), dtype=torch.int32, device="cuda:0")
    st = torch.cuda.current_stream()
    x.render_frames_device(frames[:8], buf.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    x.render_frames_device(frames, buf.data_ptr(), st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    x.render_frames_device(frames, buf.data_ptr(), None)
    rays = x.counters().rays / n
    chk = int(buf.to(torch.int64).sum().item())
    print(json.dumps({"label": label, "inflight": inflight, "ms_per_frame": round(ms, 4), "Mrays_s": round(rays / ms / 1e3, 1), "rays_per_frame": rays, "checksum": chk}))
    x.close()
    return chk

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "quick":   # tuning variants (XNART_LIB=...): the path kernel only
        tag = os.path.basename(os.environ.get("XNART_LIB", "default"))
        for k in (1, 2):
            run(tag, {}, inflight=k)
        sys.exit(0)
    a = run("path", {})
    b = run("levels", {records.OPT_PATH_KERNEL: 0})
    assert a == b, "frames differ"
    run("path", {}, inflight=2)
    run("path", {}, inflight=3)
    run("levels", {records.OPT_PATH_KERNEL: 0}, inflight=3)
