import sys, time, math
sys.path.insert(0, '/root/repo')
import numpy as np
from xna_ray_trace_b200 import capi, scenes
def cams(n): return [scenes.CameraData((math.sin(0.1*i)*72.0, 35.0, math.cos(0.1*i)*72.0), (0.0, 11.8, 0.0)) for i in range(n)]
for name in ("crate", "grid"):
    s = scenes.crate_scene() if name == "crate" else scenes.grid_scene(707)
    n = 60 if name == "crate" else 16
    cs = cams(n) if name == "crate" else [scenes.CameraData((math.sin(0.13*i)*80.0, 40.0, math.cos(0.13*i)*80.0), (0.0, 0.0, 0.0)) for i in range(n)]
    frames = [s.frame(1920, 1080, camera=c) for c in cs]
    ref = None
    for devs in ((0,), (0, 0), (0, 0, 0), (0, 0, 0, 0)):
        x = capi.Xnart(devices=devs)
        scenes.upload(x, s)
        pin = capi.PinnedFrame(1080, 1920, frames=n)
        x.render_frames(frames[:4], out=pin.array[:4])
        t0 = time.perf_counter(); out = x.render_frames(frames, out=pin.array); t = time.perf_counter() - t0
        rays = x.counters().rays
        if ref is None: ref = out.copy()
        same = bool(np.array_equal(ref, out))
        p1 = capi.PinnedFrame(1080, 1920)
        x.render_frame(frames[0], out=p1)
        t0 = time.perf_counter()
        for f in frames[:10]: x.render_frame(f, out=p1)
        t1 = (time.perf_counter() - t0) / 10
        same1 = bool(np.array_equal(p1.array, ref[9]))
        print("%s devices=%s render_frames: %.3f ms/frame %.0f Mrays/s identical=%s | render_frame (bands over the states): %.3f ms/frame identical=%s" % (name, devs, 1e3*t/n, rays/t/1e6, same, 1e3*t1, same1), flush=True)
        pin.free(); p1.free(); x.close()
