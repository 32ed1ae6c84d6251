#!/usr/bin/env python
"""tools/probe.py — time one configuration of BASELINE.json on cuda:0 and (optionally) check a
sample of its primary hits and pixels against the oracle.  Development aid, not the bench.

  python tools/probe.py crate 1920 1080 --frames 20
  python tools/probe.py grid 3840 2160 --n 707 --frames 5 --check
  python tools/probe.py soup 1920 1080 --n 10000000 --frames 3
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene", choices=["crate", "grid", "soup", "spheres", "many"])
    ap.add_argument("width", type=int)
    ap.add_argument("height", type=int)
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--depth", type=int, default=None)
    ap.add_argument("--frames", type=int, default=5)
    ap.add_argument("--check", action="store_true", help="compare a band of scanlines with the oracle")
    ap.add_argument("--count", action="store_true", help="run the instrumented kernels once and print work counters")
    ap.add_argument("--split", type=int, default=None, help="XNART_OPT_SPLIT_RAYS")
    ap.add_argument("--traversal", type=int, default=0, help="XNART_OPT_TRAVERSAL: 0 auto, 1 thread per ray, 2 octet per ray")
    args = ap.parse_args()

    from xna_ray_trace_b200 import capi, records, scenes

    t0 = time.time()
    if args.scene == "crate":
        s = scenes.crate_scene()
    elif args.scene == "grid":
        s = scenes.grid_scene(args.n or 707)
    elif args.scene == "soup":
        s = scenes.soup_scene(args.n or 1_000_000)
    elif args.scene == "spheres":
        s = scenes.spheres_scene()
    else:
        s = scenes.many_objects_scene()
    if args.depth is not None:
        s.max_reflections = args.depth
    t_gen = time.time() - t0
    x = capi.Xnart()
    x.set_option(records.OPT_TRAVERSAL, args.traversal)
    if args.split is not None:
        x.set_option(records.OPT_SPLIT_RAYS, args.split)
    t0 = time.time()
    scenes.upload(x, s)
    t_up = time.time() - t0
    tr = x.tree(0)
    print("scene %s: %d triangles, gen %.1fs, build+upload %.1fs, mesh0 tree: %d nodes, %d refs, depth %d" %
          (s.name, s.n_triangles, t_gen, t_up, len(tr["lo"]), len(tr["refs"]), tr["depth"]), flush=True)
    f = s.frame(args.width, args.height)
    pin = capi.PinnedFrame(args.height, args.width)
    x.render_frame(f, out=pin)
    ms, lv, rs, cp = [], [], [], []
    for i in range(args.frames):
        t0 = time.perf_counter()
        x.render_frame(f, out=pin)
        wall = (time.perf_counter() - t0) * 1e3
        c = x.counters()
        ms.append(c.msTotal); lv.append(c.msLevels); rs.append(c.msResolve); cp.append(c.msCopy)
    c = x.counters()
    print("frame %dx%d depth %d: total %.3f ms (levels %.3f, resolve %.3f, copy %.3f; wall %.3f) rays %d -> %.1f Mrays/s, %.1f fps" %
          (args.width, args.height, s.max_reflections, np.median(ms), np.median(lv), np.median(rs), np.median(cp), wall,
           c.rays, c.rays / (np.median(ms) * 1e-3) / 1e6, 1e3 / np.median(ms)))
    print("  rays: primary %d shadow %d reflection %d refraction %d hits %d" %
          (c.raysPrimary, c.raysShadow, c.raysReflection, c.raysRefraction, c.hits))
    if args.count:
        x.set_option(records.OPT_COUNTERS, 1)
        x.render_frame(f, out=pin)
        c = x.counters()
        x.set_option(records.OPT_COUNTERS, 0)
        rays = c.rays
        ab = 40 * rays + 32 * c.nBox + 4 * c.nTriRef + 24 * c.nTriCulled + 48 * c.nTriFull + 84 * c.hits + 4 * c.pixels
        af = 18 * c.nBox + 8 * c.nTriCulled + 51 * c.nTriFull + 120 * c.hits
        t = np.median(lv) * 1e-3
        print("  work: nBox %d (%.1f/ray) nTriRef %d (%.1f/ray) culled %d full %d | alg %.1f MB -> %.0f GB/s, %.2f T fp32-op/s" %
              (c.nBox, c.nBox / rays, c.nTriRef, c.nTriRef / rays, c.nTriCulled, c.nTriFull, ab / 1e6, ab / t / 1e9, af / t / 1e12))
    if args.check:
        from oracle import oracle as O

        o = O.Oracle()
        scenes.upload(o, s)
        y0 = args.height // 2 - 8
        t0 = time.time()
        ref = o.render_frame(f, n_threads=os.cpu_count(), rows=(y0, y0 + 16))
        diff = int((ref[y0:y0 + 16] != pin.array[y0:y0 + 16]).sum())
        print("  oracle band rows %d..%d: %d differing pixels of %d (%.1fs)" % (y0, y0 + 16, diff, 16 * args.width, time.time() - t0))
    x.close()


if __name__ == "__main__":
    main()
