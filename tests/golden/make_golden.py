#!/usr/bin/env python
"""Generates the frozen golden vectors under tests/golden/ from the CPU oracle.

The reference ships no golden vectors and cannot be executed here, so these are
produced by the oracle restatement (literal reference traversal, XNA 4.0 semantics) after it passed
the known-answer tests, and frozen so that later changes to either the oracle or libxnart.so are
caught.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from xna_ray_trace_b200 import scenes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (scene factory, frame size, number of random rays)
    "crate": (scenes.crate_scene, (320, 240), 2000),
    "spheres": (scenes.spheres_scene, (256, 256), 2000),
    "grid48": (lambda: scenes.grid_scene(48), (160, 90), 2000),
    "many": (scenes.many_objects_scene, (160, 120), 2000),
}


def golden_rays(frame, n, seed=17):
    rng = np.random.default_rng(seed)
    prim = O.primary_rays(frame)
    prim = prim[rng.choice(len(prim), size=min(n, len(prim)), replace=False)]
    org = rng.uniform(-12, 12, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    d[::9, 0] = 0.0
    return np.concatenate([prim, np.concatenate([org, d], axis=1)]).astype(np.float32)


def real_texture_golden():
    """Free_crate/Diffuse.bmp (the config-1/2 texture) cannot travel to the GPU box; the oracle's frame of the crate
    scene textured with it is frozen here, where /root/reference exists (tests/test_ingest.py)."""
    path = "/root/reference/Free_crate/Diffuse.bmp"
    if not os.path.exists(path):
        print("skipping the real-texture golden: %s is absent" % path)
        return
    from xna_ray_trace_b200 import bmp

    argb = bmp.read_bmp_argb(path)
    s = scenes.crate_scene(argb=argb)
    o = O.Oracle(mode=O.Oracle.MODE_LITERAL, semantics=0)
    scenes.upload(o, s)
    img = o.render_frame(s.frame(320, 240), n_threads=os.cpu_count())
    out = os.path.join(HERE, "crate_real_diffuse_320x240.npz")
    np.savez_compressed(out, frame=img, texture_sum=np.uint64(argb.astype(np.uint64).sum()))
    print("crate_real_diffuse", img.shape, "bytes", os.path.getsize(out))


def main():
    real_texture_golden()
    for name, (factory, (w, h), nr) in CASES.items():
        s = factory()
        o = O.Oracle(mode=O.Oracle.MODE_LITERAL, semantics=0)
        scenes.upload(o, s)
        f = s.frame(w, h)
        rays = golden_rays(f, nr)
        hits = o.trace_rays(rays)
        img = o.render_frame(f, n_threads=os.cpu_count())
        c = o.counters()
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), rays=rays, obj=hits["obj"], mesh=hits["mesh"], tri=hits["tri"],
            u=hits["u"], v=hits["v"], d=hits["d"], pos=hits["pos"], image=img, width=w, height=h,
            semantics=np.int32(0), max_reflections=np.int32(s.max_reflections),
            rays_by_type=np.array([c.raysPrimary, c.raysShadow, c.raysReflection, c.raysRefraction, c.hits], dtype=np.int64))
        print(name, "rays", len(rays), "hits", int((hits["tri"] >= 0).sum()), "image", img.shape,
              "bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))


def make_multisample():
    """Adaptive multisampling (RayTracer.cs:128-311) and its ray casts: one small frame per quality and a set of
    sub-pixel sample colours, from the oracle's literal recursion."""
    out = {}
    for name, factory in (("crate", scenes.crate_scene), ("grid48", lambda: scenes.grid_scene(48))):
        s = factory()
        o = O.Oracle(mode=O.Oracle.MODE_LITERAL, semantics=0)
        scenes.upload(o, s)
        f = s.frame(96, 72, max_reflections=2)
        rng = np.random.default_rng(23)
        xy = rng.uniform(-1.0, 97.0, size=(3000, 2)).astype(np.float32)
        xy[:, 1] *= 0.75
        out[name + "_xy"] = xy
        out[name + "_samples"] = o.cast_samples(f, xy)
        for q in (0, 2):
            out["%s_q%d" % (name, q)] = o.render_frame_multisample(f, q, n_threads=os.cpu_count())
    np.savez_compressed(os.path.join(HERE, "multisample.npz"), width=96, height=72, max_reflections=2, **out)
    print("multisample bytes", os.path.getsize(os.path.join(HERE, "multisample.npz")))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "multisample":
        make_multisample()
    else:
        main()
        make_multisample()
