"""Frozen golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle).
CPU: the oracle (both traversal modes) still reproduces them.  GPU: libxnart.so reproduces them."""
import os

import numpy as np
import pytest

from xna_ray_trace_b200 import scenes

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FACTORY = {"crate": scenes.crate_scene, "spheres": scenes.spheres_scene, "grid48": lambda: scenes.grid_scene(48),
           "many": scenes.many_objects_scene}


def _load(name):
    return np.load(os.path.join(HERE, name + ".npz"))


def _check(api, g, s):
    hits = api.trace_rays(g["rays"])
    for k in ("obj", "mesh", "tri"):
        assert np.array_equal(hits[k], g[k]), k
    for k in ("u", "v", "d", "pos"):
        assert np.array_equal(hits[k].view(np.uint32), g[k].view(np.uint32)), k
    f = s.frame(int(g["width"]), int(g["height"]))
    img = api.render_frame(f)
    assert np.array_equal(img, g["image"])
    c = api.counters()
    assert [c.raysPrimary, c.raysShadow, c.raysReflection, c.raysRefraction, c.hits] == g["rays_by_type"].tolist()


@pytest.mark.parametrize("name", sorted(FACTORY))
@pytest.mark.parametrize("mode", [0, 1])
def test_oracle_reproduces_golden(oracle_mod, name, mode):
    s = FACTORY[name]()
    o = oracle_mod.Oracle(mode=mode)
    scenes.upload(o, s)
    _check(o, _load(name), s)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(FACTORY))
def test_gpu_reproduces_golden(xnart_lib, name):
    s = FACTORY[name]()
    x = xnart_lib.Xnart()
    scenes.upload(x, s)
    _check(x, _load(name), s)
    x.close()


def _check_multisample(api):
    """multisample.npz: sample colours through the ray-cast entry, and the multisampled frames through the host
    mirror's level-synchronous refinement on top of that entry."""
    from xna_ray_trace_b200 import host

    g = _load("multisample")
    for name in ("crate", "grid48"):
        s = FACTORY[name]()
        a = api()
        scenes.upload(a, s)
        f = s.frame(int(g["width"]), int(g["height"]), max_reflections=int(g["max_reflections"]))
        assert np.array_equal(a.cast_rays(f, g[name + "_xy"]), g[name + "_samples"]), name
        tracer = host.RayTracer()
        for q in (0, 2):
            tracer.MultisampleQuality = q
            got = tracer._render_with_multisampling(a, f, int(g["width"]), int(g["height"]))
            assert np.array_equal(got, g["%s_q%d" % (name, q)]), (name, q)
        a.close()


def test_oracle_reproduces_multisample_golden(oracle_mod):
    _check_multisample(oracle_mod.Oracle)


@pytest.mark.gpu
def test_gpu_reproduces_multisample_golden(xnart_lib):
    _check_multisample(xnart_lib.Xnart)
