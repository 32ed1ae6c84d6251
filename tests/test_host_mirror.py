"""The host-side mirror (xna_ray_trace_b200.host) used the way Game1.LoadContent uses the reference's
types (Game1.cs:74-141).  On CPU the backend is swapped for the oracle to check the plumbing (world
matrices, shared meshes, lights, frame description, IsBusy/RenderCompleted contract); the GPU test
runs the same script on libxnart.so."""
import math
import time

import numpy as np
import pytest

from xna_ray_trace_b200 import host, scenes, xna


def load_content(backend, size=(96, 96)):
    """Game1.LoadContent restated with the mirror types: 4 instances of one transparent sphere."""
    mat = host.Material(reflectiveness=0.7, useTexture=False, transparent=True, refractionIndex=1.32)
    mat.InterpolateNormals = True
    tris = scenes.uv_sphere_triangles(radius=2.0, color=(1.0, 0.0, 0.0, np.float32(100) / np.float32(255)))
    sphere = host.Mesh(tris, mat, scenes.mesh_bounding_box(tris))  # Model.Tag["Meshes"], shared by all instances
    scene = host.OctreeSpatialManager(backend=backend)
    for x in range(2):
        for y in range(2):
            obj = host.SceneObject([sphere], (-7.5 + 5 * x, 2, -7.5 + 5 * y), (0, 0, 0))
            obj.Scale = (1.0, 1.0, 1.0)
            obj.Name = "Sphere (%d,%d)" % (x, y)
            scene.Bodies.append(obj)
    scene.Build()
    camera = host.Camera((0, 16, 32), (0, 0, 0), (0, 1, 0), math.pi / 4, xna.aspect_ratio(*size), 1.0, 1000.0)
    tracer = host.RayTracer()
    tracer.TextureFiltering = host.TextureFiltering.Point
    tracer.AddressMode = host.UVAddressMode.Wrap
    tracer.CurrentCamera = camera
    tracer.CurrentScene = scene
    tracer.Viewport = size
    tracer.MaxReflections = 8
    tracer.UseMultisampling = False
    light = host.SpotLight()
    light.Color = (1, 1, 1)
    light.Position = (0, 5, 20)
    light.Direction = -xna.normalize(xna.v3(0, 5, 20))
    light.SpotAngle = math.pi / 2
    light.Intensity = 1.0
    tracer.Lights.append(light)
    return tracer


def reference_image(oracle_mod, size):
    s = scenes.spheres_scene()
    o = oracle_mod.Oracle()
    scenes.upload(o, s)
    return o.render_frame(s.frame(*size), n_threads=4)


def test_mirror_matches_direct_upload(oracle_mod):
    size = (96, 96)
    tracer = load_content(oracle_mod.Oracle(), size)
    done = []
    tracer.RenderCompleted.append(lambda sender, err: done.append(err))
    img = tracer.Render()
    assert done == [None] and not tracer.IsBusy
    assert np.array_equal(img, reference_image(oracle_mod, size))
    hit = tracer.CurrentScene.GetRayIntersection((-7.5, 10, -7.5), (0, -1, 0))
    assert hit is not None and abs(hit["d"] - 6.0) < 0.05


def test_busy_contract(oracle_mod):
    tracer = load_content(oracle_mod.Oracle(), (384, 384))
    tracer.RenderCompleted.append(lambda s, e: None)
    t = tracer.RenderAsync()
    try:
        if tracer.IsBusy:  # the render is still in flight: a second request must throw
            with pytest.raises(host.InvalidOperationException):
                tracer.RenderAsync()
    finally:
        t.join()
    assert not tracer.IsBusy
    tracer.Render()  # and it can render again afterwards


def test_bad_model():
    with pytest.raises(host.ArgumentException):
        host.SceneObject(["not a mesh"])


@pytest.mark.parametrize("quality", [0, 1, 2])
def test_multisampling_host_logic_equals_literal_recursion(oracle_mod, quality):
    """UseMultisampling (RayTracer.cs:128-311): the mirror runs GetColorForQuadrant level by level in batches
    of ray casts; the oracle restates the recursion literally (including the :305 slip).  With the same ray
    caster behind both, the frames must be identical — on the mirror's own scene and on the crate scene, whose
    white ground against the black background makes the refinement test (:287-306) fire."""
    size = (56, 40)
    tracer = load_content(oracle_mod.Oracle(), size)
    tracer.UseMultisampling = True
    tracer.MultisampleQuality = quality
    img = tracer.Render()
    s = scenes.spheres_scene()
    o = oracle_mod.Oracle()
    scenes.upload(o, s)
    want = o.render_frame_multisample(s.frame(*size), quality, n_threads=4)
    assert np.array_equal(img, want)
    assert (want != o.render_frame(s.frame(*size), n_threads=4)).any()     # supersampling changes edge pixels

    s = scenes.crate_scene()
    o = oracle_mod.Oracle()
    scenes.upload(o, s)
    f = s.frame(64, 48, max_reflections=2)
    want = o.render_frame_multisample(f, quality, n_threads=4)
    got = tracer._render_with_multisampling(o, f, 64, 48)
    assert np.array_equal(got, want)
    if quality > 0:
        assert (want != o.render_frame_multisample(f, 0, n_threads=4)).any()   # refinement happened
        # the :305 slip (refined lower-right colour stored into urColor) is part of the contract and is exercised
        # here: the "repaired" recursion gives a different frame
        tracer.LR_DESTINATION = 3
        assert (tracer._render_with_multisampling(o, f, 64, 48) != want).any()
        tracer.LR_DESTINATION = 1


def test_color_helpers_match_oracle(oracle_mod):
    rng = np.random.default_rng(1)
    v = rng.uniform(-0.2, 1.2, size=(500, 3)).astype(np.float32)
    v[0] = (0.5, 2.5 / 255.0, np.nan)
    got = xna.color_from_vector3(v)
    want = np.array([oracle_mod.pack_color(tuple(float(c) for c in row)) for row in v], dtype=np.uint32)
    assert np.array_equal(got, want)
    back = xna.color_to_vector3(got)
    for i in (0, 7, 99):
        assert np.array_equal(back[i], np.array(oracle_mod.unpack_color(int(got[i])), dtype=np.float32))


@pytest.mark.gpu
def test_mirror_on_gpu(xnart_lib, oracle_mod):
    size = (256, 256)
    tracer = load_content(None, size)  # default backend: libxnart on cuda:0
    img = tracer.Render()
    assert np.array_equal(img, reference_image(oracle_mod, size))
    assert tracer.Progress == 1.0
    # adaptive multisampling on the GPU ray caster == the oracle's literal recursion
    tracer.Viewport = (96, 72)
    tracer.CurrentCamera = host.Camera((0, 16, 32), (0, 0, 0), (0, 1, 0), math.pi / 4, xna.aspect_ratio(96, 72), 1.0, 1000.0)
    tracer.UseMultisampling = True
    tracer.MultisampleQuality = 2
    ms = tracer.Render().copy()
    s = scenes.spheres_scene()
    o = oracle_mod.Oracle()
    scenes.upload(o, s)
    assert np.array_equal(ms, o.render_frame_multisample(s.frame(96, 72), 2, n_threads=8))
    tracer.UseMultisampling = False
    tracer.TextureFiltering = 9
    errs = []
    tracer.RenderCompleted.append(lambda s, e: errs.append(e))
    tracer.Render()
    assert isinstance(errs[0], host.ArgumentException)


def _rebuild_and_move(backend, oracle_mod):
    """OctreeSpatialManager.Build() rebuilds from Bodies every time it is called (OctreeSpatialManager.cs:64-99), and
    SceneObject.World / InverseWorld are read per ray (:358-364): a second Build() must not duplicate the scene, and a
    body moved after Build() renders at its new pose — inside the octree (and root bounds) it was built with."""
    size = (80, 80)
    tracer = load_content(backend, size)
    first = tracer.Render().copy()
    scene = tracer.CurrentScene
    scene.Build()                                   # idempotent: same picture, same amount of work
    assert np.array_equal(tracer.Render(), first)
    rays_once = backend.counters().rays
    scene.Build()
    tracer.Render()
    assert backend.counters().rays == rays_once
    # move one sphere (staying inside the scene's root bounds) without calling Build(): the picture is the one of a
    # scene built at the new pose
    scene.Bodies[0].Position = (-7.0, 2.0, -7.0)
    moved = tracer.Render().copy()
    assert (moved != first).any()
    s = scenes.spheres_scene()
    s.objects[0] = scenes.ObjectData([0], position=(-7.0, 2.0, -7.0))
    s.finalize()
    o = oracle_mod.Oracle()
    scenes.upload(o, s)
    assert np.array_equal(moved, o.render_frame(s.frame(*size), n_threads=4))
    # a move that leaves the root bounds keeps the stale bounds, as in the reference (the tree is only built by Build())
    scene.Bodies[1].Position = (-7.5, 5.0, -2.5)
    stale = tracer.Render().copy()
    # appending a body takes effect with the next Build()
    scene.Bodies.append(host.SceneObject(scene.Bodies[0].Meshes, (0.0, 8.0, 0.0), (0, 0, 0)))
    assert np.array_equal(tracer.Render(), stale)
    scene.Build()
    rebuilt = tracer.Render().copy()
    assert (rebuilt != stale).any()
    return [first, moved, stale, rebuilt]


def test_build_is_idempotent_and_moved_bodies_render_moved(oracle_mod):
    _rebuild_and_move(oracle_mod.Oracle(), oracle_mod)


@pytest.mark.gpu
def test_build_is_idempotent_and_moved_bodies_render_moved_on_gpu(xnart_lib, oracle_mod):
    want = _rebuild_and_move(oracle_mod.Oracle(), oracle_mod)
    got = _rebuild_and_move(xnart_lib.Xnart(), oracle_mod)
    for a, b in zip(want, got):
        assert np.array_equal(a, b)
