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


def test_bad_model_and_multisampling():
    with pytest.raises(host.ArgumentException):
        host.SceneObject(["not a mesh"])
    tr = host.RayTracer()
    tr.UseMultisampling = True
    with pytest.raises(NotImplementedError):
        tr.RenderAsync()


@pytest.mark.gpu
def test_mirror_on_gpu(xnart_lib, oracle_mod):
    size = (256, 256)
    tracer = load_content(None, size)  # default backend: libxnart on cuda:0
    img = tracer.Render()
    assert np.array_equal(img, reference_image(oracle_mod, size))
    assert tracer.Progress == 1.0
    tracer.TextureFiltering = 9
    errs = []
    tracer.RenderCompleted.append(lambda s, e: errs.append(e))
    tracer.Render()
    assert isinstance(errs[0], host.ArgumentException)
