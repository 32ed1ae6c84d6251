"""GPU parity tests proper: libxnart.so (through its C ABI) against the CPU oracle on the same
seeded inputs.  Bars (BASELINE.json north_star): closest-hit triangle index bit-exact, hit distance
within 1e-5 relative, RGB within 1/255 on >= 99.9 % of pixels.  What is asserted here is stricter:
triangle index, u, v, d and world position bit-identical for every ray, and every pixel identical."""
import numpy as np
import pytest

from xna_ray_trace_b200 import records, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def X(xnart_lib):
    import torch

    assert torch.cuda.is_available(), "the -m gpu tests need a CUDA device"
    return xnart_lib


def make_scene(name):
    return {
        "crate": scenes.crate_scene,
        "spheres": scenes.spheres_scene,
        "grid": lambda: scenes.grid_scene(120),
        "soup": lambda: scenes.soup_scene(120000, half=14.0),
        "many": scenes.many_objects_scene,
    }[name]()


def random_rays(n, extent, seed):
    rng = np.random.default_rng(seed)
    org = rng.uniform(-extent, extent, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    d[::7, 0] = 0.0   # |d| < 1e-6 branch of the slab test
    d[::11, 1] = 0.0
    d[::13, 2] = 1e-7
    return np.concatenate([org, d], axis=1).astype(np.float32)


def assert_hits_identical(a, b):
    for k in ("obj", "mesh", "tri"):
        bad = np.nonzero(a[k] != b[k])[0]
        assert bad.size == 0, "%s differs on %d rays, first %s" % (k, bad.size, bad[:5])
    for k in ("u", "v", "d", "pos"):
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k


# kernel schedules (XNART_OPT_TRAVERSAL): 0 = automatic (small opaque scenes such as the crate: ONE persistent
# path-kernel launch per frame, csrc/path.cu), 1 = one thread per ray, 2 = wavefront (persistent trace kernel;
# scenes whose OctreeSpatialManager tree has internal nodes — "many" — fall back to 1)
SCHEDULES = [0, 1, 2]


@pytest.mark.parametrize("schedule", SCHEDULES)
@pytest.mark.parametrize("name", ["crate", "spheres", "grid", "soup", "many"])
def test_trace_rays_bit_exact(X, oracle_mod, name, schedule):
    s = make_scene(name)
    f = s.frame(96, 64)
    rays = np.concatenate([oracle_mod.primary_rays(f), random_rays(20000, 16.0, 11)])
    o = oracle_mod.Oracle()
    scenes.upload(o, s)
    x = X.Xnart()
    x.set_option(records.OPT_TRAVERSAL, schedule)
    scenes.upload(x, s)
    want, got = o.trace_rays(rays), x.trace_rays(rays)
    assert (want["tri"] >= 0).sum() > 100
    assert_hits_identical(want, got)
    # secondary-style rays: start ON a surface, ignore the origin triangle (shared-mesh identity, Q6)
    hit = want["tri"] >= 0
    org = want["pos"][hit]
    rng = np.random.default_rng(3)
    d = rng.normal(size=org.shape)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays2 = np.concatenate([org, d], axis=1).astype(np.float32)
    ign = np.stack([want["mesh"][hit], want["tri"][hit]], axis=1).astype(np.int32)
    assert_hits_identical(o.trace_rays(rays2, ign), x.trace_rays(rays2, ign))
    x.close()


@pytest.mark.parametrize("schedule", SCHEDULES)
@pytest.mark.parametrize("name,wh", [("crate", (320, 240)), ("spheres", (256, 256)), ("grid", (240, 135)),
                                     ("soup", (160, 120)), ("many", (240, 180))])
def test_frames_identical(X, oracle_mod, name, wh, schedule):
    s = make_scene(name)
    f = s.frame(*wh)
    o = oracle_mod.Oracle()
    scenes.upload(o, s)
    x = X.Xnart()
    x.set_option(records.OPT_TRAVERSAL, schedule)
    scenes.upload(x, s)
    want = o.render_frame(f, n_threads=8)
    got = x.render_frame(f)
    diff = int((want != got).sum())
    assert diff == 0, "%d of %d pixels differ" % (diff, want.size)
    oc, xc = o.counters(), x.counters()
    for k in ("raysPrimary", "raysShadow", "raysReflection", "raysRefraction", "hits", "pixels"):
        assert getattr(oc, k) == getattr(xc, k), k
    x.close()


@pytest.mark.parametrize("schedule", SCHEDULES)
@pytest.mark.parametrize("name", ["crate", "grid", "many"])
def test_counters_equal_oracle_pruned(X, oracle_mod, name, schedule):
    """The instrumented kernels count exactly the work of the oracle's pruned traversal
    (N_box_min, triangle references, culled and full triangle tests: SURVEY.md §8d)."""
    s = make_scene(name)
    f = s.frame(120, 90)
    o = oracle_mod.Oracle(mode=oracle_mod.Oracle.MODE_PRUNED)
    scenes.upload(o, s)
    x = X.Xnart()
    x.set_option(records.OPT_COUNTERS, 1)
    x.set_option(records.OPT_TRAVERSAL, schedule)
    scenes.upload(x, s)
    assert np.array_equal(o.render_frame(f, n_threads=4), x.render_frame(f))
    oc, xc = o.counters(), x.counters()
    for k in ("nBox", "nTriRef", "nTriCulled", "nTriFull", "hits"):
        assert getattr(oc, k) == getattr(xc, k), (k, getattr(oc, k), getattr(xc, k))
    x.close()


def test_depths_modes_and_shards(X, oracle_mod):
    s = make_scene("crate")
    o = oracle_mod.Oracle()
    scenes.upload(o, s)
    x = X.Xnart()
    scenes.upload(x, s)
    for depth in (0, 1, 3):
        f = s.frame(160, 120, max_reflections=depth)
        assert np.array_equal(o.render_frame(f, n_threads=4), x.render_frame(f)), depth
    for am in (records.ADDRESS_CLAMP, records.ADDRESS_WRAP, records.ADDRESS_MIRROR):
        for fl in (records.FILTER_POINT, records.FILTER_BILINEAR):
            f = s.frame(160, 120, max_reflections=2, address_mode=am, filtering=fl)
            assert np.array_equal(o.render_frame(f, n_threads=4), x.render_frame(f)), (am, fl)
    # one cooperative launch per frame (optional on shallow scenes) == one launch per level (default)
    f = s.frame(200, 150)
    per_level = x.render_frame(f).copy()
    x.set_option(records.OPT_FUSE_FRAME, 1)
    n0 = x.kernel_launches()
    fused = x.render_frame(f).copy()
    assert x.kernel_launches() - n0 == 1
    assert np.array_equal(fused, per_level) and np.array_equal(fused, o.render_frame(f, n_threads=4))
    x.set_option(records.OPT_FUSE_FRAME, 0)
    f = s.frame(160, 120, semantics=records.SEMANTICS_MONOGAME)  # MonoGame: truncating colour pack, divide-form slab test
    o.set_semantics(records.SEMANTICS_MONOGAME)
    got = x.render_frame(f)
    assert np.array_equal(got, o.render_frame(f, n_threads=4)) and (got != x.render_frame(s.frame(160, 120))).any()
    o.set_semantics(records.SEMANTICS_XNA40)
    full = x.render_frame(s.frame(160, 120)).copy()
    parts = np.zeros_like(full)
    for r in range(3):
        fs = s.frame(160, 120, shard_index=r, shard_count=3, shard_band_rows=8)
        part = np.zeros_like(full)
        x.render_frame(fs, out=part)
        own = (np.arange(120) // 8) % 3 == r
        assert not part[~own].any()
        parts[own] = part[own]
    assert np.array_equal(parts, full)
    x.close()


def test_quirks_on_gpu(X, oracle_mod):
    from tests.test_oracle_kat import early_exit_mesh

    allt = early_exit_mesh()
    x = X.Xnart()
    x.add_mesh(allt, box=scenes.mesh_bounding_box(allt))
    x.add_object([0], np.eye(4), np.eye(4), scenes.mesh_bounding_box(allt))
    x.build_scene()
    r = x.trace_rays(np.array([[-0.5, -0.5, -2.0, 0, 0, 1]], dtype=np.float32))
    assert r["tri"][0] == 0 and abs(r["d"][0] - 2.5) < 1e-5  # Q1: farther triangle in the nearer leaf wins
    x.close()
    # Q6 shared-mesh ignore identity, Q7 object-space distance
    quad = scenes.make_triangles([(-1, -1, 0)], [(-1, 1, 0)], [(1, -1, 0)])
    x = X.Xnart()
    x.add_mesh(quad, box=scenes.mesh_bounding_box(quad))
    a = scenes.ObjectData([0], position=(0, 0, 0), scale=(2, 2, 2)); a.build_world()
    b = scenes.ObjectData([0], position=(0, 0, -3)); b.build_world()
    box = scenes.mesh_bounding_box(quad)
    x.add_object([0], a.world, a.inverse_world, box)
    x.add_object([0], b.world, b.inverse_world, box)
    x.build_scene()
    ray = np.array([[-0.5, -0.5, 5, 0, 0, -1]], dtype=np.float32)
    r = x.trace_rays(ray)
    assert (r["obj"][0], r["tri"][0]) == (0, 0) and abs(r["d"][0] - 2.5) < 1e-6
    assert x.trace_rays(ray, ignore=[[0, 0]])["obj"][0] == -1
    assert x.trace_rays(ray, ignore=[[0, 7]])["obj"][0] == 0
    # NaN / non-finite rays miss
    bad = np.array([[0, 0, 5, np.nan, 0, -1], [np.inf, 0, 5, 0, 0, -1], [0, 0, 5, 0, 0, 0]], dtype=np.float32)
    assert (x.trace_rays(bad)["tri"] == -1).all()
    x.close()


def test_busy_and_errors(X):
    s = make_scene("crate")
    x = X.Xnart()
    with pytest.raises(X.XnartError) as e:
        x.render_frame(s.frame(32, 24))
    assert e.value.code == records.XNART_ERR_NOT_BUILT
    scenes.upload(x, s)
    f = s.frame(32, 24)
    f.filtering = 7
    with pytest.raises(X.XnartError) as e:
        x.render_frame(f)
    assert e.value.code == records.XNART_ERR_INVALID_ARGUMENT
    assert x.progress() in (0.0, 1.0)
    x.render_frame(s.frame(32, 24))
    assert x.progress() == 1.0 and x.kernel_launches() > 0
    x.close()


def test_full_size_properties(X):
    """BASELINE sizes: properties that need no oracle run — determinism, shard union == full frame,
    pinned and pageable outputs agree, ray-count identities."""
    s = make_scene("crate")
    x = X.Xnart()
    scenes.upload(x, s)
    f = s.frame(1920, 1080)
    a = x.render_frame(f).copy()
    pin = X.PinnedFrame(1080, 1920)
    b = x.render_frame(f, out=pin)
    assert np.array_equal(a, b)
    c = x.counters()
    assert c.raysPrimary == 1920 * 1080 and c.raysShadow == c.hits * len(s.lights)
    assert c.raysReflection <= c.hits and (a >> 24 == 255).all()
    union = np.zeros_like(a)
    for r in range(2):
        x.render_frame(s.frame(1920, 1080, shard_index=r, shard_count=2, shard_band_rows=8), out=union)
    assert np.array_equal(union, a)
    pin.free()
    x.close()


def test_wavefront_equals_per_thread_at_full_size(X):
    """BASELINE size, no oracle needed: the persistent-kernel schedule and the one-thread-per-ray
    schedule are two executions of the same per-ray algorithm, so a 1080p frame of the 1M-triangle
    grid (depth 4, 2 lights: 6.9 M rays through every phase of k_trace, including the cooperative
    leaf tail) and the work counters must be identical."""
    s = scenes.grid_scene(707)
    x = X.Xnart()
    scenes.upload(x, s)
    f = s.frame(1920, 1080)
    x.set_option(records.OPT_COUNTERS, 1)
    x.set_option(records.OPT_TRAVERSAL, 2)
    a = x.render_frame(f).copy()
    ca = x.counters()
    x.set_option(records.OPT_TRAVERSAL, 1)
    b = x.render_frame(f)
    cb = x.counters()
    assert np.array_equal(a, b)
    for k in ("raysShadow", "raysReflection", "hits", "nBox", "nTriRef", "nTriCulled", "nTriFull"):
        assert getattr(ca, k) == getattr(cb, k), k
    assert ca.hits > 1_000_000 and len(np.unique(a)) > 100  # a grey-scale scene
    # determinism of the dynamic schedule: ray-to-lane assignment changes from run to run, pixels must not
    x.set_option(records.OPT_COUNTERS, 0)
    x.set_option(records.OPT_TRAVERSAL, 0)
    c = x.render_frame(f).copy()
    assert np.array_equal(a, c) and np.array_equal(c, x.render_frame(f))
    x.close()


@pytest.mark.parametrize("schedule", SCHEDULES)
def test_tiles_chunks_and_odd_sizes(X, oracle_mod, schedule):
    """8x4 pixel tiles (W % 8 == 0 and rows % 4 == 0), scanline order otherwise, and frames rendered in
    several arena chunks all give the oracle's frame."""
    s = make_scene("spheres")  # transparent: reflection + refraction queues, 4 objects sharing one mesh
    o = oracle_mod.Oracle()
    scenes.upload(o, s)
    x = X.Xnart()
    x.set_option(records.OPT_TRAVERSAL, schedule)
    scenes.upload(x, s)
    for wh in ((96, 64), (100, 64), (96, 62), (101, 59)):
        f = s.frame(*wh, max_reflections=3)
        want = o.render_frame(f, n_threads=8)
        assert np.array_equal(want, x.render_frame(f)), wh
        x.set_option(records.OPT_ARENA_BYTES, 3_000_000)   # a few thousand pixels per chunk
        assert np.array_equal(want, x.render_frame(f)), ("chunked", wh)
        x.set_option(records.OPT_ARENA_BYTES, 8 << 30)
    x.close()


def test_render_frames_flythrough(X):
    """xnart_render_frames == n calls of xnart_render_frame, into pageable and page-locked memory."""
    s = make_scene("crate")
    x = X.Xnart()
    scenes.upload(x, s)
    import math
    frames = [s.frame(320, 180, camera=scenes.CameraData((math.sin(0.4 * i) * 72.0, 35.0, math.cos(0.4 * i) * 72.0), (0.0, 11.8, 0.0)))
              for i in range(7)]
    single = np.stack([x.render_frame(f).copy() for f in frames])
    assert len({a.tobytes() for a in single}) == 7
    got = x.render_frames(frames)
    assert np.array_equal(single, got)
    assert x.counters().raysPrimary == 7 * 320 * 180
    pin = X.PinnedFrame(180, 320, frames=7)
    assert np.array_equal(single, x.render_frames(frames, out=pin.array))
    x.close()
    # several frames in flight on one GPU (a device id listed k times): same frames, host and device output
    import torch

    x3 = X.Xnart(devices=(0,), frames_in_flight=3)
    scenes.upload(x3, s)
    assert np.array_equal(single, x3.render_frames(frames, out=pin.array))
    assert np.array_equal(single[2], x3.render_frame(frames[2]))          # one frame is not split over the states
    dev = torch.zeros((7, 180, 320), dtype=torch.int32, device="cuda:0")
    x3.render_frames_device(frames, dev.data_ptr())
    assert np.array_equal(single, dev.cpu().numpy().view(np.uint32))
    st = torch.cuda.Stream()
    dev.zero_()
    torch.cuda.synchronize()
    x3.render_frames_device(frames, dev.data_ptr(), st.cuda_stream)
    st.synchronize()
    assert np.array_equal(single, dev.cpu().numpy().view(np.uint32))
    pin.free()
    x3.close()


def test_flythrough_into_avi_sink(X, tmp_path):
    """configs[3] in miniature: orbit frames through xnart_render_frames into the AviManager-style sink;
    what comes back out of the file equals what xnart_render_frame renders."""
    from xna_ray_trace_b200 import avi, flythrough

    s = make_scene("grid")
    x = X.Xnart()
    scenes.upload(x, s)
    cams = flythrough.orbit_cameras(6, radius=14.0, height_y=6.0)
    path = str(tmp_path / "fly.avi")
    mgr = avi.AviManager(path, False)
    sink = mgr.AddVideoStream(False, 30, 0, 160, 96, avi.PIXELFORMAT_24BPP_RGB)
    n, rays, _ = flythrough.render_flythrough(x, s, cams, 160, 96, sink=sink, batch=4)
    mgr.Close()
    assert n == 6 and rays > 6 * 160 * 96
    # the device-side DIB conversion writes the same file as the host-side AddFrame
    path2 = str(tmp_path / "fly2.avi")
    m2 = avi.AviManager(path2, False)
    s2 = m2.AddVideoStream(False, 30, 0, 160, 96, avi.PIXELFORMAT_24BPP_RGB)
    for c in cams:
        s2.AddFrame(x.render_frame(s.frame(160, 96, camera=c)))
    m2.Close()
    assert open(path, "rb").read() == open(path2, "rb").read()
    v = avi.AviManager(path, True).GetVideoStream()
    assert v.CountFrames == 6
    for i in (0, 3, 5):
        assert np.array_equal(v.GetBitmap(i), x.render_frame(s.frame(160, 96, camera=cams[i])))
    x.close()


def test_split_queries_give_the_same_frames(X, oracle_mod):
    """Small batches are traced as 8 segment queries per ray (k_trace<.., SPLIT>): the first non-empty segment
    reproduces the reference's first group of leaves.  Frames with splitting everywhere, nowhere and for the
    deep levels only are identical (and equal to the oracle's)."""
    for name, wh in (("grid", (320, 180)), ("soup", (200, 120))):
        s = make_scene(name)
        o = oracle_mod.Oracle()
        scenes.upload(o, s)
        x = X.Xnart()
        x.set_option(records.OPT_TRAVERSAL, 2)
        scenes.upload(x, s)
        f = s.frame(*wh)
        want = o.render_frame(f, n_threads=8)
        for split in (1 << 24, 0, 20000):
            x.set_option(records.OPT_SPLIT_RAYS, split)
            assert np.array_equal(want, x.render_frame(f)), (name, split)
        # axis-parallel directions take the slab test's containment branch inside split queries as well
        f2 = s.frame(*wh, camera=scenes.CameraData((0.0, 60.0, 0.0), (0.0, 0.0, 1e-3)))
        x.set_option(records.OPT_SPLIT_RAYS, 1 << 24)
        a = x.render_frame(f2).copy()
        x.set_option(records.OPT_SPLIT_RAYS, 0)
        assert np.array_equal(a, x.render_frame(f2))
        x.close()


@pytest.mark.parametrize("name", ["crate", "grid", "spheres"])
def test_cast_rays_equals_oracle(X, oracle_mod, name):
    """xnart_cast_rays: CastRay through arbitrary (sub-pixel) screen positions — the four casts of
    GetColorForQuadrant (RayTracer.cs:220-276) — gives the oracle's colours."""
    s = make_scene(name)
    o = oracle_mod.Oracle()
    scenes.upload(o, s)
    x = X.Xnart()
    scenes.upload(x, s)
    f = s.frame(160, 120, max_reflections=3)
    rng = np.random.default_rng(4)
    xy = rng.uniform(-2.0, 162.0, size=(20000, 2)).astype(np.float32)
    xy[:, 1] *= 0.75
    xy[:64] = np.stack(np.meshgrid(np.arange(8), np.arange(8)), -1).reshape(-1, 2)   # pixel centres as well
    want, got = o.cast_samples(f, xy), x.cast_rays(f, xy)
    assert np.array_equal(want, got) and len(np.unique(got)) > 20
    assert x.cast_rays(f, np.zeros((0, 2), dtype=np.float32)).size == 0
    # integer positions reproduce the frame's pixels
    img = x.render_frame(f)
    ys, xs = np.mgrid[0:120, 0:160]
    px = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.float32)
    assert np.array_equal(x.cast_rays(f, px).reshape(120, 160), img)
    x.close()


@pytest.mark.parametrize("schedule", SCHEDULES)
def test_edge_cases(X, oracle_mod, schedule):
    """Empty and degenerate inputs: a scene without objects, a mesh without triangles, a scene without lights,
    a 1x1 viewport, a mesh of exactly itemTreshold and itemTreshold + 1 triangles (no split / one split,
    MeshOctree.cs:86), and the deepest recursion the library accepts."""
    def both(scene):
        o = oracle_mod.Oracle()
        scenes.upload(o, scene)
        x = X.Xnart()
        x.set_option(records.OPT_TRAVERSAL, schedule)
        scenes.upload(x, scene)
        return o, x

    # no objects at all: every pixel is the background (RayTracer.cs:732)
    s = scenes.SceneData("empty")
    s.camera = scenes.CameraData((0.0, 2.0, 8.0), (0.0, 0.0, 0.0))
    s.lights.append(scenes.spot_light((0.0, 5.0, 20.0)))
    o, x = both(s)
    f = s.frame(40, 24)
    img = x.render_frame(f)
    assert (img == 0xFF000000).all() and np.array_equal(img, o.render_frame(f))
    assert (x.trace_rays(np.array([[0, 0, 5, 0, 0, -1]], dtype=np.float32))["tri"] == -1).all()
    assert x.counters().hits == 0
    x.close()

    # an object whose mesh has no triangles, next to a real one
    grid = make_scene("grid")
    s = scenes.SceneData("hollow")
    s.meshes.append(scenes.MeshData(scenes.make_triangles([], [], []), reflectiveness=0.5))
    s.meshes.append(grid.meshes[0])
    s.objects.append(scenes.ObjectData([0]))
    s.objects.append(scenes.ObjectData([1]))
    s.camera = grid.camera
    s.max_reflections = 2
    o, x = both(s)
    f = s.frame(64, 40)
    assert np.array_equal(x.render_frame(f), o.render_frame(f, n_threads=4))   # no lights: light sum 0 everywhere
    from xna_ray_trace_b200 import xna

    s.lights.append(scenes.directional_light(xna.normalize(xna.v3(0.3, 1.0, 0.2))))
    x.set_lights(s.lights); o.set_lights(s.lights)
    f1 = s.frame(1, 1)
    assert np.array_equal(x.render_frame(f1), o.render_frame(f1))
    f = s.frame(64, 40, max_reflections=29)      # kMaxLevels - 1
    assert np.array_equal(x.render_frame(f), o.render_frame(f, n_threads=4))
    with pytest.raises(X.XnartError):
        x.render_frame(s.frame(8, 8, max_reflections=30))
    x.close()

    # 50 triangles stay one leaf, 51 split once (BuildTree: Count > itemTreshold)
    for n_tri in (50, 51):
        t = scenes.grid_triangles(8)[:n_tri]
        s = scenes.SceneData("t%d" % n_tri)
        s.meshes.append(scenes.MeshData(t, reflectiveness=0.3))
        s.objects.append(scenes.ObjectData([0]))
        s.lights.append(scenes.spot_light((0.0, 80.0, 40.0)))
        s.camera = scenes.CameraData((0.0, 60.0, 90.0), (0.0, 0.0, 0.0))
        s.max_reflections = 2
        o, x = both(s)
        assert x.tree(0)["depth"] == (0 if n_tri == 50 else 1)
        f = s.frame(96, 64)
        assert np.array_equal(x.render_frame(f), o.render_frame(f, n_threads=4))
        x.close()


# ---------------------------------------------------------------------------------------------
# BASELINE.json sizes against the oracle (every pixel), and the path kernel against the level kernels
# ---------------------------------------------------------------------------------------------
def test_path_kernel_is_one_launch_and_equals_level_kernels(X):
    """Small opaque scenes render a frame with ONE persistent launch (csrc/path.cu).  It executes the same per-ray
    algorithm as the level kernels, so a 1080p crate frame, its ray counts and the instrumented work counters are
    identical; sample batches (xnart_cast_rays), shards and odd sizes (scanline pixel order) as well."""
    s = make_scene("crate")
    x = X.Xnart()
    scenes.upload(x, s)
    f = s.frame(1920, 1080)
    x.set_option(records.OPT_COUNTERS, 1)
    import torch

    dev = torch.zeros((1080, 1920), dtype=torch.int32, device="cuda:0")
    n0 = x.kernel_launches()
    x.render_frame_device(f, dev.data_ptr())
    assert x.kernel_launches() - n0 == 1          # the whole frame: one launch
    n0 = x.kernel_launches()
    a = x.render_frame(f).copy()
    assert x.kernel_launches() - n0 <= 4          # host output: a few slices, each copied while the next one renders
    assert np.array_equal(a, dev.cpu().numpy().view(np.uint32))
    ca = x.counters()
    x.set_option(records.OPT_PATH_KERNEL, 0)
    n0 = x.kernel_launches()
    b = x.render_frame(f).copy()
    assert x.kernel_launches() - n0 > 1
    cb = x.counters()
    assert np.array_equal(a, b)
    for k in ("raysPrimary", "raysShadow", "raysReflection", "raysRefraction", "hits", "pixels", "nBox", "nTriRef", "nTriCulled", "nTriFull"):
        assert getattr(ca, k) == getattr(cb, k), k
    x.set_option(records.OPT_COUNTERS, 0)
    rng = np.random.default_rng(9)
    xy = rng.uniform(-4.0, 1924.0, size=(50000, 2)).astype(np.float32)
    xy[:, 1] *= 1080.0 / 1920.0
    for wh, kw in (((1920, 1080), dict(shard_index=1, shard_count=3, shard_band_rows=8)), ((333, 217), {}), ((64, 4), {}), ((7, 3), {})):
        fr = s.frame(*wh, **kw)
        x.set_option(records.OPT_PATH_KERNEL, 0)
        want = x.render_frame(fr).copy()
        cw = x.cast_rays(fr, xy[:5000])
        x.set_option(records.OPT_PATH_KERNEL, 1)
        assert np.array_equal(want, x.render_frame(fr)), wh
        assert np.array_equal(cw, x.cast_rays(fr, xy[:5000])), wh
    assert np.array_equal(a, x.render_frame(f))   # run-to-run determinism of the dynamic pixel-to-lane schedule
    x.close()


@pytest.mark.parametrize("wh", [(640, 480), (1920, 1080)])
def test_crate_at_baseline_sizes_equals_oracle(X, oracle_mod, wh):
    """BASELINE.json configs[0] (640x480) and configs[1] (1920x1080, "pixel-diff vs CPU reference"): every pixel of
    the depth-8 Free_crate frame equals the oracle's, and so do the ray counts."""
    s = make_scene("crate")
    o = oracle_mod.Oracle()
    scenes.upload(o, s)
    x = X.Xnart()
    scenes.upload(x, s)
    f = s.frame(*wh)
    want = o.render_frame(f, n_threads=16)
    got = x.render_frame(f)
    diff = int((want != got).sum())
    assert diff == 0, "%d of %d pixels differ" % (diff, want.size)
    oc, xc = o.counters(), x.counters()
    for k in ("raysPrimary", "raysShadow", "raysReflection", "raysRefraction", "hits", "pixels"):
        assert getattr(oc, k) == getattr(xc, k), k
    assert len(np.unique(got)) > 1000
    x.close()


def test_monogame_semantics_equal_oracle(X, oracle_mod):
    """XNART_SEMANTICS_MONOGAME = MonoGame's truncating Color pack AND its divide-form BoundingBox.Intersects (nullable
    tMin / tMax, 0 only when the origin lies strictly between the planes: SURVEY.md F6 / A.1), on the device as in the
    oracle: whole frames (single-leaf trees, a deep octree, the transparent default scene, a split scene octree) and
    per-ray hits, including rays that start ON box faces, where the two frameworks disagree."""
    for name, wh in (("crate", (320, 240)), ("grid", (240, 135)), ("spheres", (128, 128)), ("many", (160, 120))):
        s = make_scene(name)
        o = oracle_mod.Oracle(semantics=records.SEMANTICS_MONOGAME)
        scenes.upload(o, s)
        x = X.Xnart()
        scenes.upload(x, s)
        f = s.frame(*wh, semantics=records.SEMANTICS_MONOGAME)
        want = o.render_frame(f, n_threads=8)
        got = x.render_frame(f)
        assert np.array_equal(want, got), (name, int((want != got).sum()))
        assert (got != x.render_frame(s.frame(*wh))).any(), name
        oc, xc = o.counters(), x.counters()
        for k in ("raysPrimary", "raysShadow", "raysReflection", "raysRefraction", "hits"):
            assert getattr(oc, k) == getattr(xc, k), (name, k)
        x.close()
    differ = 0
    for name in ("crate", "soup", "grid"):
        s = make_scene(name)
        o = oracle_mod.Oracle(semantics=records.SEMANTICS_MONOGAME)
        scenes.upload(o, s)
        ox = oracle_mod.Oracle()
        scenes.upload(ox, s)
        x = X.Xnart()
        scenes.upload(x, s)
        x.set_option(records.OPT_SEMANTICS, records.SEMANTICS_MONOGAME)
        rays = np.concatenate([oracle_mod.primary_rays(s.frame(96, 64)), random_rays(20000, 16.0, 11)])
        want = o.trace_rays(rays)
        assert_hits_identical(want, x.trace_rays(rays))
        hit = want["tri"] >= 0
        rng = np.random.default_rng(3)
        d = rng.normal(size=want["pos"][hit].shape)
        d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        rays2 = np.concatenate([want["pos"][hit], d], axis=1).astype(np.float32)   # origins on triangles, i.e. on box faces
        ign = np.stack([want["mesh"][hit], want["tri"][hit]], axis=1).astype(np.int32)
        w2 = o.trace_rays(rays2, ign)
        assert_hits_identical(w2, x.trace_rays(rays2, ign))
        differ += int((w2["tri"] != ox.trace_rays(rays2, ign)["tri"]).sum()) + int((want["tri"] != ox.trace_rays(rays)["tri"]).sum())
        x.set_option(records.OPT_SEMANTICS, records.SEMANTICS_XNA40)
        assert_hits_identical(ox.trace_rays(rays2, ign), x.trace_rays(rays2, ign))
        x.close()
    assert differ > 100, "the ray sets no longer tell the two slab forms apart"


def test_default_spheres_512_equals_oracle(X, oracle_mod):
    """The scene the reference ships (Game1.cs:44-45,98-138): 4 transparent spheres, 512x512, depth 8."""
    s = make_scene("spheres")
    o = oracle_mod.Oracle()
    scenes.upload(o, s)
    x = X.Xnart()
    scenes.upload(x, s)
    f = s.frame(512, 512)
    want = o.render_frame(f, n_threads=16)
    got = x.render_frame(f)
    assert np.array_equal(want, got), int((want != got).sum())
    x.close()


def test_grid_1M_1080p_sampled_rows_equal_oracle(X, oracle_mod):
    """The 1M-triangle octree scene of BASELINE.json at 1920x1080, depth 4, 2 lights: 64 scanlines spread over the
    frame (8 bands of 8 rows) are rendered by the oracle and must equal the GPU frame's rows."""
    s = scenes.grid_scene(707)
    o = oracle_mod.Oracle()
    scenes.upload(o, s)
    x = X.Xnart()
    scenes.upload(x, s)
    f = s.frame(1920, 1080)
    got = x.render_frame(f)
    want = np.zeros_like(got)
    rows = []
    for b in range(8):
        y0 = 40 + b * 130
        o.render_frame(f, n_threads=16, rows=(y0, y0 + 8), out=want)
        rows += list(range(y0, y0 + 8))
    bad = int((want[rows] != got[rows]).sum())
    assert bad == 0, "%d of %d sampled pixels differ" % (bad, len(rows) * 1920)
    assert len(np.unique(got[rows])) > 50
    x.close()


def test_busy_and_progress_under_concurrency(X):
    """RayTracer.RenderAsync throws while IsBusy (RayTracer.cs:62-63) and Progress is polled from another thread while
    the render runs (Game1.cs:331-341): a second call issued while a long render holds the context returns
    XNART_ERR_BUSY, and xnart_progress moves through (0, 1) without ever going backwards."""
    import threading
    import time

    s = scenes.grid_scene(707)
    x = X.Xnart()
    scenes.upload(x, s)
    small = s.frame(64, 36)
    x.render_frame(small)
    assert x.progress() == 1.0
    frames = [s.frame(1920, 1080, camera=scenes.CameraData((np.sin(0.2 * i) * 80.0, 40.0, np.cos(0.2 * i) * 80.0), (0.0, 0.0, 0.0)))
              for i in range(24)]
    out = np.zeros((len(frames), 1080, 1920), dtype=np.uint32)
    done = []
    t = threading.Thread(target=lambda: done.append(x.render_frames(frames, out=out)))
    t.start()
    busy, seen = 0, []
    deadline = time.time() + 60.0
    while t.is_alive() and time.time() < deadline:
        seen.append(x.progress())
        try:
            x.render_frame(small)
        except X.BusyError as e:
            busy += 1
            assert e.code == records.XNART_ERR_BUSY
        for bad in (lambda: x.build_scene(), lambda: x.set_object_transform(0, np.eye(4), np.eye(4))):
            try:
                bad()
            except X.XnartError:
                pass
        time.sleep(0.002)
    t.join()
    assert done and busy > 0
    mid = [p for p in seen if 0.0 < p < 1.0]
    assert len(mid) >= 3 and len(set(mid)) >= 3, seen[:20]
    started = seen.index(mid[0])
    assert all(b >= a for a, b in zip(seen[started:], seen[started + 1:])), "progress went backwards"
    assert x.progress() == 1.0
    # the long call was not disturbed by the rejected ones
    assert np.array_equal(out[5], x.render_frame(frames[5]))
    # small scenes report per pixel batch (path kernel): a big crate frame shows intermediate values as well
    c = make_scene("crate")
    xc = X.Xnart()
    scenes.upload(xc, c)
    big = [c.frame(3840, 2160)] * 40
    outc = np.zeros((len(big), 2160, 3840), dtype=np.uint32)
    tc = threading.Thread(target=lambda: xc.render_frames(big, out=outc))
    tc.start()
    vals = []
    while tc.is_alive():
        vals.append(xc.progress())
        time.sleep(0.001)
    tc.join()
    assert len({v for v in vals if 0.0 < v < 1.0}) >= 3
    x.close()
    xc.close()


def test_multi_device_context_splits_one_frame(X):
    """xnart_create with several distinct GPUs: xnart_render_frame deals 8-row bands to the devices, xnart_render_frames
    whole frames; both give the single-GPU pictures.  Needs two GPUs (skipped on the one-GPU box)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    s = make_scene("crate")
    x1 = X.Xnart(devices=(0,))
    scenes.upload(x1, s)
    x2 = X.Xnart(devices=(0, 1))
    scenes.upload(x2, s)
    f = s.frame(640, 480)
    want = x1.render_frame(f).copy()
    assert np.array_equal(want, x2.render_frame(f))
    import math
    frames = [s.frame(320, 180, camera=scenes.CameraData((math.sin(0.4 * i) * 72.0, 35.0, math.cos(0.4 * i) * 72.0), (0.0, 11.8, 0.0)))
              for i in range(5)]
    assert np.array_equal(x1.render_frames(frames), x2.render_frames(frames))
    g = scenes.grid_scene(120)
    y1, y2 = X.Xnart(devices=(0,)), X.Xnart(devices=(0, 1))
    scenes.upload(y1, g); scenes.upload(y2, g)
    fg = g.frame(320, 200)
    assert np.array_equal(y1.render_frame(fg), y2.render_frame(fg))
    for c in (x1, x2, y1, y2):
        c.close()


def test_sub_shards_over_in_flight_states(X, oracle_mod):
    """Deep scenes: a context with k in-flight states on one GPU renders ONE frame (or one rank's shard of it) as k
    sub-shards of scanline bands side by side (xnart_render_frame_device).  Pixels and ray counts are those of a
    single-state context."""
    import torch

    s = make_scene("grid")
    x1 = X.Xnart(devices=(0,))
    x1.set_option(records.OPT_TRAVERSAL, 2)
    scenes.upload(x1, s)
    x3 = X.Xnart(devices=(0,), frames_in_flight=3)
    x3.set_option(records.OPT_TRAVERSAL, 2)
    scenes.upload(x3, s)
    for wh, kw in (((320, 200), {}), ((320, 200), dict(shard_index=1, shard_count=2, shard_band_rows=8)), ((203, 77), {}), ((64, 16), {})):
        f = s.frame(*wh, **kw)
        a = torch.zeros((wh[1], wh[0]), dtype=torch.int32, device="cuda:0")
        b = torch.zeros_like(a)
        x1.render_frame_device(f, a.data_ptr())
        c1 = x1.counters()
        x3.render_frame_device(f, b.data_ptr())
        c3 = x3.counters()
        assert torch.equal(a, b), (wh, kw)
        for k in ("raysPrimary", "raysShadow", "raysReflection", "hits", "pixels"):
            assert getattr(c1, k) == getattr(c3, k), (k, wh, kw)
    o = oracle_mod.Oracle()
    scenes.upload(o, s)
    f = s.frame(320, 200)
    b = torch.zeros((200, 320), dtype=torch.int32, device="cuda:0")
    x3.render_frame_device(f, b.data_ptr())
    assert np.array_equal(o.render_frame(f, n_threads=8), b.cpu().numpy().view(np.uint32))
    x1.close()
    x3.close()


def test_frames_to_dib24_equals_host_conversion(X):
    """xnart_frames_to_dib24 (VideoStream.AddFrame's RotateNoneFlipY + LockBits as 24bpp, VideoStream.cs:344-365) writes
    the bytes avi.to_dib_bits produces on the host — for widths with and without row padding and unaligned rows."""
    import torch

    from xna_ray_trace_b200 import avi

    rng = np.random.default_rng(5)
    for w, h, n in ((64, 8, 3), (101, 7, 2), (6, 5, 1), (1920, 16, 2), (3, 2, 4)):
        frames = rng.integers(0, 2 ** 32, size=(n, h, w), dtype=np.uint64).astype(np.uint32)
        stride = (3 * w + 3) & ~3
        src = torch.from_numpy(frames.view(np.int32)).cuda()
        dst = torch.full((n, h * stride), 0x55, dtype=torch.uint8, device="cuda:0")
        X.frames_to_dib24(0, src.data_ptr(), n, w, h, stride, dst.data_ptr())
        torch.cuda.synchronize()
        got = dst.cpu().numpy()
        for k in range(n):
            bits, st = avi.to_dib_bits(frames[k], 24)
            assert st == stride and bytes(got[k]) == bytes(bits), (w, h, k)


def test_keep_counters_and_split_rule(X):
    """XNART_OPT_KEEP_COUNTERS: statistics accumulate over render calls (exact ray totals of a flythrough).
    XNART_OPT_SPLIT_HIT_PCT / XNART_OPT_SPLIT_RAYS only choose how a batch is scheduled: frames do not change."""
    s = make_scene("grid")
    x = X.Xnart()
    x.set_option(records.OPT_TRAVERSAL, 2)
    scenes.upload(x, s)
    fa, fb = s.frame(160, 96), s.frame(200, 120)
    a = x.render_frame(fa).copy()
    ra = x.counters().rays
    x.render_frame(fb)
    rb = x.counters().rays
    x.set_option(records.OPT_KEEP_COUNTERS, 1)
    x.render_frame(fa); x.render_frame(fb); x.render_frames([fa, fa])
    c = x.counters()
    assert c.rays == 3 * ra + rb and c.pixels == 3 * 160 * 96 + 200 * 120
    x.set_option(records.OPT_KEEP_COUNTERS, 0)
    x.render_frame(fa)
    assert x.counters().rays == ra
    for pct in (0, 30, 100):
        for split in (1 << 24, 5000):
            x.set_option(records.OPT_SPLIT_HIT_PCT, pct)
            x.set_option(records.OPT_SPLIT_RAYS, split)
            assert np.array_equal(a, x.render_frame(fa)), (pct, split)
    x.close()


def test_measured_peaks_are_plausible(X):
    """xnart_measure_peak (the roofline denominators bench.py reports): un-fused FP32 within 60-105 % of SMs x 128 lanes x
    clock, and the memory hierarchy ordered L1 > L2 > HBM."""
    import torch

    p = torch.cuda.get_device_properties(0)
    nominal = p.multi_processor_count * 128 * 1.965e9
    fp32 = X.measure_peak(0, X.PEAK_FP32_NOFMA, reps=2)
    l1, l2, hbm = (X.measure_peak(0, k, reps=2) for k in (X.PEAK_L1_READ, X.PEAK_L2_READ, X.PEAK_HBM_READ))
    assert 0.6 * nominal < fp32 < 1.05 * nominal, fp32
    assert l1 > l2 > hbm > 2e12, (l1, l2, hbm)
    with pytest.raises(X.XnartError):
        X.measure_peak(0, 9)


def test_transparent_scene_arena_overflow_is_retried(X, oracle_mod):
    """Scenes with transparent materials: blocking calls size the path arena for a few nodes per pixel instead of the
    2^(D+1)-1 worst case, detect an overflow on the device and render again with more room.  scenes.forking_quads_scene
    forks at nearly every hit (48 rays per pixel at depth 6, far beyond the first guess): the frame, the ray counts, a
    following flythrough call and a batch of sample casts must still be the oracle's — from a fresh context each, so
    that every kind of blocking call meets the overflow itself."""
    import torch

    s = scenes.forking_quads_scene(6)
    o = oracle_mod.Oracle()
    scenes.upload(o, s)
    f = s.frame(96, 72)
    want = o.render_frame(f, n_threads=8)
    oc = o.counters()
    assert oc.rays > 40 * 96 * 72 and len(np.unique(want)) > 100
    near = s.frame(96, 72, camera=scenes.CameraData((0.0, 0.0, 30.0), (0.0, 0.0, 0.0)))
    want_near = o.render_frame(near, n_threads=8)
    xy = np.stack(np.meshgrid(np.arange(96), np.arange(72)), -1).reshape(-1, 2).astype(np.float32)

    def fresh():
        x = X.Xnart()
        scenes.upload(x, s)
        return x

    x = fresh()
    got = x.render_frame(f)
    assert np.array_equal(want, got), int((want != got).sum())
    xc = x.counters()
    for k in ("raysShadow", "raysReflection", "raysRefraction", "hits"):
        assert getattr(oc, k) == getattr(xc, k), k
    assert np.array_equal(x.render_frame(f), want)        # the scale sticks: no second retry, same frame
    x.close()
    x = fresh()
    two = x.render_frames([near, f])
    assert np.array_equal(two[0], want_near) and np.array_equal(two[1], want)
    x.close()
    x = fresh()
    assert np.array_equal(x.cast_rays(f, xy).reshape(72, 96), want)
    x.close()
    x = fresh()
    dev = torch.zeros((72, 96), dtype=torch.int32, device="cuda:0")
    x.render_frame_device(f, dev.data_ptr())
    assert np.array_equal(dev.cpu().numpy().view(np.uint32), want)
    # on a caller's stream nothing can be retried: the arena keeps its worst-case size and the frame is right at once
    st = torch.cuda.Stream()
    dev.zero_()
    torch.cuda.synchronize()
    x.render_frame_device(f, dev.data_ptr(), st.cuda_stream)
    st.synchronize()
    assert np.array_equal(dev.cpu().numpy().view(np.uint32), want)
    x.close()


@pytest.mark.parametrize("schedule", [0, 1])
def test_path_kernel_shortcuts_equal_oracle(X, oracle_mod, schedule):
    """The exact shortcuts of k_path on the inputs that select each side of them: a camera INSIDE the
    scene root's box (the root slab test is skipped for primary rays too) and one outside; an even-width top-down view
    whose centre column has |d.x| < 1e-6 (the general slab form for those warps, the fast one for the others); an
    axis-parallel directional light next to the spot light (every shadow ray of it has two parallel axes; black
    chains fold 0 / 255f)."""
    s = scenes.crate_scene()
    s.lights.append(scenes.directional_light((0.0, -1.0, 0.0), color=(0.6, 0.7, 0.8), intensity=0.7))
    s = s.finalize() if hasattr(s, "finalize") else s
    o = oracle_mod.Oracle()
    scenes.upload(o, s)
    x = X.Xnart()
    x.set_option(records.OPT_TRAVERSAL, schedule)
    scenes.upload(x, s)
    cams = [scenes.CameraData((10.0, 15.0, 20.0), (0.0, 8.0, 0.0)),                      # inside the root box
            scenes.CameraData((0.0, 55.0, 0.0), (0.0, 0.0, 0.0), up=(0.0, 0.0, -1.0)),   # straight down, inside
            scenes.CameraData((0.0, 300.0, 0.0), (0.0, 0.0, 0.0), up=(0.0, 0.0, -1.0)),  # straight down, outside
            scenes.CameraData((90.0, 20.0, -70.0), (0.0, 11.8, 0.0))]                     # outside, grazing the ground
    for cam, wh in zip(cams, [(160, 120), (160, 120), (160, 120), (200, 96)]):
        f = s.frame(*wh, camera=cam)
        want = o.render_frame(f, n_threads=8)
        got = x.render_frame(f)
        assert len(np.unique(want)) > 50, "a frame without content tests nothing"
        diff = int((want != got).sum())
        assert diff == 0, "%d of %d pixels differ (camera %s)" % (diff, want.size, cam.position)
        oc, xc = o.counters(), x.counters()
        for k in ("raysPrimary", "raysShadow", "raysReflection", "hits"):
            assert getattr(oc, k) == getattr(xc, k), k
    x.close()


def test_crate_orbit_cameras_equal_oracle(X, oracle_mod):
    """The bench's workload shape at a size the oracle renders in milliseconds: 24 cameras on the orbit of bench.py
    (radius 72, height 35), every pixel of every frame against the oracle, through the batch call the bench times."""
    import math
    s = scenes.crate_scene()
    o = oracle_mod.Oracle()
    scenes.upload(o, s)
    x = X.Xnart(frames_in_flight=2)
    scenes.upload(x, s)
    W, H = 480, 272
    frames = []
    for i in range(24):
        rot = 2.0 * math.pi * i / 24
        cam = scenes.CameraData((math.sin(rot) * 72.0, 35.0, math.cos(rot) * 72.0), (0.0, 11.8, 0.0))
        frames.append(s.frame(W, H, camera=cam))
    got = x.render_frames(frames)
    for i, f in enumerate(frames):
        want = o.render_frame(f, n_threads=8)
        diff = int((want != got[i]).sum())
        assert diff == 0, "camera %d: %d of %d pixels differ" % (i, diff, want.size)
    x.close()
