"""Known-answer tests that pin the CPU oracle.

The reference ships no tests or golden vectors for this path, so the oracle is
pinned by hand-computed answers for each building block and by cross-checking its two
traversal implementations (literal reference algorithm vs. the pruned, order-independent
restatement the CUDA kernels follow).
"""
import math

import numpy as np
import pytest

from xna_ray_trace_b200 import scenes, xna
from xna_ray_trace_b200.bmp import read_bmp_argb, synthetic_crate_texture, write_bmp24
from xna_ray_trace_b200.records import (ADDRESS_CLAMP, ADDRESS_MIRROR, ADDRESS_WRAP, FILTER_BILINEAR, FILTER_POINT,
                                        SEMANTICS_MONOGAME, SEMANTICS_XNA40)

F = np.float32


def one_tri(v1, v2, v3):
    return scenes.make_triangles([v1], [v2], [v3])


# ---- RayExtensions.IntersectsTriangleBackfaceCulling (RayExtensions.cs:42-75) ---------------------
def test_triangle_kat(oracle_mod):
    O = oracle_mod
    t = one_tri((-1, -1, 0), (-1, 1, 0), (1, -1, 0))
    assert np.allclose(t["surfaceNormal"][0], (0, 0, 1))
    r, u, v, d = O.tri_test(t[0], (-0.5, -0.5, 5, 0, 0, -1))
    assert (r, u, v, d) == (2, 0.25, 0.25, 5.0)
    # u + v == 1 is accepted (<=)
    r, u, v, d = O.tri_test(t[0], (0, 0, 5, 0, 0, -1))
    assert (r, u, v, d) == (2, 0.5, 0.5, 5.0)
    # swapped winding flips the precomputed normal -> culled by dot(n, D) > 0
    t2 = one_tri((-1, -1, 0), (1, -1, 0), (-1, 1, 0))
    assert O.tri_test(t2[0], (-0.5, -0.5, 5, 0, 0, -1))[0] == 0
    # behind the origin: d < 0 rejected; just outside: u < 0 rejected
    assert O.tri_test(t[0], (-0.5, -0.5, -5, 0, 0, -1))[0] == 1
    assert O.tri_test(t[0], (-1.5, -0.5, 5, 0, 0, -1))[0] == 1
    # d == 0 is accepted (no epsilon, Q4)
    r, u, v, d = O.tri_test(t[0], (-0.5, -0.5, 0, 0, 0, -1))
    assert r == 2 and d == 0.0
    # ray in the triangle's plane: det == 0 -> inv = inf, products are NaN/inf, comparisons fail
    assert O.tri_test(t[0], (-5, -0.5, 0, 1, 0, 0))[0] == 1


# ---- BoundingBox.Intersects(ref Ray, out float?) (XNA 4.0 form) ----------------------------------
def test_slab_kat(oracle_mod):
    O = oracle_mod
    box = (-1, -1, -1, 1, 1, 1)
    assert O.slab(box, (0, 0, 0, 0, 0, 1)) == 0.0                 # origin inside -> 0
    assert O.slab(box, (0, 0, -5, 0, 0, 1)) == 4.0                # entry distance
    assert O.slab(box, (0, 0, 5, 0, 0, 1)) is None                # box behind the ray
    assert O.slab(box, (2, 0, -5, 0, 0, 1)) is None               # parallel axis outside the slab
    assert O.slab(box, (1, 1, -5, 0, 0, 1)) == 4.0                # parallel axis on the face: inclusive
    assert O.slab(box, (0, 0, -5, 5e-7, 0, 1)) == 4.0             # |d| < 1e-6 counts as parallel
    d = xna.normalize(xna.v3(1, 1, 1))
    t = O.slab(box, (-3, -3, -3, *d))
    assert abs(t - 2 * math.sqrt(3)) < 1e-5
    # reciprocal-multiply (XNA) vs divide (MonoGame) may differ in the last ulp; both near 2*sqrt(3)
    tm = O.slab(box, (-3, -3, -3, *d), SEMANTICS_MONOGAME)
    assert abs(tm - 2 * math.sqrt(3)) < 1e-5
    # MonoGame: origin inside -> 0; behind -> None
    assert O.slab(box, (0, 0, 0, 0, 0, 1), SEMANTICS_MONOGAME) == 0.0
    assert O.slab(box, (0, 0, 5, 0, 0, 1), SEMANTICS_MONOGAME) is None
    # exact arithmetic check of the XNA form: t = (min - p) * (1/d), not (min - p)/d
    p, dz = F(-5.3), F(0.3)
    expect = F(F(F(-1) - p) * F(F(1) / dz))
    assert O.slab(box, (0, 0, p, 0, 0, dz)) == expect


# ---- Color(Vector3) / ToVector3 ------------------------------------------------------------------
def test_color_pack_kat(oracle_mod):
    O = oracle_mod
    ch = lambda v, s=SEMANTICS_XNA40: O.pack_color((v, 0, 0), s) & 255
    assert ch(0.5) == 128 and ch(0.5, SEMANTICS_MONOGAME) == 127   # 127.5 -> half-even 128 vs truncation
    assert O.round_channel_premultiplied(2.5) == 2 and O.round_channel_premultiplied(3.5) == 4
    assert ch(float("nan")) == 0 and ch(1.7) == 255 and ch(-0.1) == 0 and ch(float("inf")) == 255
    c = O.pack_color((1.0, 0.0, 0.2))
    assert c == (255 | (0 << 8) | (51 << 16) | (255 << 24))
    assert np.array_equal(O.unpack_color(c), np.array([1.0, 0.0, F(51) / F(255)], dtype=np.float32))
    assert O.pack_color((0, 0, 0)) == 0xFF000000  # background (RayTracer.cs:732, Q10)


# ---- Material.LookupUV ---------------------------------------------------------------------------
def test_texture_kat(oracle_mod, tmp_path):
    O = oracle_mod
    W = H = 8
    argb = (np.uint32(255) << 24) | (np.arange(W * H, dtype=np.uint32).reshape(H, W) * 3)
    rgb = lambda a: np.array([((a >> 16) & 255), ((a >> 8) & 255), (a & 255)], dtype=np.float32) * (F(1) / F(255))
    # Wrap: u == 1.0 is not > 1 -> stays 1.0 -> x = W-1 ; v = 0 -> y = 0
    assert np.array_equal(O.lookup_uv(argb, (1.0, 0.0)), rgb(int(argb[0, W - 1])))
    # u = -0.25 -> 1 + fmod(-0.25, 1) = 0.75 -> x = (int)(0.75*7) = 5
    assert np.array_equal(O.lookup_uv(argb, (-0.25, 0.5)), rgb(int(argb[3, 5])))
    # u = 2.5 -> fmod -> 0.5 -> x = 3 ; v = 1.25 -> 0.25 -> y = 1
    assert np.array_equal(O.lookup_uv(argb, (2.5, 1.25)), rgb(int(argb[1, 3])))
    # Clamp
    assert np.array_equal(O.lookup_uv(argb, (7.0, -3.0), ADDRESS_CLAMP), rgb(int(argb[0, W - 1])))
    # Mirror (Material.cs:102-123): u = 1.25 -> wrapped 0.25; (int)(1.25-0.25) % 2 = 1 -> not flipped
    assert np.array_equal(O.lookup_uv(argb, (1.25, 0.0), ADDRESS_MIRROR), rgb(int(argb[H - 1, 1])))
    # BMP decode: bottom-up file -> top-down texels (GDI+ LockBits order)
    tex = synthetic_crate_texture(64)
    p = tmp_path / "t.bmp"
    write_bmp24(str(p), tex)
    assert np.array_equal(read_bmp_argb(str(p)), tex)
    # bilinear path runs and stays within [0, 1]
    c = O.lookup_uv(tex, (0.37, 0.61), ADDRESS_WRAP, FILTER_BILINEAR)
    assert np.all(c >= 0) and np.all(c <= 1.01)


def test_reference_bmp_if_present():
    import os

    path = "/root/reference/Free_crate/Diffuse.bmp"
    if not os.path.exists(path):
        pytest.skip("reference assets are not on this box")
    tex = read_bmp_argb(path)
    assert tex.shape == (512, 512) and (tex >> 24 == 255).all()


# ---- framework matrix math: numpy host mirror == C++ oracle, and hand-checked values -----------------
def test_matrix_kat(oracle_mod):
    O = oracle_mod
    eye, tgt, up = (0.0, 16.0, 32.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0)
    V = xna.create_look_at(eye, tgt, up)
    assert np.array_equal(V, O.create_look_at(eye, tgt, up))
    # camera looks down -z in view space: the target maps onto the -z axis at distance |eye|
    tv = xna.transform(xna.v3(*tgt), V)
    assert abs(tv[0]) < 1e-5 and abs(tv[1]) < 1e-5 and abs(tv[2] + math.hypot(16, 32)) < 1e-4
    P = xna.create_perspective_fov(math.pi / 4, xna.aspect_ratio(640, 480), 1.0, 1000.0)
    assert np.array_equal(P, O.create_perspective_fov(F(math.pi / 4), xna.aspect_ratio(640, 480), 1.0, 1000.0))
    assert P[2, 3] == -1 and P[3, 3] == 0 and abs(P[1, 1] - 1 / math.tan(math.pi / 8)) < 1e-6
    VP = xna.multiply(V, P)
    assert np.array_equal(VP, O.matrix_multiply(V, P))
    inv = xna.invert(VP)
    assert np.array_equal(inv, O.matrix_invert(VP))
    assert np.allclose(xna.multiply(VP, inv), np.eye(4), atol=2e-3)
    # instance transform: S * Rx*Ry*Rz * T
    o = scenes.ObjectData([0], position=(1, 2, 3), rotation=(0.3, -0.4, 1.1), scale=(2, 1, 0.5))
    o.build_world()
    assert np.array_equal(o.inverse_world, O.matrix_invert(o.world))
    assert np.allclose(xna.transform(xna.v3(0, 0, 0), o.world), (1, 2, 3))


def test_primary_ray_kat(oracle_mod):
    """Viewport.Unproject: the ray of the centre pixel passes through the camera target; pixel
    coordinates are integers (no +0.5, RayTracer.cs:412-413)."""
    O = oracle_mod
    s = scenes.crate_scene()
    f = s.frame(640, 480)
    rays = O.primary_rays(f).reshape(480, 640, 6)
    eye = np.array(s.camera.position, dtype=np.float64)
    tgt = np.array(s.camera.target, dtype=np.float64)
    r = rays[240, 320].astype(np.float64)
    want = (tgt - eye) / np.linalg.norm(tgt - eye)
    assert np.allclose(r[3:], want, atol=1e-4)
    # origin lies on the near plane (distance 1 along the view axis)
    assert abs(np.dot(r[:3] - eye, want) - 1.0) < 1e-3
    assert np.allclose(np.linalg.norm(rays[..., 3:], axis=-1), 1.0, atol=1e-5)


# ---- lights ------------------------------------------------------------------------------------------
def test_light_kat(oracle_mod):
    O = oracle_mod
    L = scenes.spot_light((0.0, 10.0, 0.0), direction=(0.0, -1.0, 0.0), spot_angle=math.pi / 2, decay=1.3)
    # fragment straight below the light, normal up: surfaceDot = 1, lightDot = 1
    c = O.light_for_fragment(L, (0, 0, 0), (0, 1, 0))
    ang = F(math.cos(float(F(F(math.pi / 2) * F(0.5)))))
    spot = F(F(1.0) * F(float(F(1.0) - ang) / math.pow(float(F(1.0) - ang), float(F(1.3)))))
    want = F(F(F(1.0) * spot) * F(1.0)) + F(1.0)
    assert np.array_equal(c, np.array([want] * 3, dtype=np.float32))
    # outside the cone -> zero; facing away -> zero
    assert not O.light_for_fragment(L, (20, 0, 0), (0, 1, 0)).any()
    assert not O.light_for_fragment(L, (0, 0, 0), (0, -1, 0)).any()
    # directional uses +Direction (not -Direction), DirectionalLight.cs:25
    D = scenes.directional_light((0, 1, 0), color=(0.5, 1, 1), intensity=2.0)
    assert np.array_equal(O.light_for_fragment(D, (0, 0, 0), (0, 1, 0)), np.array([1, 2, 2], dtype=np.float32))
    assert not O.light_for_fragment(D, (0, 0, 0), (0, -1, 0)).any()


# ---- MeshOctree.Build / SplitCuboid ----------------------------------------------------------------
def _scatter_tris(centers, size=0.01):
    c = np.asarray(centers, dtype=np.float32)
    return scenes.make_triangles(c, c + np.array([0, size, 0], np.float32), c + np.array([size, 0, 0], np.float32))


def test_octree_split_kat(oracle_mod):
    O = oracle_mod
    rng = np.random.default_rng(0)
    # 50 triangles: no split; 51: exactly one split into 8 children
    for n, want_nodes in ((50, 1), (51, 9)):
        o = O.Oracle()
        tris = _scatter_tris(rng.uniform(-1, 1, size=(n, 3)))
        o.add_mesh(tris, box=scenes.mesh_bounding_box(tris))
        tr = o.tree(0)
        assert len(tr["boxes"]) == want_nodes
    # child index = 4i+2j+k <-> (x, y, z) halves, box = (Min + size*(i,j,k), +size)
    root = tr["boxes"][0]
    size = (root[3:] - root[:3]) / F(2)
    for i in range(2):
        for j in range(2):
            for k in range(2):
                b = tr["boxes"][1 + 4 * i + 2 * j + k]
                mn = root[:3] + size * np.array([i, j, k], dtype=np.float32)
                assert np.array_equal(b[:3], mn) and np.array_equal(b[3:], mn + size)
    # every triangle lands in each child containing one of its vertices (inclusive): a vertex on
    # the split plane is in both neighbours
    o = O.Oracle()
    cs = np.concatenate([rng.uniform(-1, 1, size=(60, 3)), [[-1, -1, -1], [1, 1, 1]]]).astype(np.float32)
    tris = _scatter_tris(cs, size=0.0)  # degenerate: all three vertices coincide
    tris["v1"][0] = tris["v2"][0] = tris["v3"][0] = (0, 0, 0)  # exactly the centre of [-1,1]^3
    o.add_mesh(tris, box=scenes.mesh_bounding_box(tris))
    tr = o.tree(0)
    assert np.array_equal(tr["boxes"][0], np.array([-1, -1, -1, 1, 1, 1], dtype=np.float32))
    holders = [c for c in range(1, 9) if 0 in tr["refs"][tr["ref_start"][c]: tr["ref_start"][c] + tr["ref_count"][c]]]
    assert holders == list(range(1, 9))


def test_octree_depth_guard(oracle_mod):
    """>threshold triangles sharing one vertex recurse forever in the reference (MeshOctree.cs:84-96);
    the native builders report an error instead."""
    O = oracle_mod
    o = O.Oracle()
    tris = _scatter_tris(np.zeros((60, 3)), size=0.0)
    with pytest.raises(O.OracleError) as e:
        o.add_mesh(tris, box=scenes.mesh_bounding_box(tris))
    assert e.value.code == 5


# ---- traversal quirks ----------------------------------------------------------------------------
def _facing(tri, direction):
    """Re-wind a one-triangle record so that it faces a ray travelling along `direction`."""
    if float(np.dot(tri["surfaceNormal"][0], direction)) > 0:
        tri = scenes.make_triangles(tri["v1"], tri["v3"], tri["v2"])
    return tri


def early_exit_mesh():
    """Root [-1,1]^3 split once.  A ray along +z at x=y=-0.5 enters child 0 (z<0) at t=1 and child 1
    (z>0) at t=2.  Triangle B has a vertex in child 0 (so it is stored there) and is hit at d=2.5;
    triangle A lies only in z>0 leaves and is hit at d=2.2."""
    rng = np.random.default_rng(1)
    fill = []
    for sx, sy, sz in ((-1, 1, -1), (1, -1, -1), (1, 1, -1), (1, 1, 1)):  # children 2, 4, 6, 7
        c = rng.uniform(0.3, 0.9, size=(20, 3)) * np.array([sx, sy, sz])
        fill.append(c)
    fill = _scatter_tris(np.concatenate(fill))
    corners = _scatter_tris([[-1, -1, -1], [1, 1, 1]], size=0.0)
    d = np.array([0, 0, 1], dtype=np.float32)
    B = _facing(scenes.make_triangles([(-0.9, -0.9, -0.9)], [(-0.9, 0.1, 0.85)], [(0.1, -0.9, 0.85)]), d)
    A = _facing(scenes.make_triangles([(-0.9, -0.9, 0.2)], [(-0.9, 0.5, 0.2)], [(0.5, -0.9, 0.2)]), d)
    allt = np.concatenate([B, A, fill, corners])
    allt["id"] = np.arange(len(allt))
    return allt


def test_early_exit_by_leaf_entry_quirk(oracle_mod):
    """Q1: the walk stops after the first leaf GROUP (by slab entry distance) that yields a hit, so a
    nearer triangle stored only in a later leaf is not found (MeshOctree.cs:281,300-301)."""
    O = oracle_mod
    allt = early_exit_mesh()
    ray = np.array([[-0.5, -0.5, -2.0, 0, 0, 1]], dtype=np.float32)
    assert O.tri_test(allt[0], ray[0])[0] == 2 and abs(O.tri_test(allt[0], ray[0])[3] - 2.5) < 1e-5
    assert O.tri_test(allt[1], ray[0])[0] == 2 and abs(O.tri_test(allt[1], ray[0])[3] - 2.2) < 1e-5
    for mode in (0, 1):
        o = O.Oracle(mode=mode)
        o.add_mesh(allt, box=scenes.mesh_bounding_box(allt))
        o.add_object([0], np.eye(4), np.eye(4), scenes.mesh_bounding_box(allt))
        o.build_scene()
        tr = o.tree(0)
        assert len(tr["boxes"]) == 9
        leaf = lambda c: set(tr["refs"][tr["ref_start"][c]: tr["ref_start"][c] + tr["ref_count"][c]].tolist())
        assert 0 in leaf(1 + 0) and 1 not in leaf(1 + 0) and 1 in leaf(1 + 1)
        r = o.trace_rays(ray)
        assert r["tri"][0] == 0, "reference behaviour: the farther triangle B (stored in the nearer leaf) wins"
        assert abs(r["d"][0] - 2.5) < 1e-5


def test_shared_mesh_ignore_quirk(oracle_mod):
    """Q6: ignoreTriangle is the identity of a Triangle object shared by every instance of a model,
    so a ray leaving triangle k of instance A also skips triangle k of instance B."""
    O = oracle_mod
    quad = scenes.make_triangles([(-1, -1, 0)], [(-1, 1, 0)], [(1, -1, 0)])  # faces +z
    for mode in (0, 1):
        o = O.Oracle(mode=mode)
        o.add_mesh(quad, box=scenes.mesh_bounding_box(quad))
        a = scenes.ObjectData([0], position=(0, 0, 0)); a.build_world()
        b = scenes.ObjectData([0], position=(0, 0, -3)); b.build_world()
        box = scenes.mesh_bounding_box(quad)
        o.add_object([0], a.world, a.inverse_world, box)
        o.add_object([0], b.world, b.inverse_world, box)
        o.build_scene()
        ray = np.array([[-0.5, -0.5, 5, 0, 0, -1]], dtype=np.float32)
        r = o.trace_rays(ray)
        assert (r["obj"][0], r["tri"][0], r["d"][0]) == (0, 0, 5.0)
        # ignoring (mesh 0, tri 0) hides it in BOTH instances
        r = o.trace_rays(ray, ignore=[[0, 0]])
        assert r["obj"][0] == -1
        # a different ignore key hides nothing
        r = o.trace_rays(ray, ignore=[[0, 7]])
        assert r["obj"][0] == 0


def test_object_space_distance_quirk(oracle_mod):
    """Q7: d is the object-space distance; a scaled instance reports d / scale."""
    O = oracle_mod
    quad = scenes.make_triangles([(-1, -1, 0)], [(-1, 1, 0)], [(1, -1, 0)])
    o = O.Oracle()
    o.add_mesh(quad, box=scenes.mesh_bounding_box(quad))
    a = scenes.ObjectData([0], scale=(2, 2, 2)); a.build_world()
    o.add_object([0], a.world, a.inverse_world, scenes.mesh_bounding_box(quad))
    o.build_scene()
    r = o.trace_rays(np.array([[-0.5, -0.5, 5, 0, 0, -1]], dtype=np.float32))
    assert r["tri"][0] == 0 and abs(r["d"][0] - 2.5) < 1e-6
    assert np.allclose(r["pos"][0], (-0.5, -0.5, 0), atol=1e-6)


# ---- literal reference algorithm == pruned order-independent restatement ---------------------------
@pytest.mark.parametrize("name", ["spheres", "grid", "many", "soup", "crate"])
def test_literal_equals_pruned(oracle_mod, name):
    O = oracle_mod
    s = {"spheres": scenes.spheres_scene, "grid": lambda: scenes.grid_scene(48),
         "many": scenes.many_objects_scene, "soup": lambda: scenes.soup_scene(30000, half=10.0),
         "crate": scenes.crate_scene}[name]()
    rng = np.random.default_rng(5)
    n = 4000
    # rays: primary rays of the scene camera + random rays starting inside the scene bounds
    f = s.frame(64, 48)
    prim = O.primary_rays(f)
    org = rng.uniform(-12, 12, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    d[::7, 0] = 0.0  # axis-parallel components exercise the |d| < 1e-6 branch
    d[::11, 1] = 0.0
    rays = np.concatenate([prim, np.concatenate([org, d], axis=1)]).astype(np.float32)
    out = []
    for mode in (0, 1):
        o = O.Oracle(mode=mode)
        scenes.upload(o, s)
        out.append(o.trace_rays(rays))
    a, b = out
    assert (a["tri"] >= 0).sum() > 50
    for k in ("obj", "mesh", "tri"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("u", "v", "d", "pos"):
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
    # and whole frames (shadow, reflection and refraction rays start ON surfaces)
    imgs = []
    for mode in (0, 1):
        o = O.Oracle(mode=mode)
        scenes.upload(o, s)
        imgs.append(o.render_frame(f, n_threads=4))
    assert np.array_equal(imgs[0], imgs[1])


def test_render_properties(oracle_mod):
    O = oracle_mod
    s = scenes.crate_scene()
    o = O.Oracle()
    scenes.upload(o, s)
    f = s.frame(96, 72)
    img = o.render_frame(f)
    assert (img >> 24 == 255).all()
    c = o.counters()
    assert c.raysPrimary == 96 * 72 and c.pixels == 96 * 72
    assert c.raysShadow == c.hits * len(s.lights)
    # depth 0: no reflection rays at all
    f0 = s.frame(96, 72, max_reflections=0)
    o.render_frame(f0)
    assert o.counters().raysReflection == 0
    # multi-threaded scanline dispenser gives the same image
    assert np.array_equal(img, o.render_frame(f, n_threads=4))
    # row sharding renders only the shard's bands
    fs = s.frame(96, 72, shard_index=1, shard_count=2, shard_band_rows=8)
    part = o.render_frame(fs)
    bands = (np.arange(72) // 8) % 2 == 1
    assert np.array_equal(part[bands], img[bands]) and not part[~bands].any()
