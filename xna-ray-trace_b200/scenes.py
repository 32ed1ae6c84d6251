"""Synthetic scenes for the trace path, built with the triangle-record
conventions of the reference's content processor (RayTracePipeline/TracerModelProcessor.cs:133-307):

* `surfaceNormal = Normalize(Cross(edge2, edge1))`, edge1 = v2-v1, edge2 = v3-v1 (:199-203) — the
  front face is the side from which v1,v2,v3 appear clockwise;
* `Mesh.MeshBoundingBox` grows from `new BoundingBox()` = (0,0,0)-(0,0,0) (:246), so it always
  contains the origin;
* triangle `id` is the running index inside the mesh (:167,221), `color` = DiffuseColor (:227).

Scenes are plain data (`SceneData`); `upload(api, scene)` feeds them to any object exposing the
xnart C-ABI surface (the libxnart binding, or — in tests — the oracle binding).
"""
import math
from dataclasses import dataclass, field

import numpy as np

from . import xna
from .records import (ADDRESS_WRAP, FILTER_POINT, LIGHT_DIRECTIONAL, LIGHT_SPOT, SEMANTICS_XNA40,
                      TRIANGLE_DTYPE, make_frame)

F = np.float32


# ----------------------------------------------------------------------------------------
# vectorised strict-float32 helpers (N x 3 arrays)
# ----------------------------------------------------------------------------------------
def _cross(a, b):
    return np.stack(
        [a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2], a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]],
        axis=1,
    ).astype(np.float32)


def _normalize(v):
    s = (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        f = (F(1.0) / np.sqrt(s.astype(np.float64)).astype(np.float32)).astype(np.float32)
        return (v * f[:, None]).astype(np.float32)


def make_triangles(v1, v2, v3, n1=None, n2=None, n3=None, uv1=None, uv2=None, uv3=None, color=(1, 1, 1, 1)):
    """Triangle records following TracerModelProcessor.FindVertices (:168-229)."""
    v1, v2, v3 = (np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3) for a in (v1, v2, v3))
    n = v1.shape[0]
    t = np.zeros(n, dtype=TRIANGLE_DTYPE)
    t["v1"], t["v2"], t["v3"] = v1, v2, v3
    e1 = (v2 - v1).astype(np.float32)
    e2 = (v3 - v1).astype(np.float32)
    sn = _normalize(_cross(e2, e1))
    t["surfaceNormal"] = sn
    t["n1"] = sn if n1 is None else _normalize(np.asarray(n1, dtype=np.float32).reshape(-1, 3))
    t["n2"] = sn if n2 is None else _normalize(np.asarray(n2, dtype=np.float32).reshape(-1, 3))
    t["n3"] = sn if n3 is None else _normalize(np.asarray(n3, dtype=np.float32).reshape(-1, 3))
    if uv1 is not None:
        t["uv1"], t["uv2"], t["uv3"] = uv1, uv2, uv3
    t["color"] = np.asarray(color, dtype=np.float32)
    t["id"] = np.arange(n, dtype=np.int32)
    return t


def mesh_bounding_box(tris):
    """TracerModelProcessor.CreateBoundingBox (:244-307): starts from (0,0,0)-(0,0,0)."""
    pts = np.concatenate([tris["v1"], tris["v2"], tris["v3"]], axis=0) if len(tris) else np.zeros((1, 3), np.float32)
    mn = np.minimum(pts.min(axis=0), 0).astype(np.float32)
    mx = np.maximum(pts.max(axis=0), 0).astype(np.float32)
    return np.concatenate([mn, mx]).astype(np.float32)


# ----------------------------------------------------------------------------------------
# scene description
# ----------------------------------------------------------------------------------------
@dataclass
class MeshData:
    tris: np.ndarray
    reflectiveness: float = 0.0
    use_texture: bool = False
    interpolate_normals: bool = True  # TracerModelProcessor.cs:30
    transparent: bool = False
    refraction_index: float = 0.0
    argb: np.ndarray = None
    box: np.ndarray = None
    item_threshold: int = 50  # MeshOctree.cs:42

    def __post_init__(self):
        if self.box is None:
            self.box = mesh_bounding_box(self.tris)


@dataclass
class ObjectData:
    mesh_ids: list
    position: tuple = (0.0, 0.0, 0.0)
    rotation: tuple = (0.0, 0.0, 0.0)
    scale: tuple = (1.0, 1.0, 1.0)
    world: np.ndarray = None
    inverse_world: np.ndarray = None
    local_box: np.ndarray = None

    def build_world(self):
        """SceneObject.BuildWorld (SceneObject.cs:183-199): World = S * (Rx*Ry*Rz) * T."""
        s = xna.create_scale(self.scale)
        r = xna.multiply(xna.multiply(xna.create_rotation_x(self.rotation[0]), xna.create_rotation_y(self.rotation[1])),
                         xna.create_rotation_z(self.rotation[2]))
        t = xna.create_translation(self.position)
        self.world = xna.multiply(xna.multiply(s, r), t)
        self.inverse_world = xna.invert(self.world)


@dataclass
class CameraData:
    position: tuple
    target: tuple
    up: tuple = (0.0, 1.0, 0.0)
    fov: float = math.pi / 4  # MathHelper.PiOver4 (Game1.cs:111)
    near: float = 1.0
    far: float = 1000.0

    def view(self):
        return xna.create_look_at(self.position, self.target, self.up)  # Camera.cs:40-47

    def projection(self, width, height):
        return xna.create_perspective_fov(F(self.fov), xna.aspect_ratio(width, height), self.near, self.far)  # Camera.cs:49-54


@dataclass
class SceneData:
    name: str
    meshes: list = field(default_factory=list)
    objects: list = field(default_factory=list)
    lights: list = field(default_factory=list)
    camera: CameraData = None
    max_reflections: int = 8  # Game1.cs:126
    scene_item_threshold: int = 20  # OctreeSpatialManager.cs:50
    filtering: int = FILTER_POINT  # Game1.cs:120
    address_mode: int = ADDRESS_WRAP  # Game1.cs:121

    def finalize(self):
        for o in self.objects:
            if o.world is None:
                o.build_world()
            if o.local_box is None:
                # SceneObject ctor (SceneObject.cs:129-131): CreateMerged(default box, mesh boxes)
                mn, mx = np.zeros(3, np.float32), np.zeros(3, np.float32)
                for mid in o.mesh_ids:
                    mn = np.minimum(mn, self.meshes[mid].box[:3])
                    mx = np.maximum(mx, self.meshes[mid].box[3:])
                o.local_box = np.concatenate([mn, mx]).astype(np.float32)
        return self

    def frame(self, width, height, camera=None, **kw):
        cam = camera or self.camera
        kw.setdefault("max_reflections", self.max_reflections)
        kw.setdefault("filtering", self.filtering)
        kw.setdefault("address_mode", self.address_mode)
        kw.setdefault("semantics", SEMANTICS_XNA40)
        return make_frame(cam.view(), cam.projection(width, height), width, height, **kw)

    @property
    def n_triangles(self):
        return sum(len(m.tris) for m in self.meshes)


def upload(api, scene):
    """Feed a SceneData to an xnart-shaped API object (add_mesh/add_object/build_scene/set_lights)."""
    scene.finalize()
    for m in scene.meshes:
        api.add_mesh(m.tris, reflectiveness=m.reflectiveness, use_texture=m.use_texture,
                     interpolate_normals=m.interpolate_normals, transparent=m.transparent,
                     refraction_index=m.refraction_index, argb=m.argb, box=m.box, item_threshold=m.item_threshold)
    for o in scene.objects:
        api.add_object(o.mesh_ids, o.world, o.inverse_world, o.local_box)
    api.build_scene(scene.scene_item_threshold)
    api.set_lights(scene.lights)
    return api


def spot_light(position, direction=None, color=(1, 1, 1), spot_angle=math.pi / 2, intensity=1.0, decay=1.3):
    """SpotLight as configured in Game1.cs:132-138 (direction defaults to -Normalize(position))."""
    pos = np.asarray(position, dtype=np.float32)
    if direction is None:
        direction = -xna.normalize(pos)
    return dict(kind=LIGHT_SPOT, position=pos, direction=np.asarray(direction, np.float32), color=color,
                intensity=intensity, decayExponent=decay, spotAngle=F(spot_angle))


def directional_light(direction, color=(1, 1, 1), intensity=1.0):
    return dict(kind=LIGHT_DIRECTIONAL, direction=np.asarray(direction, np.float32), color=color, intensity=intensity,
                position=(0, 0, 0), decayExponent=0.0, spotAngle=0.0)


# ----------------------------------------------------------------------------------------
# geometry
# ----------------------------------------------------------------------------------------
def _orient_outward(v1, v2, v3, center):
    """Swap v2/v3 where needed so that Normalize(Cross(e2,e1)) points away from `center`."""
    sn = _cross((v3 - v1).astype(np.float32), (v2 - v1).astype(np.float32))
    cen = ((v1 + v2 + v3) / F(3.0)) - np.asarray(center, dtype=np.float32)
    flip = (sn * cen).sum(axis=1) < 0
    v2n = np.where(flip[:, None], v3, v2)
    v3n = np.where(flip[:, None], v2, v3)
    return v1, v2n.astype(np.float32), v3n.astype(np.float32), flip


def _crate_from(verts64, quads, quv):
    """The crate mesh from FBX control points (Z up), quads and per-corner UVs: content scale 0.6
    (RayTraceProjectContent.contentproj:69), Z-up -> Y-up, fan triangulation with the winding reversed."""
    verts = np.asarray(verts64, dtype=np.float64) * 0.6
    verts = np.stack([verts[:, 0], verts[:, 2], -verts[:, 1]], axis=1).astype(np.float32)  # (x, z, -y)
    v1, v2, v3, u1, u2, u3 = [], [], [], [], [], []
    for qi, q in enumerate(quads):
        uvq = quv[qi] if np.ndim(quv) == 3 else quv
        for a, b, c in ((0, 1, 2), (0, 2, 3)):
            # reversed winding so that Cross(e2, e1) is the outward normal
            v1.append(verts[q[a]]); v2.append(verts[q[c]]); v3.append(verts[q[b]])
            u1.append(uvq[a]); u2.append(uvq[c]); u3.append(uvq[b])
    v1, v2, v3 = (np.array(a, dtype=np.float32) for a in (v1, v2, v3))
    center = verts.mean(axis=0)
    sn = _normalize(_cross((v3 - v1).astype(np.float32), (v2 - v1).astype(np.float32)))
    cen = ((v1 + v2 + v3) / F(3.0)) - center
    assert ((sn * cen).sum(axis=1) > 0).all(), "crate winding must face outward"
    return make_triangles(v1, v2, v3, uv1=np.array(u1, dtype=np.float32), uv2=np.array(u2, dtype=np.float32),
                          uv3=np.array(u3, dtype=np.float32), color=(1, 1, 1, 1))


def crate_triangles():
    """Crate_Fragile.FBX geometry: 8 vertices, 6 quads, scale 0.6, Z-up -> Y-up.  The asset does not
    travel to the GPU box, so its numbers are restated here; tests/test_ingest.py checks them against the file."""
    h, H = 19.685039520263672, 39.370079040527344
    verts = [(-h, -h, 0), (-h, -h, H), (h, -h, H), (h, -h, 0), (h, h, 0), (h, h, H), (-h, h, H), (-h, h, 0)]
    quads = [(3, 2, 1, 0), (7, 6, 5, 4), (0, 1, 6, 7), (4, 5, 2, 3), (2, 5, 6, 1), (4, 3, 0, 7)]
    quv = np.array([(1, 0), (1, 1), (0, 1), (0, 0)], dtype=np.float32)
    return _crate_from(verts, quads, quv)


def crate_triangles_from_fbx(path):
    """The same mesh with control points, polygons and UVs read from the shipped binary asset
    (RayTraceProjectContent/Crate_Fragile.FBX, ingest.read_fbx_meshes)."""
    from . import ingest

    (m,) = ingest.read_fbx_meshes(path)
    mapping, ref, data, index = m["uvs"]
    assert mapping == "ByPolygonVertex" and all(len(q) == 4 for q in m["polygons"])
    pv, quv = 0, []
    for q in m["polygons"]:
        quv.append([data[int(index[pv + k])] if index is not None else data[pv + k] for k in range(4)])
        pv += 4
    return _crate_from(m["vertices"], m["polygons"], np.asarray(quv, dtype=np.float32))


def ground_triangles(half=60.0, y=0.0, color=(1, 1, 1, 1)):
    p00, p10, p01, p11 = (-half, y, -half), (half, y, -half), (-half, y, half), (half, y, half)
    v1 = np.array([p00, p10], dtype=np.float32)
    v2 = np.array([p10, p11], dtype=np.float32)
    v3 = np.array([p01, p01], dtype=np.float32)
    uv1 = np.array([(0, 0), (1, 0)], dtype=np.float32)
    uv2 = np.array([(1, 0), (1, 1)], dtype=np.float32)
    uv3 = np.array([(0, 1), (0, 1)], dtype=np.float32)
    t = make_triangles(v1, v2, v3, uv1=uv1, uv2=uv2, uv3=uv3, color=color)
    assert (t["surfaceNormal"][:, 1] > 0.99).all()
    return t


def uv_sphere_triangles(radius=1.0, segments=32, rings=16, color=(1, 1, 1, 1)):
    """UV sphere with the topology of the shipped Sphere.fbx (482 vertices / 960 triangles)."""
    v1, v2, v3 = [], [], []

    def p(ring, seg):
        th = math.pi * ring / rings
        ph = 2.0 * math.pi * (seg % segments) / segments
        return (radius * math.sin(th) * math.cos(ph), radius * math.cos(th), radius * math.sin(th) * math.sin(ph))

    for r in range(rings):
        for s in range(segments):
            a, b, c, d = p(r, s), p(r, s + 1), p(r + 1, s), p(r + 1, s + 1)
            if r != 0:
                v1.append(a); v2.append(b); v3.append(c)
            if r != rings - 1:
                v1.append(b); v2.append(d); v3.append(c)
    v1, v2, v3 = (np.array(a, dtype=np.float32) for a in (v1, v2, v3))
    v1, v2, v3, _ = _orient_outward(v1, v2, v3, (0, 0, 0))
    return make_triangles(v1, v2, v3, n1=v1, n2=v2, n3=v3, color=color)


def height(x, z):
    return 4.0 * (np.sin(0.13 * x) * np.cos(0.11 * z) + 0.5 * np.sin(0.37 * x + 0.2) * np.sin(0.29 * z)
                  + 0.25 * np.cos(0.71 * x - 0.53 * z))


def grid_triangles(n_quads=707, half=50.0, seed=1, color=(0.8, 0.8, 0.8, 1.0)):
    """Displaced grid of n_quads^2 quads on [-half, half]^2 (C3: 707 -> 999,698 triangles)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = n_quads + 1
    xs = np.linspace(-half, half, n)
    X, Z = np.meshgrid(xs, xs, indexing="ij")
    Y = height(X, Z) + rng.uniform(-0.01, 0.01, size=X.shape)
    P = np.stack([X, Y, Z], axis=-1).astype(np.float32)
    # analytic-ish vertex normals from central differences of the displaced surface
    gx = np.gradient(Y, xs, axis=0)
    gz = np.gradient(Y, xs, axis=1)
    N = np.stack([-gx, np.ones_like(gx), -gz], axis=-1).astype(np.float32)
    p00, p10, p01, p11 = P[:-1, :-1], P[1:, :-1], P[:-1, 1:], P[1:, 1:]
    n00, n10, n01, n11 = N[:-1, :-1], N[1:, :-1], N[:-1, 1:], N[1:, 1:]
    v1 = np.stack([p00, p10], axis=2).reshape(-1, 3)
    v2 = np.stack([p10, p11], axis=2).reshape(-1, 3)
    v3 = np.stack([p01, p01], axis=2).reshape(-1, 3)
    m1 = np.stack([n00, n10], axis=2).reshape(-1, 3)
    m2 = np.stack([n10, n11], axis=2).reshape(-1, 3)
    m3 = np.stack([n01, n01], axis=2).reshape(-1, 3)
    return make_triangles(v1, v2, v3, n1=m1, n2=m2, n3=m3, color=color)


def soup_triangles(n=10_000_000, half=100.0, edge=0.5, seed=2, color=(0.9, 0.9, 0.9, 1.0)):
    """Random triangle soup (C5): centroid ~ U[-half,half]^3, two edge vectors ~ U[-edge,edge]^3."""
    rng = np.random.Generator(np.random.PCG64(seed))
    c = rng.uniform(-half, half, size=(n, 3)).astype(np.float32)
    e1 = rng.uniform(-edge, edge, size=(n, 3)).astype(np.float32)
    e2 = rng.uniform(-edge, edge, size=(n, 3)).astype(np.float32)
    return make_triangles(c, (c + e1).astype(np.float32), (c + e2).astype(np.float32), color=color)


def cube_triangles(size=1.0, color=(1, 1, 1, 1)):
    s = size / 2.0
    verts = np.array([(-s, -s, -s), (s, -s, -s), (s, s, -s), (-s, s, -s), (-s, -s, s), (s, -s, s), (s, s, s), (-s, s, s)],
                     dtype=np.float32)
    quads = [(0, 1, 2, 3), (5, 4, 7, 6), (4, 0, 3, 7), (1, 5, 6, 2), (3, 2, 6, 7), (4, 5, 1, 0)]
    v1, v2, v3 = [], [], []
    for q in quads:
        for a, b, c in ((0, 1, 2), (0, 2, 3)):
            v1.append(verts[q[a]]); v2.append(verts[q[b]]); v3.append(verts[q[c]])
    v1, v2, v3 = (np.array(a, dtype=np.float32) for a in (v1, v2, v3))
    v1, v2, v3, _ = _orient_outward(v1, v2, v3, (0, 0, 0))
    return make_triangles(v1, v2, v3, color=color)


# ----------------------------------------------------------------------------------------
# the configurations of BASELINE.json
# ----------------------------------------------------------------------------------------
def crate_scene(argb=None):
    """C1/C2: textured Free_crate box on a reflective ground quad, one spot light, depth 8."""
    from .bmp import synthetic_crate_texture

    if argb is None:
        argb = synthetic_crate_texture()
    s = SceneData("free_crate")
    s.meshes.append(MeshData(crate_triangles(), reflectiveness=0.5, use_texture=True, interpolate_normals=True, argb=argb))
    s.meshes.append(MeshData(ground_triangles(), reflectiveness=0.5, use_texture=False, interpolate_normals=False))
    s.objects.append(ObjectData([0]))
    s.objects.append(ObjectData([1]))
    s.lights.append(spot_light((20.0, 60.0, 50.0)))
    s.camera = CameraData((40.0, 35.0, 60.0), (0.0, 11.8, 0.0))
    s.max_reflections = 8
    return s.finalize()


def orbit_camera(frame_index, n_frames=600, radius=90.0, height_y=30.0, target=(0.0, 0.0, 0.0)):
    """Flythrough camera in the style of Game1.cs:261,266 (C4)."""
    rot = 2.0 * math.pi * frame_index / n_frames
    return CameraData((math.sin(rot) * radius, height_y, math.cos(rot) * radius), target)


def grid_scene(n_quads=707, max_reflections=4):
    """C3/C4: ~1M-triangle displaced grid in one MeshOctree, spot + directional light."""
    s = SceneData("grid_%d" % n_quads)
    s.meshes.append(MeshData(grid_triangles(n_quads), reflectiveness=0.5, interpolate_normals=True))
    s.objects.append(ObjectData([0]))
    s.lights.append(spot_light((0.0, 80.0, 40.0)))
    s.lights.append(directional_light(xna.normalize(xna.v3(0.3, 1.0, 0.2))))
    s.camera = CameraData((0.0, 40.0, 80.0), (0.0, 0.0, 0.0))
    s.max_reflections = max_reflections
    return s.finalize()


def soup_scene(n=10_000_000, max_reflections=8, half=100.0):
    """C5: incoherent-ray stress — mirror triangle soup, one directional light."""
    s = SceneData("soup_%d" % n)
    s.meshes.append(MeshData(soup_triangles(n, half=half), reflectiveness=1.0, interpolate_normals=False))
    s.objects.append(ObjectData([0]))
    s.lights.append(directional_light(xna.normalize(xna.v3(0.2, 0.5, 1.0))))
    s.camera = CameraData((0.0, 0.0, 2.5 * half), (0.0, 0.0, 0.0))
    s.max_reflections = max_reflections
    return s.finalize()


def forking_quads_scene(depth=6):
    """Five parallel transparent quads, three facing the camera and two facing away (invisible to primary rays: back
    faces), refraction index 1: every hit forks into a reflection that turns round and a refraction that goes straight
    on, both towards more quads — the binary path tree of RayTracer.cs:586-702 at nearly full width (48 rays per pixel
    at depth 6).  A stress case for the path arena of transparent scenes."""
    def quad(z, flip):
        a, b, c = [(-5, -5, z), (5, 5, z)], [(-5, 5, z), (5, -5, z)], [(5, -5, z), (-5, 5, z)]
        if flip:
            b, c = c, b
        return make_triangles(a, b, c, color=(0.6, 0.8, 1.0, 0.5))

    s = SceneData("forking_quads")
    for z, flip in ((0.0, False), (-2.0, False), (-4.0, False), (2.0, True), (4.0, True)):
        s.meshes.append(MeshData(quad(z, flip), reflectiveness=0.5, transparent=True, refraction_index=1.0))
        s.objects.append(ObjectData([len(s.meshes) - 1]))
    s.lights.append(spot_light((0.0, 3.0, 12.0)))
    s.camera = CameraData((0.3, 0.2, 10.0), (0.0, 0.0, 0.0))
    s.max_reflections = depth
    return s.finalize()


def spheres_scene():
    """The shipped default scene (Game1.cs:98-138): 4 instances of one transparent sphere."""
    s = SceneData("default_spheres")
    # DiffuseColor (255,0,0,100).ToVector4()
    col = (F(255) / F(255), F(0), F(0), F(100) / F(255))
    s.meshes.append(MeshData(uv_sphere_triangles(radius=2.0, color=col), reflectiveness=0.7, interpolate_normals=True,
                             transparent=True, refraction_index=1.32))
    for x in range(2):
        for y in range(2):
            s.objects.append(ObjectData([0], position=(-7.5 + 5 * x, 2.0, -7.5 + 5 * y)))
    s.lights.append(spot_light((0.0, 5.0, 20.0)))
    s.camera = CameraData((0.0, 16.0, 32.0), (0.0, 0.0, 0.0))
    s.max_reflections = 8
    return s.finalize()


def many_objects_scene(n_objects=40, seed=3):
    """>20 rotated/scaled instances of two shared meshes: exercises the scene-level octree
    (OctreeSpatialManager.cs:101-113, 218-248) and quirks Q6-Q8."""
    rng = np.random.Generator(np.random.PCG64(seed))
    s = SceneData("many_objects")
    s.meshes.append(MeshData(uv_sphere_triangles(radius=1.0, segments=12, rings=8, color=(0.2, 0.7, 1.0, 1.0)),
                             reflectiveness=0.4, interpolate_normals=True))
    s.meshes.append(MeshData(cube_triangles(1.5, color=(1.0, 0.6, 0.2, 0.5)), reflectiveness=0.2,
                             interpolate_normals=False, transparent=True, refraction_index=1.2))
    s.meshes.append(MeshData(ground_triangles(40.0, -6.0, color=(0.9, 0.9, 0.9, 1.0)), reflectiveness=0.3,
                             interpolate_normals=False))
    for i in range(n_objects):
        pos = rng.uniform(-12, 12, size=3)
        pos[1] = rng.uniform(-4, 6)
        rot = rng.uniform(-math.pi, math.pi, size=3)
        sc = rng.uniform(0.6, 1.8, size=3)
        mesh_ids = [i % 2] if i % 5 else [0, 1]
        s.objects.append(ObjectData(mesh_ids, position=tuple(pos), rotation=tuple(rot), scale=tuple(sc)))
    s.objects.append(ObjectData([2]))
    s.lights.append(spot_light((5.0, 30.0, 25.0)))
    s.lights.append(directional_light(xna.normalize(xna.v3(-0.3, 1.0, 0.4)), color=(0.4, 0.4, 0.5)))
    s.camera = CameraData((0.0, 14.0, 34.0), (0.0, 0.0, 0.0))
    s.max_reflections = 4
    return s.finalize()
