"""Mesh ingest: FBX 6.x (ASCII 6.1 and binary 6100) -> Triangle records, following the
conventions of the reference's build-time content processor.

The reference bakes its models at BUILD time: XNA's closed-source `FbxImporter` turns an .fbx into a
`NodeContent` tree and `TracerModelProcessor` (RayTracePipeline/TracerModelProcessor.cs) walks it:

  * `base.Process` is the stock `ModelProcessor`: its `Scale` / `RotationX,Y,Z` parameters are baked into the
    scene (`MeshHelper.TransformScene`) before anything else (:107);
  * `FindVertices` (:133-242): per `MeshContent`, positions = `Vector3.Transform(v, AbsoluteTransform)`
    (:178-180), vertex normals = `Normalize(Vector3.Transform(n, Transpose(Invert(AbsoluteTransform))))`
    (:141,191-197), `surfaceNormal = Normalize(Cross(v3 - v1, v2 - v1))` (:199-203), `id` = running index
    (:221), `color` = `DiffuseColor.ToVector4()` (:227) unless vertex colours are used (:224-225);
  * `CreateBoundingBox` (:244-307) grows a box that STARTS at (0,0,0)-(0,0,0), so `Mesh.MeshBoundingBox`
    always contains the origin.

This module restates the second and third bullet exactly (binary32, XNA operator order via xna.py) and
replaces the closed importer by a small reader for the FBX files the reference ships: ASCII FBX 6.1
(`RayTraceProjectContent/*.fbx`: Blender exports) and binary FBX 6100 (`Crate_Fragile.FBX`, the Free_crate asset of
BASELINE.json configs[0..1], and the other 3ds Max exports); both parse into the same node tree.  What the closed importer does to axes, winding and
texture coordinates cannot be checked here — PARITY UNPINNED for those three choices, which are made as
SURVEY.md Appendix A.3/A.4 recommends and kept in one place (`IMPORTER`):

  * node transform = Lcl Scaling * RotX * RotY * RotZ * Lcl Translation (row vectors, degrees, FBX eEulerXYZ);
  * polygons are fan-triangulated (a, b, c, d, ...) -> (a, b, c), (a, c, d), ... and each triangle is
    REVERSED, so that `Cross(edge2, edge1)` points out of a mesh whose FBX polygons are counter-clockwise
    seen from outside (XNA's front faces are clockwise);
  * v -> 1 - v for texture coordinates (FBX has v up, XNA v down).

Scene graph: a mesh node's transform is composed with its parents' (`Connections`: local * parent * ... in row-vector
order).  What this reader does not model — PreRotation / PostRotation with RotationActive, a RotationOrder other than
XYZ, non-zero Geometric* offsets, pivots — raises NotImplementedError instead of silently placing the mesh wrongly.

`scenes.crate_scene` restates the crate's geometry from SURVEY.md A.4 (the asset does not travel to the GPU box);
tests/test_ingest.py checks that restatement against the geometry read from `Crate_Fragile.FBX` itself.
"""
import math
import re
import struct

import numpy as np

from . import xna
from .records import TRIANGLE_DTYPE

F = np.float32

IMPORTER = {"reverse_winding": True, "flip_v": True}


# ---------------------------------------------------------------------------------------------
# ASCII FBX 6.1 reader (just what a mesh needs)
# ---------------------------------------------------------------------------------------------
class FbxNode:
    __slots__ = ("name", "values", "children")

    def __init__(self, name, values):
        self.name, self.values, self.children = name, values, []

    def child(self, name):
        for c in self.children:
            if c.name == name:
                return c
        return None

    def all(self, name):
        return [c for c in self.children if c.name == name]


_TOKEN = re.compile(r'"[^"]*"|[^,\s]+')


def _values(text):
    out = []
    for t in _TOKEN.findall(text):
        if t.startswith('"'):
            out.append(t[1:-1])
        else:
            try:
                out.append(int(t))
            except ValueError:
                try:
                    out.append(float(t))
                except ValueError:
                    out.append(t)
    return out


def parse_fbx_ascii(text):
    """`Name: v, v, ... {` / `}` / continuation lines starting with ','  ->  tree of FbxNode."""
    root = FbxNode("", [])
    stack = [root]
    last = None
    for raw in text.splitlines():
        line = raw.strip()
        if not line or line.startswith(";"):
            continue
        if line == "}":
            stack.pop()
            last = None
            continue
        if line.startswith(","):               # a long array continues
            if last is not None:
                last.values.extend(_values(line))
            continue
        m = re.match(r"^([A-Za-z_][\w]*):\s*(.*)$", line)
        if not m:
            if last is not None:               # continuation without a leading comma
                last.values.extend(_values(line))
            continue
        name, rest = m.group(1), m.group(2)
        opens = rest.endswith("{")
        if opens:
            rest = rest[:-1]
        node = FbxNode(name, _values(rest))
        stack[-1].children.append(node)
        last = node
        if opens:
            stack.append(node)
    return root


# ---------------------------------------------------------------------------------------------
# binary FBX 6100 reader (same FbxNode tree as the ASCII reader)
# ---------------------------------------------------------------------------------------------
_BIN_MAGIC = b"Kaydara FBX Binary  \x00\x1a\x00"


def _bin_name(v):
    """Binary files store object names as `Name\0\1Class`; the ASCII form is `Class::Name`."""
    if isinstance(v, str) and "\x00\x01" in v:
        name, cls = v.split("\x00\x01", 1)
        return cls + "::" + name
    return v


def parse_fbx_binary(data):
    """Node record: u32 endOffset, u32 numProperties, u32 propertyListLen, u8 nameLen, name, properties, nested
    records up to endOffset, the list closed by a 13-byte null record.  Property = type byte + payload:
    Y i16, C u8, I i32, F f32, D f64, L i64, S / R u32 length + bytes, and (7000+) f d l i b arrays."""
    import zlib

    if data[:len(_BIN_MAGIC)] != _BIN_MAGIC:
        raise ValueError("not a binary FBX file")
    version = struct.unpack_from("<I", data, 23)[0]
    if version >= 7500:
        raise NotImplementedError("binary FBX %d uses 64-bit node records" % version)
    scalar = {"Y": "<h", "C": "<B", "I": "<i", "F": "<f", "D": "<d", "L": "<q"}
    array = {"f": "<f4", "d": "<f8", "l": "<i8", "i": "<i4", "b": "u1"}

    def node_at(pos):
        end, nprop, plen, nlen = struct.unpack_from("<IIIB", data, pos)
        if end == 0:
            return None, pos + 13
        name = data[pos + 13:pos + 13 + nlen].decode("ascii", "replace")
        q = pos + 13 + nlen
        vals = []
        for _ in range(nprop):
            t = chr(data[q])
            q += 1
            if t in scalar:
                v = struct.unpack_from(scalar[t], data, q)[0]
                q += struct.calcsize(scalar[t])
                vals.append(float(v) if t in "FD" else int(v))
            elif t in "SR":
                n = struct.unpack_from("<I", data, q)[0]
                raw = data[q + 4:q + 4 + n]
                q += 4 + n
                vals.append(_bin_name(raw.decode("latin-1")) if t == "S" else raw)
            elif t in array:
                cnt, enc, clen = struct.unpack_from("<III", data, q)
                raw = data[q + 12:q + 12 + clen]
                q += 12 + clen
                if enc == 1:
                    raw = zlib.decompress(raw)
                a = np.frombuffer(raw, dtype=array[t], count=cnt)
                vals.extend(float(x) for x in a) if t in "fd" else vals.extend(int(x) for x in a)
            else:
                raise ValueError("unknown FBX property type %r at offset %d" % (t, q - 1))
        node = FbxNode(name, vals)
        while q < end:
            child, q = node_at(q)
            if child is None:
                break
            node.children.append(child)
        return node, end

    root = FbxNode("", [version])
    pos = 27
    while pos + 13 <= len(data):
        n, pos = node_at(pos)
        if n is None:
            break
        root.children.append(n)
    return root


def parse_fbx_file(path):
    """Node tree of an FBX 6.x file, ASCII or binary."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:len(_BIN_MAGIC)] == _BIN_MAGIC:
        return parse_fbx_binary(data)
    return parse_fbx_ascii(data.decode("utf-8", "replace"))


def _lcl(props, key, default):
    for p in props.all("Property") if props is not None else []:
        if p.values and p.values[0] == key:
            return tuple(float(v) for v in p.values[-3:])
    return default


def _layer(model, elem, data, index):
    """(mapping, reference, data array, index array or None) of the first layer element `elem`."""
    le = model.child(elem)
    if le is None:
        return None
    mapping = le.child("MappingInformationType").values[0]
    ref = le.child("ReferenceInformationType").values[0]
    d = np.asarray(le.child(data).values, dtype=np.float64)
    idx = le.child(index)
    return mapping, ref, d, (np.asarray(idx.values, dtype=np.int64) if idx is not None and idx.values else None)


def _prop3(props, key):
    for p in props.all("Property") if props is not None else []:
        if p.values and p.values[0] == key and len(p.values) >= 3:
            try:
                return tuple(float(v) for v in p.values[-3:])
            except (TypeError, ValueError):
                return None
    return None


def _prop1(props, key):
    for p in props.all("Property") if props is not None else []:
        if p.values and p.values[0] == key:
            return p.values[-1]
    return None


def _check_supported(name, props):
    """Transform features this reader does not model must not pass silently (a mesh would land in the wrong pose)."""
    for key in ("GeometricTranslation", "GeometricRotation", "RotationOffset", "RotationPivot", "ScalingOffset", "ScalingPivot"):
        v = _prop3(props, key)
        if v is not None and any(abs(c) > 0.0 for c in v):
            raise NotImplementedError("%s: non-zero %s is not supported" % (name, key))
    g = _prop3(props, "GeometricScaling")
    if g is not None and any(abs(c - 1.0) > 0.0 for c in g):
        raise NotImplementedError("%s: GeometricScaling is not supported" % name)
    if _prop1(props, "RotationActive") in (1, True):
        for key in ("PreRotation", "PostRotation"):
            v = _prop3(props, key)
            if v is not None and any(abs(c) > 0.0 for c in v):
                raise NotImplementedError("%s: %s with RotationActive is not supported" % (name, key))
    ro = _prop1(props, "RotationOrder")
    if ro not in (None, 0, "0"):
        raise NotImplementedError("%s: RotationOrder %r is not supported (only XYZ)" % (name, ro))


def read_fbx_meshes(path):
    """Mesh models of an FBX 6.x file (ASCII or binary): list of dicts with name, lcl (translation, rotation in degrees,
    scaling), parents (the lcl triples of the chain of parent models, nearest first), vertices (N, 3), polygons
    (lists of vertex indices), normals / uvs as (mapping, reference, data, index), axes (UpAxis, UpAxisSign,
    FrontAxis, FrontAxisSign, CoordAxis, CoordAxisSign, UnitScaleFactor from GlobalSettings, when present)."""
    root = parse_fbx_file(path)
    objects = root.child("Objects")
    models = {}
    for m in (objects.all("Model") if objects is not None else []):
        if m.values:
            models[str(m.values[0])] = m
    parent_of = {}
    conns = root.child("Connections")
    for c in (conns.all("Connect") if conns is not None else []):
        if len(c.values) >= 3 and c.values[0] == "OO":
            parent_of.setdefault(str(c.values[1]), str(c.values[2]))
    axes = {}
    gs = objects.child("GlobalSettings") if objects is not None else None
    gp = gs.child("Properties60") if gs is not None else None
    for key in ("UpAxis", "UpAxisSign", "FrontAxis", "FrontAxisSign", "CoordAxis", "CoordAxisSign", "UnitScaleFactor"):
        v = _prop1(gp, key)
        if v is not None:
            axes[key] = v

    def lcl_of(model):
        props = model.child("Properties60")
        return (_lcl(props, "Lcl Translation", (0.0, 0.0, 0.0)), _lcl(props, "Lcl Rotation", (0.0, 0.0, 0.0)),
                _lcl(props, "Lcl Scaling", (1.0, 1.0, 1.0)))

    out = []
    for key, m in models.items():
        if len(m.values) < 2 or m.values[1] != "Mesh" or m.child("Vertices") is None:
            continue
        props = m.child("Properties60")
        _check_supported(key, props)
        parents, p, seen = [], parent_of.get(key), set()
        while p is not None and p in models and p not in seen:      # "Model::Scene" (the root) is not in Objects
            seen.add(p)
            _check_supported(p, models[p].child("Properties60"))
            parents.append(lcl_of(models[p]))
            p = parent_of.get(p)
        verts = np.asarray(m.child("Vertices").values, dtype=np.float64).reshape(-1, 3)
        polys, cur = [], []
        for i in m.child("PolygonVertexIndex").values:
            if i < 0:
                cur.append(-i - 1)             # a negative index closes the polygon: index = -(i) - 1
                polys.append(cur)
                cur = []
            else:
                cur.append(i)
        normals = _layer(m, "LayerElementNormal", "Normals", "NormalsIndex")
        uvs = _layer(m, "LayerElementUV", "UV", "UVIndex")
        if normals is not None:
            normals = (normals[0], normals[1], normals[2].reshape(-1, 3), normals[3])
        if uvs is not None:
            uvs = (uvs[0], uvs[1], uvs[2].reshape(-1, 2), uvs[3])
        tex = None
        for t in (objects.all("Texture") if objects is not None else []):
            fn = t.child("RelativeFilename") or t.child("FileName")
            if fn is not None and fn.values:
                tex = str(fn.values[0])
                break
        out.append({"name": key.replace("Model::", ""), "lcl": lcl_of(m), "parents": parents,
                    "vertices": verts, "polygons": polys, "normals": normals, "uvs": uvs, "axes": axes, "texture": tex})
    return out


# ---------------------------------------------------------------------------------------------
# TracerModelProcessor
# ---------------------------------------------------------------------------------------------
def _node_transform(lcl):
    t, r, s = lcl
    m = xna.multiply(xna.create_scale(s), xna.create_rotation_x(F(math.radians(r[0]))))
    m = xna.multiply(m, xna.create_rotation_y(F(math.radians(r[1]))))
    m = xna.multiply(m, xna.create_rotation_z(F(math.radians(r[2]))))
    return xna.multiply(m, xna.create_translation(t))


def _layer_value(layer, poly_vertex, vertex, polygon):
    mapping, ref, data, index = layer
    if mapping in ("ByVertice", "ByVertex", "ByControlPoint"):
        k = vertex
    elif mapping == "ByPolygonVertex":
        k = poly_vertex
    elif mapping == "ByPolygon":
        k = polygon
    else:                                       # AllSame
        k = 0
    if ref == "IndexToDirect" and index is not None:
        k = int(index[k])
    return data[k]


def process_model(meshes, scale=1.0, rotation_deg=(0.0, 0.0, 0.0), diffuse_color=(0, 0, 0, 0)):
    """TracerModelProcessor.Process for the meshes of one file: list of (triangles, meshBoundingBox).

    scale / rotation_deg are the stock ModelProcessor parameters (`ProcessorParameters_Scale` etc. in
    RayTraceProjectContent.contentproj), diffuse_color the processor's DiffuseColor as bytes (R, G, B, A);
    unspecified parameters take CLR defaults — DiffuseColor (0, 0, 0, 0)."""
    # ModelProcessor: rotation (Z, X, Y) then scale, baked into the scene before FindVertices
    rz = xna.create_rotation_z(F(math.radians(rotation_deg[2])))
    rx = xna.create_rotation_x(F(math.radians(rotation_deg[0])))
    ry = xna.create_rotation_y(F(math.radians(rotation_deg[1])))
    scene = xna.multiply(xna.multiply(xna.multiply(rz, rx), ry), xna.create_scale((scale, scale, scale)))
    color = np.asarray(diffuse_color, dtype=F) / F(255.0)          # Color.ToVector4()
    out = []
    for m in meshes:
        local = _node_transform(m["lcl"])
        for parent in m.get("parents", ()):                        # row vectors: local * parent * grandparent * ...
            local = xna.multiply(local, _node_transform(parent))
        absolute = xna.multiply(local, scene)                      # MeshContent.AbsoluteTransform
        inv_t = xna.invert(absolute).T.copy()                      # Transpose(Invert(absoluteTransform)) (:141)
        corners = []
        pv = 0
        for pi, poly in enumerate(m["polygons"]):
            idx = list(range(pv, pv + len(poly)))
            pv += len(poly)
            for k in range(1, len(poly) - 1):                      # fan triangulation
                tri = [(poly[0], idx[0]), (poly[k], idx[k]), (poly[k + 1], idx[k + 1])]
                if IMPORTER["reverse_winding"]:
                    tri = [tri[0], tri[2], tri[1]]
                corners.append((pi, tri))
        t = np.zeros(len(corners), dtype=TRIANGLE_DTYPE)
        for ti, (pi, tri) in enumerate(corners):
            vs, ns, uv = [], [], []
            for vertex, poly_vertex in tri:
                p = m["vertices"][vertex].astype(F)
                vs.append(xna.transform(p, absolute))              # :178-180
                if m["normals"] is not None:
                    n = np.asarray(_layer_value(m["normals"], poly_vertex, vertex, pi), dtype=F)
                    ns.append(xna.normalize(xna.transform(n, inv_t)))   # :191-197
                if m["uvs"] is not None:
                    c = np.asarray(_layer_value(m["uvs"], poly_vertex, vertex, pi), dtype=F)
                    uv.append(np.array([c[0], F(1.0) - c[1]] if IMPORTER["flip_v"] else [c[0], c[1]], dtype=F))
            e1 = (vs[1] - vs[0]).astype(F)
            e2 = (vs[2] - vs[0]).astype(F)
            sn = xna.normalize(xna.cross(e2, e1))                  # :199-203
            rec = t[ti]
            rec["v1"], rec["v2"], rec["v3"] = vs
            if ns:
                rec["n1"], rec["n2"], rec["n3"] = ns
            else:
                rec["n1"] = rec["n2"] = rec["n3"] = sn
            if uv:
                rec["uv1"], rec["uv2"], rec["uv3"] = uv
            rec["surfaceNormal"] = sn
            rec["color"] = color                                   # :227
            rec["id"] = ti                                         # :221
        out.append((t, create_bounding_box(t)))
    return out


def create_bounding_box(tris):
    """TracerModelProcessor.CreateBoundingBox (:244-307): min/max over all vertices, seeded with (0,0,0)."""
    mn, mx = np.zeros(3, dtype=F), np.zeros(3, dtype=F)
    for key in ("v1", "v2", "v3"):
        if len(tris):
            mn = np.minimum(mn, tris[key].min(axis=0))
            mx = np.maximum(mx, tris[key].max(axis=0))
    return np.concatenate([mn, mx]).astype(F)


def load_model(path, **processor_parameters):
    """Content.Load<Model>(name).Tag["Meshes"] for an FBX 6.x file (ASCII or binary): [(triangles, meshBoundingBox), ...]."""
    return process_model(read_fbx_meshes(path), **processor_parameters)
