"""Host-side mirror of the reference's types on the trace path, over the libxnart C ABI.

The reference's host code is C# (no C# toolchain exists in this image — SURVEY.md F8), so this is the
Python stand-in for the shim a maintainer would write (INTEGRATION.md shows the C# P/Invoke form).
Names, argument meaning and error behaviour follow the reference:

    Material / RayTracerTexture     RayTracerTypeLibrary/Material.cs:26-69, RayTracerTexture.cs:10-33
    Mesh (+ Init)                   RayTracerTypeLibrary/Mesh.cs:9-39
    SceneObject                     RayTraceProject/RayTraceProject/SceneObject.cs:117-199
    OctreeSpatialManager            RayTraceProject/RayTraceProject/Spatial/OctreeSpatialManager.cs:35-99
    Camera                          RayTraceProject/RayTraceProject/Camera.cs:20-170
    SpotLight / DirectionalLight    RayTraceProject/RayTraceProject/SpotLight.cs, DirectionalLight.cs
    RayTracer                       RayTraceProject/RayTraceProject/RayTracer.cs:13-126

`RayTracer.RenderAsync()` keeps the reference's contract (dispatcher thread, IsBusy, Progress,
RenderCompleted) and replaces the scanline-worker fan-out of RenderInternal (RayTracer.cs:105-120)
with ONE call of xnart_render_frame.  Nothing here computes pixels on the CPU.
"""
import math
import threading

import numpy as np

from . import capi, xna
from .bmp import read_bmp_argb
from .records import (ADDRESS_CLAMP, ADDRESS_MIRROR, ADDRESS_WRAP, FILTER_BILINEAR, FILTER_POINT, LIGHT_DIRECTIONAL,
                      LIGHT_SPOT, TRIANGLE_DTYPE, make_frame)


class InvalidOperationException(RuntimeError):
    """System.InvalidOperationException (RayTracer.cs:26-27, 62-63)."""


class ArgumentException(ValueError):
    """System.ArgumentException (Material.cs:85,97; SceneObject.cs:123-124)."""


class TextureFiltering:  # Material.cs:12-16
    Point, Bilinear = FILTER_POINT, FILTER_BILINEAR


class UVAddressMode:  # Material.cs:18-23
    Clamp, Wrap, Mirror = ADDRESS_CLAMP, ADDRESS_WRAP, ADDRESS_MIRROR


class RayTracerTexture:
    """RayTracerTexture.cs: ColorData = 32-bit 0xAARRGGBB texels, top-down rows."""

    def __init__(self, path=None, color_data=None):
        self.ColorData = read_bmp_argb(path) if path is not None else color_data


class Material:
    def __init__(self, reflectiveness=0.0, useTexture=False, transparent=False, refractionIndex=0.0, textureFilePath=None):
        self.Reflectiveness = reflectiveness
        self.UseTexture = useTexture
        self.Transparent = transparent
        self.RefractionIndex = refractionIndex
        self.TextureFilePath = textureFilePath
        self.InterpolateNormals = False
        self.Texture = None

    def Init(self):  # Material.cs:59-69
        if self.UseTexture and self.Texture is None:
            self.Texture = RayTracerTexture(self.TextureFilePath)


class Mesh:
    def __init__(self, triangles, material, boundingBox):
        self.Triangles = np.ascontiguousarray(triangles, dtype=TRIANGLE_DTYPE)
        self.MeshMaterial = material
        self.MeshBoundingBox = np.asarray(boundingBox, dtype=np.float32).reshape(6)
        self.itemTreshold = 50  # MeshOctree.cs:42 (sic)
        self._ids = {}  # backend -> mesh id (a Mesh is shared by all instances of a model, SceneObject.cs:126-127)

    def Init(self, backend=None):
        """Mesh.Init (Mesh.cs:27-32): builds the MeshOctree — natively, inside xnart_add_mesh."""
        if backend is None or backend in self._ids:
            return
        m = self.MeshMaterial
        m.Init()
        argb = m.Texture.ColorData if (m.UseTexture and m.Texture is not None) else None
        self._ids[backend] = backend.add_mesh(
            self.Triangles, reflectiveness=m.Reflectiveness, use_texture=m.UseTexture,
            interpolate_normals=m.InterpolateNormals, transparent=m.Transparent, refraction_index=m.RefractionIndex,
            argb=argb, box=self.MeshBoundingBox, item_threshold=self.itemTreshold)


def LoadModel(path, Scale=1.0, RotationX=0.0, RotationY=0.0, RotationZ=0.0, DiffuseColor=(0, 0, 0, 0), Reflectiveness=0.0,
              UseTexture=False, Transparent=False, RefractionIndex=0.0, TextureFilePath=None, InterpolateNormals=True):
    """`Content.Load<Model>(name).Tag["Meshes"]` for an ASCII FBX 6.1 asset: what FbxImporter + TracerModelProcessor
    (RayTracePipeline/TracerModelProcessor.cs:105-242) bake at build time, with the processor parameters of
    RayTraceProjectContent.contentproj as keyword arguments (`ingest.py`; importer conventions unpinned)."""
    from . import ingest

    mat = Material(Reflectiveness, UseTexture, Transparent, RefractionIndex, TextureFilePath)  # CreateMaterial (:121-131)
    mat.InterpolateNormals = InterpolateNormals
    models = ingest.load_model(path, scale=Scale, rotation_deg=(RotationX, RotationY, RotationZ), diffuse_color=DiffuseColor)
    return [Mesh(tris, mat, box) for tris, box in models]


class SceneObject:
    """An instance of a model's meshes (SceneObject.cs).  `meshes` plays the role of Model.Tag["Meshes"]."""

    def __init__(self, meshes, pos=(0, 0, 0), rot=(0, 0, 0)):
        if not meshes or not all(isinstance(m, Mesh) for m in meshes):
            raise ArgumentException("Model must be built using the TracerModelProcessor")  # SceneObject.cs:123-124
        self.Meshes = list(meshes)
        self.Position, self.Rotation, self.Scale = tuple(pos), tuple(rot), (1.0, 1.0, 1.0)
        self.Name = ""
        mn, mx = np.zeros(3, np.float32), np.zeros(3, np.float32)  # default(BoundingBox) merged with the mesh boxes (:131)
        for m in self.Meshes:
            mn, mx = np.minimum(mn, m.MeshBoundingBox[:3]), np.maximum(mx, m.MeshBoundingBox[3:])
        self.BoundingBox = np.concatenate([mn, mx]).astype(np.float32)

    def _build_world(self):  # SceneObject.cs:183-199
        s = xna.create_scale(self.Scale)
        r = xna.multiply(xna.multiply(xna.create_rotation_x(self.Rotation[0]), xna.create_rotation_y(self.Rotation[1])),
                         xna.create_rotation_z(self.Rotation[2]))
        self._world = xna.multiply(xna.multiply(s, r), xna.create_translation(self.Position))
        self._inverse = xna.invert(self._world)

    @property
    def World(self):
        self._build_world()
        return self._world

    @property
    def InverseWorld(self):
        self._build_world()
        return self._inverse


class OctreeSpatialManager:
    """ISpatialManager (ISpatialManager.cs:10-16): Bodies + Build(); ray queries run on the device."""

    def __init__(self, backend=None):
        self.Bodies = []
        self.itemTreshold = 20  # OctreeSpatialManager.cs:50
        self.backend = backend if backend is not None else capi.Xnart()
        self.built = False

    def Build(self):  # OctreeSpatialManager.cs:64-99
        """Rebuilds from `Bodies`, as the reference does on every call: the backend's object list is cleared first, so a
        second Build() (after appending a body, say) does not duplicate the scene."""
        self.backend.clear_objects()
        self._poses = []
        for obj in self.Bodies:
            for m in obj.Meshes:
                m.Init(self.backend)
            self.backend.add_object([m._ids[self.backend] for m in obj.Meshes], obj.World, obj.InverseWorld, obj.BoundingBox)
            self._poses.append((tuple(obj.Position), tuple(obj.Rotation), tuple(obj.Scale)))
        self._built_bodies = list(self.Bodies)
        self.backend.build_scene(self.itemTreshold)
        self.built = True

    def _sync_transforms(self):
        """The reference reads SceneObject.World / InverseWorld per ray (OctreeSpatialManager.cs:358-364): a body moved
        after Build() renders at its new pose, with the octree it was built with.  Poses that changed since the last
        upload are sent to the backend (xnart_set_object_transform); bodies added or removed need a new Build()."""
        if not self.built:
            return
        for i, obj in enumerate(self._built_bodies):
            pose = (tuple(obj.Position), tuple(obj.Rotation), tuple(obj.Scale))
            if pose != self._poses[i]:
                self.backend.set_object_transform(i, obj.World, obj.InverseWorld)
                self._poses[i] = pose

    def GetRayIntersection(self, ray_position, ray_direction, ignoreTriangle=None):
        """ISpatialManager.GetRayIntersection for one ray; ignoreTriangle = (Mesh, triangle index) or None."""
        self._sync_transforms()
        ign = None
        if ignoreTriangle is not None:
            ign = [[ignoreTriangle[0]._ids[self.backend], ignoreTriangle[1]]]
        r = self.backend.trace_rays(np.array([list(ray_position) + list(ray_direction)], dtype=np.float32), ign)
        if r["tri"][0] < 0:
            return None
        return dict(mesh=int(r["mesh"][0]), triangle=int(r["tri"][0]), u=float(r["u"][0]), v=float(r["v"][0]),
                    d=float(r["d"][0]), worldPosition=r["pos"][0])


class Camera:
    def __init__(self, position, target, up, fieldOfView, aspectRatio, nearClippingPlane, farClippingPlane):
        self.Position, self.Target, self.Up = position, target, up
        self.FieldOfView, self.AspectRatio = fieldOfView, aspectRatio
        self.NearClippingPlane, self.FarClippingPlane = nearClippingPlane, farClippingPlane

    @property
    def View(self):  # rebuilt on every read, Camera.cs:153-160
        return xna.create_look_at(self.Position, self.Target, self.Up)

    @property
    def Projection(self):  # Camera.cs:163-170
        return xna.create_perspective_fov(self.FieldOfView, self.AspectRatio, self.NearClippingPlane, self.FarClippingPlane)


class SpotLight:
    IsPositionable = True

    def __init__(self):  # SpotLight.cs:31-35
        self.Position = (0.0, 0.0, 0.0)
        self.Direction = (0.0, -1.0, 0.0)
        self.Color = (1.0, 1.0, 1.0)
        self.DecayExponent = 1.3
        self.Intensity = 1.0
        self.SpotAngle = math.pi / 2

    def _record(self):
        return dict(kind=LIGHT_SPOT, position=self.Position, direction=self.Direction, color=self.Color,
                    intensity=self.Intensity, decayExponent=self.DecayExponent, spotAngle=np.float32(self.SpotAngle))


class DirectionalLight:
    IsPositionable = False

    def __init__(self):  # DirectionalLight.cs:18-21
        self.Direction = (0.0, -1.0, 0.0)
        self.Color = (1.0, 1.0, 1.0)
        self.Intensity = 1.0

    def _record(self):
        return dict(kind=LIGHT_DIRECTIONAL, position=(0, 0, 0), direction=self.Direction, color=self.Color,
                    intensity=self.Intensity, decayExponent=0.0, spotAngle=0.0)


class RayTracer:
    """RayTracer.cs:13-126 with the worker fan-out replaced by one xnart_render_frame call."""

    def __init__(self):
        self.CurrentScene = None
        self.CurrentCamera = None
        self.CurrentTarget = None      # (H, W) uint32 array standing in for RenderTarget2D.SetData<Color>
        self.Viewport = None           # (width, height) — GraphicsDevice.Viewport
        self.MaxReflections = 0        # no initialiser in the reference (RayTracer.cs:33); Game1.cs:126 sets 8
        self.TextureFiltering = TextureFiltering.Point
        self.AddressMode = UVAddressMode.Wrap
        self.Lights = []
        self.UseMultisampling = False
        self.MultisampleQuality = 1
        self.RenderCompleted = []      # callables(sender, error) — the RenderCompleted event
        self._busy = False
        self._lock = threading.Lock()
        self._thread = None

    @property
    def IsBusy(self):
        return self._busy

    @property
    def Progress(self):  # RayTracer.cs:43-46
        return self.CurrentScene.backend.progress() if self.CurrentScene is not None else 0.0

    def RenderAsync(self):
        with self._lock:
            if self._busy:
                raise InvalidOperationException("Cannot render while the ray tracer is busy.")  # RayTracer.cs:62-63
            self._busy = True
        self._thread = threading.Thread(target=self._render_internal, name="RenderDispatcherThread", daemon=True)
        self._thread.start()
        return self._thread

    # ---- adaptive multisampling (RayTracer.cs:128-311) -------------------------------------------
    TRESHOLD = np.float32(0.5)  # RayTracer.cs:340
    LR_DESTINATION = 1          # RayTracer.cs:305 stores the refined lower-right colour into urColor (index 1), not lrColor (3)

    def _render_with_multisampling(self, be, frame, w, h):
        """RenderInternalWithMultisampling / RenderFirstPass (RayTracer.cs:128-213): every pixel (x, y) gets
        GetColorForQuadrant(x, y, 1.0f, 0).  The recursion is run level by level: all quadrants of one
        iteration depth are cast as ONE batch through `xnart_cast_rays` (4 rays each), the refinement test
        (:279-306) picks the quadrants of the next depth, and the colours are combined on the way back."""
        ys, xs = np.mgrid[0:h, 0:w]
        cx = xs.astype(np.float32).ravel()
        cy = ys.astype(np.float32).ravel()
        size = np.full(cx.shape, 1.0, dtype=np.float32)
        return self._color_for_quadrants(be, frame, cx, cy, size, 0).reshape(h, w)

    def _color_for_quadrants(self, be, frame, cx, cy, size, iteration):
        """GetColorForQuadrant (RayTracer.cs:215-311) for arrays of quadrants of one iteration depth."""
        from . import xna

        n = len(cx)
        if n == 0:
            return np.zeros(0, dtype=np.uint32)
        q = size * np.float32(0.25)                                        # quarterSize (:218)
        sx = np.stack([cx - q, cx + q, cx - q, cx + q], axis=1)             # ul, ur, ll, lr (:221,237,251,265)
        sy = np.stack([cy - q, cy - q, cy + q, cy + q], axis=1)
        col = be.cast_rays(frame, np.stack([sx, sy], axis=2).reshape(-1, 2)).reshape(n, 4)
        if iteration < self.MultisampleQuality:                             # :279
            v = xna.color_to_vector3(col)                                   # (n, 4, 3)
            average = (((v[:, 0] + v[:, 1]) + v[:, 2]) + v[:, 3]) * np.float32(0.25)   # Vector3 / 4.0f multiplies by 1f/4f
            avg_len = xna.length3(average)
            refine = np.abs(avg_len[:, None] - xna.length3(v)) > self.TRESHOLD         # (n, 4)
            half = size / np.float32(2.0)
            # the four recursive calls (:289-306); the result of the lower-right one is stored into urColor (:305)
            for k, dst in ((0, 0), (1, 1), (2, 2), (3, self.LR_DESTINATION)):
                m = refine[:, k]
                if m.any():
                    col[m, dst] = self._color_for_quadrants(be, frame, sx[m, k], sy[m, k], half[m], iteration + 1)
        v = xna.color_to_vector3(col)
        total = ((v[:, 0] + v[:, 1]) + v[:, 2]) + v[:, 3]
        return xna.color_from_vector3(total * np.float32(0.25))             # :309

    def Render(self):
        """Synchronous convenience: RenderAsync + wait."""
        self.RenderAsync().join()
        return self.CurrentTarget

    def _render_internal(self):  # RayTracer.cs:103-126
        err = None
        try:
            be = self.CurrentScene.backend
            self.CurrentScene._sync_transforms()
            w, h = self.Viewport
            be.set_lights([L._record() for L in self.Lights])
            frame = make_frame(self.CurrentCamera.View, self.CurrentCamera.Projection, w, h,
                               max_reflections=self.MaxReflections, filtering=self.TextureFiltering,
                               address_mode=self.AddressMode)
            if self.CurrentTarget is None or self.CurrentTarget.shape != (h, w):
                self.CurrentTarget = np.zeros((h, w), dtype=np.uint32)
            try:
                if self.UseMultisampling:
                    self.CurrentTarget[...] = self._render_with_multisampling(be, frame, w, h)
                else:
                    be.render_frame(frame, out=self.CurrentTarget)
            except capi.XnartError as e:
                if e.code == 1:
                    raise ArgumentException(str(e))
                raise
        except Exception as e:  # surfaced through the completion event, like AsyncCompletedEventArgs.Error
            err = e
        finally:
            self._busy = False  # OnRenderCompleted (RayTracer.cs:431-437)
            for cb in list(self.RenderCompleted):
                cb(self, err)
        if err is not None and not self.RenderCompleted:
            raise err
