"""Host-side scene model and its flattening into the C ABI's plain arrays.

Mirrors, for the hot path only, the objects the reference instantiates from a scene file:
``Scene`` (src/render/scene.cpp:22-100), ``Rectangle`` (src/shapes/rectangle.cpp), ``Cube``
(src/shapes/cube.cpp:109-165), ``Mesh`` (src/render/mesh.cpp), animated ``Instance`` + ``ShapeGroup``
(src/shapes/instance.cpp, src/render/shapegroup.cpp; created by the XML rewrite src/core/xml.cpp:1166-1192),
``SmoothDiffuse`` / ``TwoSidedBRDF``, ``PointLight`` / ``AreaLight``, ``PerspectiveCamera``
(src/sensors/perspective.cpp:172-198), ``HDRFilm`` + reconstruction filters.

``Scene.flatten()`` produces a :class:`FlatScene` holding numpy buffers plus the ctypes
``dtof_scene_desc`` that points into them (see include/dtof.h).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Union

import numpy as np

from . import _abi
from .transform import AnimatedTransform, Transform4, perspective_projection

__all__ = [
    "Bsdf", "Shape", "PointLight", "SpotLight", "DirectionalLight", "ConstantEmitter", "Film", "CorrelatedSampler", "PerspectiveSensor", "Scene", "FlatScene",
    "rectangle", "cube", "mesh",
]

f32 = np.float32


# ------------------------------------------------------------------------------------------------
@dataclass
class Bsdf:
    """``diffuse`` (src/bsdfs/diffuse.cpp) or ``conductor`` (src/bsdfs/conductor.cpp: `reflectance` is its
    specular_reflectance, `eta` / `k` the complex index of refraction, material "none" = (0, 1)), optionally wrapped in
    ``twosided`` (src/bsdfs/twosided.cpp); or ``dielectric`` (src/bsdfs/dielectric.cpp: `eta[0]` = int_ior / ext_ior,
    `reflectance` / `k` = specular_reflectance / specular_transmittance; never two-sided)."""
    reflectance: Sequence[float] = (0.5, 0.5, 0.5)   # SmoothDiffuse default reflectance 0.5
    twosided: bool = False
    kind: int = _abi.BSDF_DIFFUSE
    eta: Sequence[float] = (0.0, 0.0, 0.0)
    k: Sequence[float] = (1.0, 1.0, 1.0)
    alpha: Sequence[float] = (0.0, 0.0)          # roughconductor: alpha_u, alpha_v
    distribution: int = 0                        # roughconductor: 0 beckmann, 1 ggx


@dataclass
class Shape:
    kind: str                                   # 'rectangle' | 'cube' | 'mesh'
    to_world: Union[Transform4, AnimatedTransform, None] = None
    bsdf: Optional[Bsdf] = None
    radiance: Optional[Sequence[float]] = None  # attached area emitter (src/emitters/area.cpp)
    flip_normals: bool = False
    # 'mesh' payload (object space)
    positions: Optional[np.ndarray] = None
    normals: Optional[np.ndarray] = None
    texcoords: Optional[np.ndarray] = None
    faces: Optional[np.ndarray] = None
    id: str = ""
    # the mesh file had no normals and face_normals is false: smooth vertex normals are computed when the shape is
    # flattened, on the positions in the space the reference computes them in (world space for a static shape)
    smooth_normals: bool = False

    @property
    def animated(self) -> bool:
        return isinstance(self.to_world, AnimatedTransform) and self.to_world.size() > 1


def rectangle(**kw) -> Shape:
    return Shape("rectangle", **kw)


def cube(**kw) -> Shape:
    return Shape("cube", **kw)


def mesh(positions, faces, normals=None, texcoords=None, **kw) -> Shape:
    return Shape("mesh", positions=np.asarray(positions, f32).reshape(-1, 3),
                 faces=np.asarray(faces, np.uint32).reshape(-1, 3),
                 normals=None if normals is None else np.asarray(normals, f32).reshape(-1, 3),
                 texcoords=None if texcoords is None else np.asarray(texcoords, f32).reshape(-1, 2), **kw)


@dataclass
class PointLight:
    """src/emitters/point.cpp:65-85: position = ``position`` or translation of ``to_world``."""
    position: Sequence[float] = (0.0, 0.0, 0.0)
    intensity: Sequence[float] = (1.0, 1.0, 1.0)


@dataclass
class SpotLight:
    """src/emitters/spot.cpp (without a projection texture): `to_world` places the light (it looks along +z);
    `cutoff_angle` / `beam_width` in degrees (defaults 20 and 3/4 of the cutoff angle, spot.cpp:102-106)."""
    to_world: Transform4 = field(default_factory=Transform4.identity)
    intensity: Sequence[float] = (1.0, 1.0, 1.0)
    cutoff_angle: float = 20.0
    beam_width: Optional[float] = None


@dataclass
class DirectionalLight:
    """src/emitters/directional.cpp: distant light; `direction` is the direction the light travels in
    (to_world * (0, 0, 1), or the normalised `direction` property)."""
    direction: Sequence[float] = (0.0, 0.0, 1.0)
    irradiance: Sequence[float] = (1.0, 1.0, 1.0)


@dataclass
class ConstantEmitter:
    """src/emitters/constant.cpp: constant environment emitter (`radiance`, RGB); at most one per scene. Its bounding
    sphere comes from the scene's geometry (set_scene, :73-82) and is derived by the library at upload."""
    radiance: Sequence[float] = (1.0, 1.0, 1.0)


@dataclass
class Film:
    """``hdrfilm`` geometry (src/films/hdrfilm.cpp, src/render/film.cpp:17-80)."""
    width: int = 768
    height: int = 576
    crop_offset: Sequence[int] = (0, 0)
    crop_size: Optional[Sequence[int]] = None
    rfilter: str = "gaussian"          # film.cpp:49-54: default reconstruction filter is 'gaussian'
    rfilter_radius: Optional[float] = None   # tent 'radius' (default 1)
    gaussian_stddev: float = 0.5
    mitchell_b: float = 1.0 / 3.0      # mitchell 'B', 'C' (src/rfilters/mitchell.cpp:52-60)
    mitchell_c: float = 1.0 / 3.0
    lanczos_lobes: int = 3             # lanczos 'lobes' = its radius (src/rfilters/lanczos.cpp:38-42)

    def abi(self) -> _abi.Film:
        cw, ch = self.crop_size if self.crop_size is not None else (self.width, self.height)
        kinds = {"box": _abi.RFILTER_BOX, "tent": _abi.RFILTER_TENT, "gaussian": _abi.RFILTER_GAUSSIAN,
                 "mitchell": _abi.RFILTER_MITCHELL, "catmullrom": _abi.RFILTER_CATMULLROM, "lanczos": _abi.RFILTER_LANCZOS}
        if self.rfilter not in kinds:
            raise ValueError(f"rfilter '{self.rfilter}' is not a reconstruction filter (box|tent|gaussian|mitchell|catmullrom|lanczos)")
        if self.rfilter == "box":
            radius = 0.5
        elif self.rfilter == "tent":
            radius = 1.0 if self.rfilter_radius is None else float(self.rfilter_radius)
        elif self.rfilter in ("mitchell", "catmullrom"):
            radius = 2.0
        elif self.rfilter == "lanczos":
            radius = float(int(self.lanczos_lobes))
        else:
            radius = 4.0 * float(self.gaussian_stddev)     # src/rfilters/gaussian.cpp:50-53
        return _abi.Film(int(cw), int(ch), int(self.crop_offset[0]), int(self.crop_offset[1]), kinds[self.rfilter],
                         radius, float(self.gaussian_stddev), float(self.mitchell_b), float(self.mitchell_c))


@dataclass
class CorrelatedSampler:
    """``correlated`` sampler properties (src/render/sampler.cpp:13-14, src/samplers/correlated.cpp:17-23)."""
    sample_count: int = 4
    seed: int = 0
    time_correlate_number: int = 2
    path_correlate_number: Optional[int] = None
    kind: str = "correlated"   # "independent": PCG32Sampler only, usable by the non-Doppler integrators (path, velocity)

    def __post_init__(self):
        if self.path_correlate_number is None:
            self.path_correlate_number = self.time_correlate_number


def parse_fov(fov: float, fov_axis: str, aspect: float) -> float:
    """src/render/sensor.cpp:149-203 (the 'fov' branch; double precision)."""
    fov_axis = fov_axis.lower()
    if fov_axis == "smaller":
        fov_axis = "y" if aspect > 1 else "x"
    elif fov_axis == "larger":
        fov_axis = "x" if aspect > 1 else "y"
    if fov_axis == "x":
        result = fov
    elif fov_axis == "y":
        result = math.degrees(2.0 * math.atan(math.tan(0.5 * math.radians(fov)) * aspect))
    elif fov_axis == "diagonal":
        diagonal = 2.0 * math.tan(0.5 * math.radians(fov))
        width = diagonal / math.sqrt(1.0 + 1.0 / (aspect * aspect))
        result = math.degrees(2.0 * math.atan(width * 0.5))
    else:
        raise ValueError("The 'fov_axis' parameter must be set to one of 'smaller', 'larger', 'diagonal', 'x', or 'y'!")
    if result <= 0.0 or result >= 180.0:
        raise ValueError("The horizontal field of view must be in the range [0, 180]!")
    return result


@dataclass
class PerspectiveSensor:
    to_world: Transform4 = field(default_factory=Transform4.identity)
    fov: float = 45.0
    fov_axis: str = "x"
    near_clip: float = 1e-2
    far_clip: float = 1e4
    shutter_open: float = 0.0
    shutter_close: float = 0.0
    film: Film = field(default_factory=Film)
    sampler: CorrelatedSampler = field(default_factory=CorrelatedSampler)

    def abi(self) -> _abi.Camera:
        if self.shutter_close < self.shutter_open:  # src/render/sensor.cpp:18-20
            raise ValueError("Shutter opening time must be less than or equal to the shutter closing time!")
        if self.near_clip <= 0 or self.near_clip >= self.far_clip:
            raise ValueError("The 'near_clip' parameter must be greater than zero and smaller than 'far_clip'.")
        tw = self.to_world.astype(f32)
        if tw.has_scale():
            raise ValueError("Scale factors in the camera-to-world transformation are not allowed!")
        fa = self.film
        size = (fa.width, fa.height)
        crop = fa.crop_size if fa.crop_size is not None else size
        x_fov = f32(parse_fov(float(self.fov), self.fov_axis, size[0] / float(size[1])))
        c2s = perspective_projection(size, crop, fa.crop_offset, x_fov, f32(self.near_clip), f32(self.far_clip))
        s2c = c2s.inverse().matrix.astype(f32)
        cam = _abi.Camera()
        cam.to_world[:] = tw.m34().tolist()
        cam.sample_to_camera[:] = s2c.reshape(16).tolist()
        cam.near_clip = float(f32(self.near_clip))
        cam.far_clip = float(f32(self.far_clip))
        so, sc = f32(self.shutter_open), f32(self.shutter_close)
        cam.shutter_open = float(so)
        cam.shutter_open_time = float(f32(sc - so))
        return cam


# ------------------------------------------------------------------------------------------------
_CUBE_V = np.array([
    [1, -1, -1], [1, -1, 1], [-1, -1, 1], [-1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, 1, 1], [1, 1, 1],
    [1, -1, -1], [1, 1, -1], [1, 1, 1], [1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1], [-1, -1, 1],
    [-1, -1, 1], [-1, 1, 1], [-1, 1, -1], [-1, -1, -1], [1, 1, -1], [1, -1, -1], [-1, -1, -1], [-1, 1, -1]], f32)
_CUBE_N = np.repeat(np.array([[0, -1, 0], [0, 1, 0], [1, 0, 0], [0, 0, 1], [-1, 0, 0], [0, 0, -1]], f32), 4, axis=0)
_CUBE_UV = np.tile(np.array([[0, 1], [1, 1], [1, 0], [0, 0]], f32), (6, 1))
_CUBE_F = np.array([[0, 1, 2], [3, 0, 2], [4, 5, 6], [7, 4, 6], [8, 9, 10], [11, 8, 10], [12, 13, 14], [15, 12, 14],
                    [16, 17, 18], [19, 16, 18], [20, 21, 22], [23, 20, 22]], np.uint32)


def _normalize_rows(v: np.ndarray) -> np.ndarray:
    # dr::normalize in float32: v * rsqrt(squared_norm)
    out = np.empty_like(v, dtype=f32)
    for i, r in enumerate(v.astype(f32)):
        sq = f32(r[0] * r[0])
        sq = f32(np.float64(r[1]) * np.float64(r[1]) + np.float64(sq))
        sq = f32(np.float64(r[2]) * np.float64(r[2]) + np.float64(sq))
        out[i] = r * f32(f32(1) / np.sqrt(sq, dtype=f32))
    return out


@dataclass
class _FlatMesh:
    positions: np.ndarray
    faces: np.ndarray
    normals: Optional[np.ndarray]
    texcoords: Optional[np.ndarray]
    bsdf: int
    emitter: int
    flip: int
    kind: int
    rect_to_world: np.ndarray


def _flatten_shape(sh: Shape, trafo: Transform4) -> _FlatMesh:
    """Object -> (static: world, animated: object) space triangle mesh, float32."""
    t32 = trafo.astype(f32)
    rect_m = np.zeros(12, f32)
    flip = int(sh.flip_normals)
    if sh.kind == "rectangle":
        if sh.flip_normals:   # rectangle.cpp:91-94: baked into to_world
            t32 = (trafo @ Transform4.scale((1.0, 1.0, -1.0), dtype=trafo.matrix.dtype)).astype(f32)
            flip = 0
        corners = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], f32)
        pos = np.stack([t32.transform_affine_point(c) for c in corners])
        uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], f32)
        # winding such that normalize(cross(p1-p0, p2-p0)) == normalize(to_world * Normal(0,0,1))
        det = np.linalg.det(t32.matrix[:3, :3].astype(np.float64))
        faces = np.array([[0, 1, 2], [0, 2, 3]] if det > 0 else [[0, 2, 1], [0, 3, 2]], np.uint32)
        return _FlatMesh(pos, faces, None, uv, 0, -1, flip, _abi.SHAPE_RECTANGLE, t32.m34())
    if sh.kind == "cube":
        pos = np.stack([t32.transform_affine_point(v) for v in _CUBE_V])
        nrm = _normalize_rows(np.stack([t32.transform_normal(n) for n in _CUBE_N]))
        return _FlatMesh(pos, _CUBE_F.copy(), nrm, _CUBE_UV.copy(), 0, -1, flip, _abi.SHAPE_MESH, rect_m)
    if sh.kind == "mesh":
        ident = np.array_equal(t32.matrix, np.eye(4, dtype=f32))
        pos = sh.positions.astype(f32) if ident else np.stack([t32.transform_affine_point(v) for v in sh.positions])
        nrm = None
        if sh.smooth_normals and sh.normals is None:
            # Mesh::recompute_vertex_normals (src/render/mesh.cpp:283-345) runs after the loader applied to_world
            from .meshio import vertex_normals
            nrm = vertex_normals(np.ascontiguousarray(pos, f32), np.ascontiguousarray(sh.faces, np.uint32))
        elif sh.normals is not None:
            nrm = sh.normals.astype(f32) if ident else _normalize_rows(
                np.stack([t32.transform_normal(n) for n in sh.normals]))
        uv = None if sh.texcoords is None else sh.texcoords.astype(f32)
        return _FlatMesh(np.ascontiguousarray(pos, f32), np.ascontiguousarray(sh.faces, np.uint32), nrm, uv, 0, -1,
                         flip, _abi.SHAPE_MESH, rect_m)
    raise ValueError(f"shape type '{sh.kind}' is outside the hot-path scope (rectangle|cube|mesh)")


class FlatScene:
    """Owns the numpy buffers and the ctypes ``dtof_scene_desc`` pointing into them."""

    def __init__(self, meshes: List[_FlatMesh], instances: List[_abi.Instance], bsdfs: List[_abi.Bsdf],
                 emitters: List[_abi.Emitter], camera: _abi.Camera, film: _abi.Film):
        self._keep = []
        self.n_triangles = int(sum(m.faces.shape[0] for m in meshes))
        self.meshes = (_abi.Mesh * max(1, len(meshes)))()
        for i, m in enumerate(meshes):
            pos = np.ascontiguousarray(m.positions, f32)
            fac = np.ascontiguousarray(m.faces, np.uint32)
            self._keep += [pos, fac]
            cm = self.meshes[i]
            cm.n_vertices, cm.n_faces = pos.shape[0], fac.shape[0]
            cm.positions, cm.faces = _abi.as_fp(pos), _abi.as_up(fac)
            if m.normals is not None:
                n = np.ascontiguousarray(m.normals, f32)
                self._keep.append(n)
                cm.normals = _abi.as_fp(n)
            if m.texcoords is not None:
                t = np.ascontiguousarray(m.texcoords, f32)
                self._keep.append(t)
                cm.texcoords = _abi.as_fp(t)
            cm.bsdf, cm.emitter, cm.flip_normals, cm.kind = m.bsdf, m.emitter, m.flip, m.kind
            cm.rect_to_world[:] = m.rect_to_world.tolist()
        self.instances = (_abi.Instance * max(1, len(instances)))(*instances)
        self.bsdfs = (_abi.Bsdf * max(1, len(bsdfs)))(*bsdfs)
        self.emitters = (_abi.Emitter * max(1, len(emitters)))(*emitters)
        self.desc = _abi.SceneDesc(len(meshes), self.meshes, len(instances), self.instances, len(bsdfs), self.bsdfs,
                                   len(emitters), self.emitters, camera, film)
        self.width, self.height = film.width, film.height


@dataclass
class Scene:
    shapes: List[Shape] = field(default_factory=list)
    emitters: List[PointLight] = field(default_factory=list)
    sensor: PerspectiveSensor = field(default_factory=PerspectiveSensor)
    integrator: object = None   # DopplerToFPathIntegrator (integrator.py)

    def fingerprint(self) -> bytes:
        """Digest of everything flatten() reads (shapes, transforms, BSDFs, emitters, sensor, film, sampler): the key of the
        upload cache in integrator.render / render_distributed, so that an edited scene (a moved shape, another film size,
        another sensor) is flattened and uploaded again instead of silently rendering the stale copy. Arrays above 1 MiB are
        identified by buffer address, shape and their first / last elements instead of being hashed."""
        import dataclasses
        import hashlib
        h = hashlib.blake2b(digest_size=16)
        seen = set()

        def walk(x):
            if isinstance(x, np.ndarray):
                h.update(repr((x.shape, str(x.dtype))).encode())
                if x.nbytes <= (1 << 20):
                    h.update(np.ascontiguousarray(x).tobytes())
                elif x.size:
                    flat = x.reshape(-1)
                    h.update(repr((x.__array_interface__["data"][0], flat[0].item(), flat[-1].item())).encode())
            elif dataclasses.is_dataclass(x) and not isinstance(x, type):
                if id(x) in seen:
                    return
                seen.add(id(x))
                h.update(type(x).__name__.encode())
                for f in dataclasses.fields(x):
                    h.update(f.name.encode())
                    walk(getattr(x, f.name))
            elif isinstance(x, (list, tuple)):
                h.update(b"[")
                for y in x:
                    walk(y)
                h.update(b"]")
            elif isinstance(x, dict):
                for k in sorted(x):
                    h.update(repr(k).encode())
                    walk(x[k])
            elif hasattr(x, "matrix") and isinstance(getattr(x, "matrix"), np.ndarray):     # Transform4
                walk(x.matrix)
            elif hasattr(x, "times") and hasattr(x, "transforms"):                         # AnimatedTransform
                walk(list(x.times))
                walk(list(x.transforms))
            elif hasattr(x, "__dict__") and not callable(x):
                h.update(type(x).__name__.encode())
                walk({k: v for k, v in vars(x).items() if not k.startswith("_")})
            else:
                h.update(repr(x).encode())

        walk(self.shapes)
        walk(self.emitters)
        walk(self.sensor)
        return h.digest()

    def flatten(self) -> FlatScene:
        bsdfs: List[_abi.Bsdf] = []
        bsdf_index = {}

        def bsdf_id(b: Optional[Bsdf]) -> int:
            b = b if b is not None else Bsdf()   # Shape default BSDF: diffuse (src/render/shape.cpp:60-65)
            def rgb(v):
                v = np.atleast_1d(np.asarray(v, f32))
                return tuple(float(x) for x in ((v.tolist() * 3)[:3] if v.size == 1 else v.tolist()))
            conductor = b.kind != _abi.BSDF_DIFFUSE   # every other kind uses eta / k
            key = (b.kind, bool(b.twosided), tuple(float(f32(x)) for x in b.reflectance),
                   rgb(b.eta) if conductor else (0.0,) * 3, rgb(b.k) if conductor else (0.0,) * 3,
                   (float(f32(b.alpha[0])), float(f32(b.alpha[1]))), int(b.distribution))
            if key not in bsdf_index:
                bsdf_index[key] = len(bsdfs)
                bsdfs.append(_abi.Bsdf(b.kind, int(b.twosided), (C.c_float * 3)(*key[2]), (C.c_float * 3)(*key[3]),
                                       (C.c_float * 3)(*key[4]), (C.c_float * 2)(*key[5]), key[6], 0))
            return bsdf_index[key]

        meshes: List[_FlatMesh] = []
        instances: List[_abi.Instance] = []
        emitters: List[_abi.Emitter] = []
        ident12 = Transform4.identity(f32).m34()

        def rgb3(v):
            v = np.atleast_1d(np.asarray(v, f32))
            return (C.c_float * 3)(*((v.tolist() * 3)[:3] if v.size == 1 else v.tolist()))

        # Emitter order = order of appearance among the scene's children (scene.cpp:40-64); area emitters take
        # the slot of their parent shape. Here: shapes first in list order, then free-standing emitters, unless
        # `scene_order` was recorded by the XML loader.
        order = getattr(self, "scene_order", None) or ([("shape", i) for i in range(len(self.shapes))] +
                                                       [("emitter", i) for i in range(len(self.emitters))])
        static = [s for s in self.shapes if not s.animated]
        moving = [s for s in self.shapes if s.animated]
        mesh_of_shape = {}
        for s in static:
            tw = s.to_world if isinstance(s.to_world, Transform4) else (
                s.to_world.transforms[0] if isinstance(s.to_world, AnimatedTransform) and s.to_world.size() == 1
                else Transform4.identity())
            fm = _flatten_shape(s, tw)
            fm.bsdf = bsdf_id(s.bsdf)
            mesh_of_shape[id(s)] = len(meshes)
            meshes.append(fm)
        if static:
            instances.append(_abi.Instance(0, len(static), 0, 0.0, 0.0, (C.c_float * 12)(*ident12.tolist()),
                                           (C.c_float * 12)(*ident12.tolist())))
        for s in moving:
            if s.radiance is not None:   # shapegroup.cpp:27-30
                raise ValueError("Instancing of emitters is not supported")
            fm = _flatten_shape(s, Transform4.identity())
            fm.bsdf = bsdf_id(s.bsdf)
            at: AnimatedTransform = s.to_world
            # Instance::embree_geometry (instance.cpp:295-310): matrices at get_min_time / get_max_time
            t0, t1 = at.get_min_time(), at.get_max_time()
            m0 = Transform4(at.eval(t0), np.eye(4, dtype=f32)).m34()
            m1 = Transform4(at.eval(t1), np.eye(4, dtype=f32)).m34()
            instances.append(_abi.Instance(len(meshes), 1, 1, float(f32(t0)), float(f32(t1)),
                                           (C.c_float * 12)(*m0.tolist()), (C.c_float * 12)(*m1.tolist())))
            mesh_of_shape[id(s)] = len(meshes)
            meshes.append(fm)
        for kind, i in order:
            if kind == "shape":
                s = self.shapes[i]
                if s.radiance is not None:
                    mi_ = mesh_of_shape[id(s)]
                    meshes[mi_].emitter = len(emitters)
                    emitters.append(_abi.Emitter(_abi.EMITTER_AREA, mi_, (C.c_float * 3)(0, 0, 0), rgb3(s.radiance)))
            else:
                e = self.emitters[i]
                if isinstance(e, DirectionalLight):
                    emitters.append(_abi.Emitter(_abi.EMITTER_DIRECTIONAL, 0, (C.c_float * 3)(*[float(f32(x)) for x in e.direction]),
                                                 rgb3(e.irradiance)))
                elif isinstance(e, SpotLight):
                    m = np.asarray(e.to_world.matrix, np.float64)
                    inv = np.asarray(e.to_world.inverse_transpose, np.float64).T[:3, :3].astype(f32)   # tracked inverse
                    cutoff = f32(e.cutoff_angle)
                    beam = f32(e.beam_width) if e.beam_width is not None else cutoff * f32(3.0) / f32(4.0)
                    deg = f32(np.pi / 180.0)
                    emitters.append(_abi.Emitter(_abi.EMITTER_SPOT, 0, (C.c_float * 3)(*[float(f32(x)) for x in m[:3, 3]]),
                                                 rgb3(e.intensity), (C.c_float * 9)(*inv.reshape(-1).tolist()),
                                                 float(cutoff * deg), float(beam * deg)))
                elif isinstance(e, ConstantEmitter):
                    if any(em.kind == _abi.EMITTER_CONSTANT for em in emitters):
                        raise ValueError("Only one environment emitter can be specified per scene.")   # scene.cpp:53-55
                    emitters.append(_abi.Emitter(_abi.EMITTER_CONSTANT, 0, (C.c_float * 3)(0, 0, 0), rgb3(e.radiance)))
                else:
                    emitters.append(_abi.Emitter(_abi.EMITTER_POINT, 0, (C.c_float * 3)(*[float(f32(x)) for x in e.position]),
                                                 rgb3(e.intensity)))
        return FlatScene(meshes, instances, bsdfs, emitters, self.sensor.abi(), self.sensor.film.abi())
