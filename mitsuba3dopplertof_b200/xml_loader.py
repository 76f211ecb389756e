"""Minimal scene-XML reader for the subset the hot path needs, so a reference `scene.xml`
(e.g. configs_example/scene.xml) loads unchanged.

Follows src/core/xml.cpp: `<default>` + `$name` substitution (:300-330, command-line `-Dk=v` overrides),
`<transform>` ops composed by LEFT-multiplication in double precision (:820-1007), `<animation>` with
`<transform time=..>` keyframes (:520-524,882-900,996-1007), `<ref id>` resolution, and the rewrite of a
shape with an animated `to_world` into shapegroup + instance (:1166-1192; done in scene.Scene.flatten).
Anything outside the hot-path scope raises.
"""
from __future__ import annotations

import os
import re
import xml.etree.ElementTree as ET
from typing import Dict, Optional

import numpy as np

from .integrator import DopplerToFPathIntegrator, PathIntegrator, VelocityIntegrator
from .scene import (Bsdf, ConstantEmitter, CorrelatedSampler, DirectionalLight, Film, PerspectiveSensor, PointLight, Scene, Shape,
                    SpotLight)
from .transform import AnimatedTransform, Transform4

__all__ = ["load_file", "load_string"]


def _floats(s: str):
    return [float(t) for t in re.split(r"[\s,]+", s.strip()) if t]


# child tags that are objects or structure, handled by the code that walks the parent (never property values)
_STRUCTURAL = frozenset(("transform", "animation", "bsdf", "emitter", "ref", "sampler", "film", "rfilter", "shape", "sensor",
                         "integrator", "default", "include", "alias"))


class _Loader:
    def __init__(self, params: Optional[Dict[str, str]], base_dir: str):
        self.defaults: Dict[str, str] = dict(params or {})
        self.cli = set(self.defaults)
        self.base_dir = base_dir
        self.bsdfs: Dict[str, Bsdf] = {}

    def sub(self, s: str) -> str:
        def rep(m):
            k = m.group(1)
            if k not in self.defaults:
                raise ValueError(f"undefined parameter ${k}")
            return self.defaults[k]
        return re.sub(r"\$(\w+)", rep, s)

    def attr(self, node, name, default=None):
        v = node.get(name)
        return default if v is None else self.sub(v)

    # ---- values -----------------------------------------------------------------------------------
    def props(self, node, skip=()):
        out = {}
        for ch in node:
            if ch.tag in skip:
                continue
            n = ch.get("name")
            if ch.tag == "float":
                out[n] = float(self.attr(ch, "value"))
            elif ch.tag == "integer":
                out[n] = int(self.attr(ch, "value"))
            elif ch.tag == "boolean":
                out[n] = self.attr(ch, "value").strip().lower() == "true"
            elif ch.tag == "string":
                out[n] = self.attr(ch, "value")
            elif ch.tag in ("rgb", "spectrum"):
                if ch.get("value") is None:   # <spectrum filename=...>: a tabulated spectrum, not a constant
                    raise ValueError(f"<{ch.tag}> '{n}' without a constant value is outside the supported subset")
                v = _floats(self.attr(ch, "value"))
                if len(v) not in (1, 3):
                    raise ValueError(f"<{ch.tag}> '{n}': only constant / RGB values are in scope")
                out[n] = tuple(v * 3) if len(v) == 1 else tuple(v)
            elif ch.tag in ("point", "vector"):
                out[n] = self.vec(ch)
            elif ch.tag in _STRUCTURAL and not (ch.tag == "ref" and n is not None):
                continue   # child objects: the caller handles them (or refuses them)
            else:
                # a <texture>, a named <ref> (a texture / spectrum bound to a property) or any tag this subset does not
                # know: the value would silently fall back to the plugin's default (0.5 reflectance, unit radiance ...)
                raise ValueError(f"<{ch.tag}{' name=' + repr(n) if n else ''}> inside <{node.tag}"
                                 f"{' type=' + repr(node.get('type')) if node.get('type') else ''}> is outside the supported "
                                 "subset: only constant float / integer / boolean / string / rgb / spectrum / point / vector "
                                 "properties are read (textures and spatially varying values are not in scope)")
        return out

    def vec(self, node, default=0.0):
        if node.get("value") is not None:
            v = _floats(self.attr(node, "value"))
            return tuple(v * 3) if len(v) == 1 else tuple(v)
        return tuple(float(self.attr(node, k, default)) for k in "xyz")

    def transform(self, node) -> Transform4:
        t = Transform4.identity(np.float64)
        for op in node:
            if op.tag == "matrix":
                v = _floats(self.attr(op, "value"))
                if len(v) == 9:
                    m = np.eye(4)
                    m[:3, :3] = np.array(v).reshape(3, 3)
                    v = m.reshape(16)
                if len(v) != 16:
                    raise ValueError("matrix: expected 16 or 9 values")
                # stof<Float=double> then Transform4f(matrix) in double
                o = Transform4.from_matrix(v)
            elif op.tag == "translate":
                o = Transform4.translate(self.vec(op))
            elif op.tag == "scale":
                o = Transform4.scale(self.vec(op, 1.0))
            elif op.tag == "rotate":
                o = Transform4.rotate(self.vec(op), float(self.attr(op, "angle")))
            elif op.tag in ("lookat", "look_at"):
                origin, target = _floats(self.attr(op, "origin")), _floats(self.attr(op, "target"))
                up = _floats(self.attr(op, "up", "0,0,0"))
                if not any(up):
                    raise ValueError("lookat without 'up' is outside the supported subset")
                o = Transform4.look_at(origin, target, up)
            else:
                raise ValueError(f"unsupported transform op <{op.tag}>")
            t = o @ t
        return t

    def animation(self, node) -> AnimatedTransform:
        at = AnimatedTransform()
        for kf in node:
            if kf.tag != "transform":
                raise ValueError("<animation> may only contain <transform time=..> keyframes")
            at.append(float(self.attr(kf, "time")), self.transform(kf))
        return at

    # ---- objects ----------------------------------------------------------------------------------
    def bsdf(self, node) -> Bsdf:
        typ = self.attr(node, "type")
        if typ == "twosided":
            inner = [c for c in node if c.tag in ("bsdf", "ref")]
            if len(inner) != 1:
                raise ValueError("twosided with two different BRDFs is outside the hot-path scope")
            b = self.bsdf_or_ref(inner[0])
            from . import _abi
            if b.kind in (_abi.BSDF_DIELECTRIC, _abi.BSDF_THINDIELECTRIC, _abi.BSDF_ROUGHDIELECTRIC):   # twosided.cpp:102-103
                raise ValueError("Only materials without a transmission component can be nested!")
            return Bsdf(b.reflectance, True, b.kind, b.eta, b.k, b.alpha, b.distribution)
        if typ == "diffuse":
            p = self.props(node)
            unknown = set(p) - {"reflectance"}
            if unknown:
                raise ValueError(f"diffuse: unreferenced property {sorted(unknown)}")
            return Bsdf(p.get("reflectance", (0.5, 0.5, 0.5)), False)
        if typ == "conductor":   # SmoothConductor ctor, src/bsdfs/conductor.cpp:213-230
            from . import _abi
            p = self.props(node)
            unknown = set(p) - {"specular_reflectance", "material", "eta", "k"}
            if unknown:
                raise ValueError(f"conductor: unreferenced property {sorted(unknown)}")
            material = p.get("material", "none")
            if material != "none":   # conductor.cpp:221-229: a named material replaces eta / k
                if "eta" in p:
                    raise ValueError("Should specify either (eta, k) or material, not both.")
                p["eta"], p["k"] = conductor_ior(material)
            return Bsdf(p.get("specular_reflectance", (1.0, 1.0, 1.0)), False, _abi.BSDF_CONDUCTOR,
                        p.get("eta", (0.0, 0.0, 0.0)), p.get("k", (1.0, 1.0, 1.0)))
        if typ == "roughconductor":   # RoughConductor ctor, src/bsdfs/roughconductor.cpp:160-211
            from . import _abi
            p = self.props(node)
            unknown = set(p) - {"specular_reflectance", "material", "eta", "k", "distribution", "alpha", "alpha_u", "alpha_v",
                                "sample_visible"}
            if unknown:
                raise ValueError(f"roughconductor: unreferenced property {sorted(unknown)}")
            material = p.get("material", "none")
            if material != "none":   # conductor.cpp:221-229: a named material replaces eta / k
                if "eta" in p:
                    raise ValueError("Should specify either (eta, k) or material, not both.")
                p["eta"], p["k"] = conductor_ior(material)
            distr = p.get("distribution", "beckmann")
            if distr not in ("beckmann", "ggx"):
                raise ValueError(f'Specified an invalid distribution "{distr}", must be "beckmann" or "ggx"!')
            if not p.get("sample_visible", True):
                raise ValueError("roughconductor with sample_visible=false is outside the hot-path scope")
            if "alpha_u" in p or "alpha_v" in p:
                if "alpha_u" not in p or "alpha_v" not in p:
                    raise ValueError("Microfacet model: both 'alpha_u' and 'alpha_v' must be specified.")
                if "alpha" in p:
                    raise ValueError("Microfacet model: please specifyeither 'alpha' or 'alpha_u'/'alpha_v'.")
                alpha = (float(p["alpha_u"]), float(p["alpha_v"]))
            else:
                alpha = (float(p.get("alpha", 0.1)),) * 2
            return Bsdf(p.get("specular_reflectance", (1.0, 1.0, 1.0)), False, _abi.BSDF_ROUGHCONDUCTOR,
                        p.get("eta", (0.0, 0.0, 0.0)), p.get("k", (1.0, 1.0, 1.0)), alpha, 1 if distr == "ggx" else 0)
        if typ == "roughdielectric":   # RoughDielectric ctor, src/bsdfs/roughdielectric.cpp:161-213
            from . import _abi
            p = self.props(node)
            unknown = set(p) - {"int_ior", "ext_ior", "specular_reflectance", "specular_transmittance", "distribution", "alpha",
                                "alpha_u", "alpha_v", "sample_visible"}
            if unknown:
                raise ValueError(f"roughdielectric: unreferenced property {sorted(unknown)}")
            int_ior, ext_ior = lookup_ior(p.get("int_ior", "bk7")), lookup_ior(p.get("ext_ior", "air"))
            if int_ior < 0 or ext_ior < 0 or int_ior == ext_ior:
                raise ValueError("The interior and exterior indices of refraction must be positive and differ!")
            distr = str(p.get("distribution", "beckmann")).lower()
            if distr not in ("beckmann", "ggx"):
                raise ValueError(f'Specified an invalid distribution "{distr}", must be "beckmann" or "ggx"!')
            if not p.get("sample_visible", True):
                raise ValueError("roughdielectric with sample_visible=false is outside the hot-path scope")
            if "alpha_u" in p or "alpha_v" in p:
                if "alpha_u" not in p or "alpha_v" not in p:
                    raise ValueError("Microfacet model: both 'alpha_u' and 'alpha_v' must be specified.")
                if "alpha" in p:
                    raise ValueError("Microfacet model: please specifyeither 'alpha' or 'alpha_u'/'alpha_v'.")
                alpha = (float(p["alpha_u"]), float(p["alpha_v"]))
            else:
                alpha = (float(p.get("alpha", 0.1)),) * 2
            eta = float(np.float32(int_ior) / np.float32(ext_ior))
            return Bsdf(p.get("specular_reflectance", (1.0, 1.0, 1.0)), False, _abi.BSDF_ROUGHDIELECTRIC, (eta, 0.0, 0.0),
                        p.get("specular_transmittance", (1.0, 1.0, 1.0)), alpha, 1 if distr == "ggx" else 0)
        if typ == "plastic":   # SmoothPlastic ctor, src/bsdfs/plastic.cpp:157-183
            from . import _abi
            p = self.props(node)
            unknown = set(p) - {"int_ior", "ext_ior", "diffuse_reflectance", "specular_reflectance", "nonlinear"}
            if unknown:
                raise ValueError(f"plastic: unreferenced property {sorted(unknown)}")
            int_ior, ext_ior = lookup_ior(p.get("int_ior", "polypropylene")), lookup_ior(p.get("ext_ior", "air"))
            if int_ior < 0 or ext_ior < 0:
                raise ValueError("The interior and exterior indices of refraction must be positive!")
            eta = float(np.float32(int_ior) / np.float32(ext_ior))
            return Bsdf(p.get("diffuse_reflectance", (0.5, 0.5, 0.5)), False, _abi.BSDF_PLASTIC,
                        (eta, 1.0 if p.get("nonlinear", False) else 0.0, 0.0), p.get("specular_reflectance", (1.0, 1.0, 1.0)))
        if typ in ("dielectric", "thindielectric"):   # dielectric.cpp:199-228, thindielectric.cpp:104-126
            from . import _abi
            p = self.props(node)
            unknown = set(p) - {"int_ior", "ext_ior", "specular_reflectance", "specular_transmittance"}
            if unknown:
                raise ValueError(f"{typ}: unreferenced property {sorted(unknown)}")
            int_ior, ext_ior = lookup_ior(p.get("int_ior", "bk7")), lookup_ior(p.get("ext_ior", "air"))
            if int_ior < 0 or ext_ior < 0:
                raise ValueError("The interior and exterior indices of refraction must be positive!")
            eta = float(np.float32(int_ior) / np.float32(ext_ior))
            return Bsdf(p.get("specular_reflectance", (1.0, 1.0, 1.0)), False,
                        _abi.BSDF_DIELECTRIC if typ == "dielectric" else _abi.BSDF_THINDIELECTRIC, (eta, 0.0, 0.0),
                        p.get("specular_transmittance", (1.0, 1.0, 1.0)))
        raise ValueError(f"bsdf type '{typ}' is outside the hot-path scope (diffuse|conductor|roughconductor|dielectric|thindielectric|roughdielectric|plastic|twosided)")

    def bsdf_or_ref(self, node) -> Bsdf:
        if node.tag == "ref":
            rid = self.attr(node, "id")
            if rid not in self.bsdfs:
                raise ValueError(f"reference to unknown id '{rid}'")
            return self.bsdfs[rid]
        return self.bsdf(node)

    def shape(self, node) -> Shape:
        typ = self.attr(node, "type")
        p = self.props(node)
        sh = Shape({"rectangle": "rectangle", "cube": "cube", "obj": "mesh", "ply": "mesh", "serialized": "mesh"}.get(typ, typ),
                   id=node.get("id", ""))
        sh.flip_normals = bool(p.pop("flip_normals", False))
        for ch in node:
            if ch.tag == "transform" and ch.get("name") == "to_world":
                sh.to_world = self.transform(ch)
            elif ch.tag == "animation" and ch.get("name") == "to_world":
                sh.to_world = self.animation(ch)
            elif ch.tag in ("bsdf", "ref"):
                sh.bsdf = self.bsdf_or_ref(ch)
            elif ch.tag == "emitter":
                if self.attr(ch, "type") != "area":
                    raise ValueError("only 'area' emitters can be attached to shapes")
                sh.radiance = self.props(ch).get("radiance", (1.0, 1.0, 1.0))
        if typ in ("obj", "ply", "serialized"):
            from .meshio import load_mesh, load_serialized
            fn = p.pop("filename")
            face_normals = bool(p.pop("face_normals", False))
            if typ == "serialized":   # the plugin, not the file extension, selects the loader (serialized.cpp:229-248)
                pos, faces, nrm, uv = load_serialized(os.path.join(self.base_dir, fn), int(p.pop("shape_index", 0)))
            else:
                flip_uv = bool(p.pop("flip_tex_coords", True)) if typ == "obj" else True   # obj.cpp:151 (ply.cpp has no such property)
                pos, faces, nrm, uv = load_mesh(os.path.join(self.base_dir, fn), face_normals=face_normals, compute_missing=False,
                                                flip_tex_coords=flip_uv)
            if face_normals:
                nrm = None
            sh.positions, sh.faces, sh.normals, sh.texcoords = pos, faces, nrm, uv
            sh.smooth_normals = nrm is None and not face_normals   # computed at flatten time, after to_world
        elif typ not in ("rectangle", "cube"):
            raise ValueError(f"shape type '{typ}' is outside the hot-path scope (rectangle|cube|obj|ply|serialized)")
        if p:
            raise ValueError(f"shape '{typ}': unreferenced property {sorted(p)}")
        return sh

    def sensor(self, node) -> PerspectiveSensor:
        if self.attr(node, "type") != "perspective":
            raise ValueError("only the 'perspective' sensor is in the hot-path scope")
        p = self.props(node)
        s = PerspectiveSensor()
        # a sensor without a <sampler> gets the reference's default: `independent`, 4 spp (src/render/sensor.cpp:47-48)
        s.sampler = CorrelatedSampler(4, 0, 1, 1, kind="independent")
        for ch in node:
            if ch.tag == "transform" and ch.get("name") == "to_world":
                s.to_world = self.transform(ch)
            elif ch.tag == "sampler":
                kind = self.attr(ch, "type")
                if kind not in ("correlated", "independent"):
                    raise ValueError(f"sampler '{kind}' is outside the hot-path scope (correlated; independent for path/velocity)")
                sp = self.props(ch)
                if kind == "independent":   # PCG32Sampler only: the stream `path` / `velocity` draw from (sampler.cpp:115-134)
                    s.sampler = CorrelatedSampler(int(sp.pop("sample_count", 4)), int(sp.pop("seed", 0)), 1, 1, kind="independent")
                else:
                    s.sampler = CorrelatedSampler(int(sp.pop("sample_count", 4)), int(sp.pop("seed", 0)),
                                                  int(sp.pop("time_correlate_number", 2)), sp.pop("path_correlate_number", None))
                if sp:   # e.g. use_stratified_sampling_for_each_interval is an INTEGRATOR property
                    raise ValueError(f"{kind} sampler: unreferenced property {sorted(sp)}")
            elif ch.tag == "film":
                fp = self.props(ch)
                f = Film(int(fp.pop("width", 768)), int(fp.pop("height", 576)))
                if "crop_width" in fp or "crop_height" in fp:
                    f.crop_size = (int(fp.pop("crop_width", f.width)), int(fp.pop("crop_height", f.height)))
                    f.crop_offset = (int(fp.pop("crop_offset_x", 0)), int(fp.pop("crop_offset_y", 0)))
                # sample_border (film.cpp:35, integrator.cpp:176-178) enlarges the sampled region by the filter's border: not
                # built, and silently ignoring it would change the border pixels -- refuse. The kernels produce RGB (+ weight).
                if fp.pop("sample_border", False):
                    raise ValueError("hdrfilm: sample_border=true is outside the hot-path scope")
                if fp.pop("pixel_format", "rgb") != "rgb":
                    raise ValueError("hdrfilm: only pixel_format=\"rgb\" is inside the hot-path scope")
                for k in ("file_format", "component_format", "compensate"):
                    fp.pop(k, None)
                if fp:
                    raise ValueError(f"hdrfilm: unreferenced property {sorted(fp)}")
                for rf in ch:
                    if rf.tag == "rfilter":
                        f.rfilter = self.attr(rf, "type")
                        rp = self.props(rf)
                        f.rfilter_radius = rp.get("radius")
                        f.gaussian_stddev = float(rp.get("stddev", 0.5))
                        f.mitchell_b = float(rp.get("B", 1.0 / 3.0))
                        f.mitchell_c = float(rp.get("C", 1.0 / 3.0))
                        f.lanczos_lobes = int(rp.get("lobes", 3))
                        allowed = {"tent": {"radius"}, "gaussian": {"stddev"}, "mitchell": {"B", "C"}, "lanczos": {"lobes"}}
                        unknown = set(rp) - allowed.get(f.rfilter, set())
                        if unknown:
                            raise ValueError(f"rfilter '{f.rfilter}': unreferenced property {sorted(unknown)}")
                s.film = f
        s.fov = float(p.pop("fov", 45.0)) if "fov" in p else s.fov
        s.fov_axis = p.pop("fov_axis", "x")
        s.near_clip = float(p.pop("near_clip", 1e-2))
        s.far_clip = float(p.pop("far_clip", 1e4))
        s.shutter_open = float(p.pop("shutter_open", 0.0))
        s.shutter_close = float(p.pop("shutter_close", 0.0))
        if p:
            raise ValueError(f"perspective: unreferenced property {sorted(p)}")
        return s

    def load(self, root) -> Scene:
        if root.tag != "scene":
            raise ValueError("root element must be <scene>")
        sc = Scene()
        order = []
        for node in root:
            if node.tag == "default":
                k = node.get("name")
                if k not in self.cli:
                    self.defaults[k] = self.sub(node.get("value"))
            elif node.tag == "integrator":
                typ = self.attr(node, "type")
                if typ == "dopplertofpath":
                    sc.integrator = DopplerToFPathIntegrator(**self.props(node))
                elif typ == "velocity":
                    sc.integrator = VelocityIntegrator(**self.props(node))
                elif typ == "path":
                    sc.integrator = PathIntegrator(**self.props(node))
                else:
                    raise ValueError(f"integrator '{typ}' is outside the hot-path scope (dopplertofpath|velocity|path)")
            elif node.tag == "sensor":
                sc.sensor = self.sensor(node)
            elif node.tag == "bsdf":
                b = self.bsdf(node)
                if node.get("id"):
                    self.bsdfs[node.get("id")] = b
            elif node.tag == "shape":
                order.append(("shape", len(sc.shapes)))
                sc.shapes.append(self.shape(node))
            elif node.tag == "emitter":
                typ = self.attr(node, "type")
                if typ not in ("point", "constant", "spot", "directional"):
                    raise ValueError(f"emitter '{typ}' is outside the hot-path scope (point|spot|directional|area|constant)")
                p = self.props(node)
                if typ == "directional":   # DirectionalEmitter ctor, src/emitters/directional.cpp:65-91
                    unknown = set(p) - {"irradiance", "direction"}
                    if unknown:
                        raise ValueError(f"emitter 'directional': unreferenced property {sorted(unknown)}")
                    tw = None
                    for ch in node:
                        if ch.tag == "transform" and ch.get("name") == "to_world":
                            tw = self.transform(ch)
                    if "direction" in p:
                        if tw is not None:
                            raise ValueError("Only one of the parameters 'direction' and 'to_world' can be specified at the same time!'")
                        d = np.asarray(p["direction"], np.float32)
                        for _ in range(2):   # normalize(direction), then look_at(0, direction, up) normalises again; column 2 = d
                            d = d * (np.float32(1) / np.sqrt(np.float32(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])))
                    else:
                        m = np.asarray((tw or Transform4.identity()).matrix, np.float64).astype(np.float32)
                        d = m[:3, 2]                                                   # to_world.transform_affine((0, 0, 1))
                    order.append(("emitter", len(sc.emitters)))
                    sc.emitters.append(DirectionalLight(tuple(float(x) for x in d), p.get("irradiance", (1.0,) * 3)))
                    continue
                if typ == "spot":   # SpotLight ctor, src/emitters/spot.cpp:89-114
                    unknown = set(p) - {"intensity", "cutoff_angle", "beam_width"}
                    if unknown:
                        raise ValueError(f"emitter 'spot': unreferenced property {sorted(unknown)} (projection textures are out of scope)")
                    tw = Transform4.identity()
                    for ch in node:
                        if ch.tag == "transform" and ch.get("name") == "to_world":
                            tw = self.transform(ch)
                    order.append(("emitter", len(sc.emitters)))
                    sc.emitters.append(SpotLight(tw, p.get("intensity", (1.0,) * 3), float(p.get("cutoff_angle", 20.0)),
                                                 float(p["beam_width"]) if "beam_width" in p else None))
                    continue
                if typ == "constant":
                    unknown = set(p) - {"radiance"}
                    if unknown:
                        raise ValueError(f"emitter 'constant': unreferenced property {sorted(unknown)}")
                    order.append(("emitter", len(sc.emitters)))
                    sc.emitters.append(ConstantEmitter(p.get("radiance", (1.0,) * 3)))
                    continue
                pos = p.get("position")
                for ch in node:
                    if ch.tag == "transform" and ch.get("name") == "to_world":
                        if pos is not None:
                            raise ValueError("Only one of the parameters 'position' and 'to_world' can be specified")
                        pos = tuple(self.transform(ch).astype(np.float32).matrix[:3, 3].tolist())
                order.append(("emitter", len(sc.emitters)))
                sc.emitters.append(PointLight(pos if pos is not None else (0.0, 0.0, 0.0), p.get("intensity", (1.0,) * 3)))
            else:
                raise ValueError(f"unsupported top-level element <{node.tag}>")
        unused = self.cli - set(re.findall(r"\$(\w+)", ET.tostring(root, encoding="unicode")))
        if unused:   # xml.cpp:1069 'Unused parameter'
            raise ValueError(f"Unused parameter \"{sorted(unused)[0]}\"!")
        sc.scene_order = order
        if sc.integrator is None:
            raise ValueError("scene has no integrator")
        return sc


def load_file(path: str, **params) -> Scene:
    """`mi.load_file(path, **params)`: params override `<default>` values like `-Dkey=value`."""
    tree = ET.parse(path)
    return _Loader({k: str(v) for k, v in params.items()}, os.path.dirname(os.path.abspath(path))).load(tree.getroot())


# include/mitsuba/render/ior.h:23-49
_IOR = {"vacuum": 1.0, "helium": 1.000036, "hydrogen": 1.000132, "air": 1.000277, "carbon dioxide": 1.00045,
        "water": 1.3330, "acetone": 1.36, "ethanol": 1.361, "carbon tetrachloride": 1.461, "glycerol": 1.4729,
        "benzene": 1.501, "silicone oil": 1.52045, "bromine": 1.661, "water ice": 1.31, "fused quartz": 1.458,
        "pyrex": 1.470, "acrylic glass": 1.49, "polypropylene": 1.49, "bk7": 1.5046, "sodium chloride": 1.544,
        "amber": 1.55, "pet": 1.5750, "diamond": 2.419}


_CONDUCTORS = None


def conductor_ior(material: str):
    """RGB (eta, k) of a named conductor material: the table the reference derives from its measured spectra
    (complex_ior_from_file, include/mitsuba/render/ior.h:100-143), generated by tests/golden/make_conductor_table.py."""
    global _CONDUCTORS
    if _CONDUCTORS is None:
        import json
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "conductor_ior.json")) as f:
            _CONDUCTORS = json.load(f)
    if material not in _CONDUCTORS:
        raise ValueError(f'conductor material "{material}" is not in the table of named materials: ' + ", ".join(sorted(_CONDUCTORS)))
    v = _CONDUCTORS[material]
    return tuple(v["eta"]), tuple(v["k"])


def lookup_ior(value) -> float:
    """`lookup_ior` (include/mitsuba/render/ior.h:52-98): a number, or the name of a material of the table."""
    if isinstance(value, (int, float)):
        return float(value)
    name = str(value).strip().lower()
    try:
        return float(name)
    except ValueError:
        pass
    if name not in _IOR:
        raise ValueError(f'Unable to find an IOR value for "{name}"! Valid choices are: ' + ", ".join(_IOR))
    return _IOR[name]


def load_string(xml: str, base_dir: str = ".", **params) -> Scene:
    return _Loader({k: str(v) for k, v in params.items()}, base_dir).load(ET.fromstring(xml))
