"""Procedural stand-in for BASELINE config 5 ("large multi-million-triangle motion-blurred scene").

The tutorial scenes are not in the reference repository (SURVEY.md finding 8), so the large scene is generated in
memory, deterministically: a cube-sphere whose radius is displaced by a seeded sum of sinusoids (a function of the
direction only, so the six faces meet without cracks), flat shaded, placed as ONE animated instance (translation +
small rotation over the exposure) inside the static slab room of tests/scenes/c5_slabroom.xml.
"""
from __future__ import annotations

import numpy as np

from .scene import Scene, Shape
from .transform import AnimatedTransform, Transform4

__all__ = ["displaced_cube_sphere", "large_scene"]


def displaced_cube_sphere(n: int, seed: int = 1234, radius: float = 1.0, amplitude: float = 0.08, octaves: int = 6):
    """6 faces x n x n quads x 2 = 12 n^2 triangles. Returns (positions float32 (V,3), faces uint32 (F,3))."""
    rng = np.random.RandomState(seed)
    freq = rng.uniform(2.0, 40.0, size=(octaves, 3))
    phase = rng.uniform(0.0, 2.0 * np.pi, size=octaves)
    amp = amplitude / (1.0 + np.arange(octaves))
    u = np.linspace(-1.0, 1.0, n + 1)
    a, b = np.meshgrid(u, u, indexing="xy")
    one = np.ones_like(a)
    face_dirs = [(one, a, b), (-one, b, a), (b, one, a), (a, -one, b), (a, b, one), (b, a, -one)]
    pos, faces = [], []
    i = np.arange(n)
    ii, jj = np.meshgrid(i, i, indexing="xy")
    v00 = (jj * (n + 1) + ii).ravel()
    quad = np.stack([v00, v00 + 1, v00 + n + 2, v00, v00 + n + 2, v00 + n + 1], axis=1).reshape(-1, 3)
    for k, (x, y, z) in enumerate(face_dirs):
        d = np.stack([x, y, z], axis=-1).reshape(-1, 3)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        r = np.full(d.shape[0], radius)
        for o in range(octaves):
            r += amp[o] * np.sin(d @ freq[o] + phase[o])
        pos.append((d * r[:, None]).astype(np.float32))
        faces.append((quad + k * (n + 1) * (n + 1)).astype(np.uint32))
    return np.concatenate(pos), np.concatenate(faces)


def large_scene(base: Scene, n: int = 592, seed: int = 1234) -> Scene:
    """`base` = the loaded slab room (its movers are kept); adds the displaced sphere (12 n^2 triangles, n = 592 ->
    4.2 M) as an animated instance: centre (0, 0.75, 0), radius 0.45, moving +0.02 in x and turning 1 degree about y
    between t = 0 and t = T (two keyframes; the motion model is the reference's linear matrix interpolation)."""
    pos, faces = displaced_cube_sphere(n, seed)
    T = float(base.integrator.time)
    k0 = Transform4.translate((0.0, 0.75, 0.0)) @ Transform4.scale((0.45, 0.45, 0.45))
    k1 = Transform4.translate((0.02, 0.75, 0.0)) @ Transform4.rotate((0.0, 1.0, 0.0), 1.0) @ Transform4.scale((0.45, 0.45, 0.45))
    at = AnimatedTransform()
    at.append(0.0, k0)
    at.append(T, k1)
    bsdf = next((s.bsdf for s in base.shapes if s.bsdf is not None), None)
    sphere = Shape("mesh", to_world=at, bsdf=bsdf, positions=pos, faces=faces, id="DisplacedSphere")
    base.shapes.append(sphere)
    if getattr(base, "scene_order", None) is not None:
        base.scene_order.append(("shape", len(base.shapes) - 1))
    return base
