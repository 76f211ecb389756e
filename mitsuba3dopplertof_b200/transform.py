"""Host-side transforms mirroring the reference's ``Transform4f`` / ``AnimatedTransform``.

Reference: include/mitsuba/core/transform.h:24-380 (Transform: matrix + tracked inverse transpose),
:382-551 (AnimatedTransform: linear 4x4 matrix interpolation between keyframes 0 and 1, :440-466),
src/core/transform.cpp:22-36 (append).  XML transforms are composed in double precision
(``Properties::Float = double``, include/mitsuba/core/properties.h:55; src/core/xml.cpp:820-1007) and
narrowed to float32 when a plugin reads them; sensor-internal transforms (perspective projection)
are composed in float32 with Dr.Jit's fused multiply-add chains, which `_matmul_f32` emulates.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

__all__ = ["Transform4", "AnimatedTransform", "perspective_projection"]


def _fma32(a, b, c):
    """float32 fused multiply-add emulated through float64 (a*b is exact in binary64)."""
    return np.float32(np.float64(a) * np.float64(b) + np.float64(c))


def _matmul_f32(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Dr.Jit Matrix4f product (ext/drjit/include/drjit/matrix.h): per output column j,
    sum = a.col(0) * b(0,j); sum = fmadd(a.col(i), b(i,j), sum)."""
    out = np.zeros((4, 4), dtype=np.float32)
    for j in range(4):
        for r in range(4):
            s = np.float32(a[r, 0] * b[0, j])
            for i in range(1, 4):
                s = _fma32(a[r, i], b[i, j], s)
            out[r, j] = s
    return out


@dataclass
class Transform4:
    """A 4x4 transform with its tracked inverse transpose, in ``dtype`` precision."""

    matrix: np.ndarray = field(default_factory=lambda: np.eye(4))
    inverse_transpose: np.ndarray = field(default_factory=lambda: np.eye(4))

    # ---- construction ---------------------------------------------------------------------
    @staticmethod
    def identity(dtype=np.float64) -> "Transform4":
        return Transform4(np.eye(4, dtype=dtype), np.eye(4, dtype=dtype))

    @staticmethod
    def from_matrix(m: Sequence[float], dtype=np.float64) -> "Transform4":
        m = np.asarray(m, dtype=dtype).reshape(4, 4)
        inv_t = np.linalg.inv(m.astype(np.float64)).T.astype(dtype)
        return Transform4(m.copy(), inv_t)

    @staticmethod
    def translate(v, dtype=np.float64) -> "Transform4":
        m, it = np.eye(4, dtype=dtype), np.eye(4, dtype=dtype)
        m[:3, 3] = np.asarray(v, dtype=dtype)
        it[3, :3] = -np.asarray(v, dtype=dtype)
        return Transform4(m, it)

    @staticmethod
    def scale(v, dtype=np.float64) -> "Transform4":
        v = np.asarray(v, dtype=dtype)
        m, it = np.eye(4, dtype=dtype), np.eye(4, dtype=dtype)
        m[0, 0], m[1, 1], m[2, 2] = v
        it[0, 0], it[1, 1], it[2, 2] = (dtype(1) / v)
        return Transform4(m, it)

    @staticmethod
    def rotate(axis, angle_deg: float, dtype=np.float64) -> "Transform4":
        # dr::rotate<Matrix4>(axis, angle): Rodrigues with a normalised axis
        a = np.asarray(axis, dtype=np.float64)
        a = a / np.linalg.norm(a)
        ang = math.radians(float(angle_deg))
        s, c = math.sin(ang), math.cos(ang)
        x, y, z = a
        m = np.eye(4, dtype=np.float64)
        m[:3, :3] = [
            [c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s],
            [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s],
            [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c)],
        ]
        m = m.astype(dtype)
        return Transform4(m, m.copy())

    @staticmethod
    def look_at(origin, target, up, dtype=np.float64) -> "Transform4":
        o = np.asarray(origin, dtype=np.float64)
        t = np.asarray(target, dtype=np.float64)
        u = np.asarray(up, dtype=np.float64)
        d = (t - o) / np.linalg.norm(t - o)
        left = np.cross(u, d)
        left /= np.linalg.norm(left)
        new_up = np.cross(d, left)
        m = np.eye(4)
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, new_up, d, o
        return Transform4.from_matrix(m, dtype)

    # ---- algebra -----------------------------------------------------------------------------
    def __matmul__(self, other: "Transform4") -> "Transform4":
        if self.matrix.dtype == np.float32:
            return Transform4(_matmul_f32(self.matrix, other.matrix),
                              _matmul_f32(self.inverse_transpose, other.inverse_transpose))
        return Transform4(self.matrix @ other.matrix, self.inverse_transpose @ other.inverse_transpose)

    def inverse(self) -> "Transform4":
        # transform.h:70-72: just shuffles
        return Transform4(self.inverse_transpose.T.copy(), self.matrix.T.copy())

    def astype(self, dtype) -> "Transform4":
        return Transform4(self.matrix.astype(dtype), self.inverse_transpose.astype(dtype))

    # ---- application (float32, fma chains as in transform.h:96-150) ----------------------------
    def transform_affine_point(self, p) -> np.ndarray:
        m = self.matrix.astype(np.float32)
        out = np.empty(3, dtype=np.float32)
        for r in range(3):
            a = m[r, 3]
            for i in range(3):
                a = _fma32(m[r, i], np.float32(p[i]), a)
            out[r] = a
        return out

    def transform_vector(self, v) -> np.ndarray:
        m = self.matrix.astype(np.float32)
        out = np.empty(3, dtype=np.float32)
        for r in range(3):
            a = np.float32(m[r, 0] * np.float32(v[0]))
            for i in range(1, 3):
                a = _fma32(m[r, i], np.float32(v[i]), a)
            out[r] = a
        return out

    def transform_normal(self, n) -> np.ndarray:
        m = self.inverse_transpose.astype(np.float32)
        out = np.empty(3, dtype=np.float32)
        for r in range(3):
            a = np.float32(m[r, 0] * np.float32(n[0]))
            for i in range(1, 3):
                a = _fma32(m[r, i], np.float32(n[i]), a)
            out[r] = a
        return out

    def has_scale(self) -> bool:
        # Transform::has_scale: squared column norms of the upper 3x3 differ from 1 by > 1e-3
        m = self.matrix.astype(np.float64)[:3, :3]
        return bool(np.any(np.abs((m * m).sum(axis=0) - 1.0) > 1e-3))

    def m34(self) -> np.ndarray:
        """Row-major 3x4 float32 block handed to the C ABI."""
        return np.ascontiguousarray(self.matrix.astype(np.float32)[:3, :4]).reshape(12)


@dataclass
class AnimatedTransform:
    """Keyframed transform. Only keyframes 0 and 1 take part in ``eval`` (transform.h:451-456)."""

    times: List[float] = field(default_factory=list)
    transforms: List[Transform4] = field(default_factory=list)

    def append(self, time: float, trafo: Transform4) -> None:
        # src/core/transform.cpp:22-36
        if self.times and time <= self.times[-1]:
            raise ValueError("AnimatedTransform::append(): time values must be strictly monotonically increasing!")
        self.times.append(float(np.float32(time)))
        self.transforms.append(trafo.astype(np.float32))

    def size(self) -> int:
        return len(self.times)

    def get_min_time(self) -> float:
        return min(self.times) if self.times else 100.0   # transform.h:503-511

    def get_max_time(self) -> float:
        return max(self.times) if self.times else -100.0  # transform.h:514-522

    def eval(self, time: float) -> np.ndarray:
        if self.size() <= 1:
            return self.transforms[0].matrix.astype(np.float32) if self.transforms else np.eye(4, dtype=np.float32)
        t0, t1 = np.float32(self.times[0]), np.float32(self.times[1])
        t = np.float32(min(max((np.float32(time) - t0) / (t1 - t0), np.float32(0)), np.float32(1)))
        m0, m1 = self.transforms[0].matrix, self.transforms[1].matrix
        return (m0 * (np.float32(1) - t) + m1 * t).astype(np.float32)


def perspective_projection(film_size, crop_size, crop_offset, fov_x: float, near: float, far: float) -> Transform4:
    """``perspective_projection<float>`` (include/mitsuba/render/sensor.h:227-262) with
    ``Transform::perspective`` (include/mitsuba/core/transform.h:216-233), all in float32."""
    f32 = np.float32
    fw, fh = f32(film_size[0]), f32(film_size[1])
    rel_size = (f32(crop_size[0]) / fw, f32(crop_size[1]) / fh)
    rel_off = (f32(crop_offset[0]) / fw, f32(crop_offset[1]) / fh)
    aspect = fw / fh
    near, far, fov_x = f32(near), f32(far), f32(fov_x)

    recip = f32(1) / (far - near)
    # dr::tan(dr::deg_to_rad(fov * .5f)) -- Dr.Jit's float tan polynomial; libm tanf agrees to <= 1 ulp
    tan = f32(math.tan(float(f32(fov_x * f32(0.5)) * f32(math.pi / 180.0))))
    cot = f32(1) / tan
    trafo = np.zeros((4, 4), dtype=f32)
    trafo[0, 0] = cot
    trafo[1, 1] = cot
    trafo[2, 2] = far * recip
    trafo[2, 3] = -near * far * recip
    trafo[3, 2] = f32(1)
    inv = np.zeros((4, 4), dtype=f32)
    inv[0, 0] = tan
    inv[1, 1] = tan
    inv[3, 3] = f32(1) / near
    inv[2, 3] = f32(1)
    inv[3, 2] = (near - far) / (far * near)
    persp = Transform4(trafo, inv.T.copy())

    def sc(v):
        return Transform4.scale(v, dtype=f32)

    def tr(v):
        return Transform4.translate(v, dtype=f32)

    return (sc((f32(1) / rel_size[0], f32(1) / rel_size[1], f32(1)))
            @ tr((-rel_off[0], -rel_off[1], f32(0)))
            @ sc((f32(-0.5), f32(-0.5) * aspect, f32(1)))
            @ tr((f32(-1), f32(-1) / aspect, f32(0)))
            @ persp)
