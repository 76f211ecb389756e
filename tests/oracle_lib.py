"""ctypes wrapper around oracle/libdtof_oracle.so (TEST INFRASTRUCTURE; never imported by the package)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from mitsuba3dopplertof_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "libdtof_oracle.so"], check=True)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(ORACLE_DIR, "libdtof_oracle.so")
        src = os.path.join(ORACLE_DIR, "dtof_oracle.cpp")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        L = C.CDLL(path)
        u32, u64, f, fp = C.c_uint32, C.c_uint64, C.c_float, C.POINTER(C.c_float)
        sig = {
            "dtof_oracle_tea32": (None, [u32, u32, C.c_int, C.POINTER(u32)]),
            "dtof_oracle_pcg32": (None, [u64, u64, u32, C.POINTER(u32), fp, C.POINTER(u64)]),
            "dtof_oracle_permute_kensler": (u32, [u32, u32, u32]),
            "dtof_oracle_sincos": (None, [f, fp, fp]),
            "dtof_oracle_rfilter_eval": (f, [C.POINTER(_abi.Film), f]),
            "dtof_oracle_waveform_lowpass": (f, [f, u32]),
            "dtof_oracle_waveform": (f, [f, u32]),
            "dtof_oracle_modulation_weight": (f, [C.POINTER(_abi.Params), f, f]),
            "dtof_oracle_square_to_cosine_hemisphere": (None, [f, f, fp]),
            "dtof_oracle_coordinate_system": (None, [fp, fp, fp]),
            "dtof_oracle_seed_lane": (None, [C.POINTER(_abi.Params), u64, u32, C.POINTER(u64)]),
            "dtof_oracle_time_sample": (f, [C.POINTER(_abi.Params), u64, u32, u32]),
            "dtof_oracle_camera_ray": (None, [C.POINTER(_abi.Camera), f, f, fp, fp, fp]),
            "dtof_oracle_film_put": (None, [C.POINTER(_abi.Film), f, f, fp, C.POINTER(C.c_double)]),
            "dtof_oracle_pass_info": (C.c_int, [C.POINTER(_abi.SceneDesc), C.POINTER(_abi.Params), C.POINTER(_abi.PassInfo)]),
            "dtof_oracle_scene_create": (C.c_void_p, [C.POINTER(_abi.SceneDesc), C.c_int]),
            "dtof_oracle_scene_destroy": (None, [C.c_void_p]),
            "dtof_oracle_trace_samples": (C.c_int, [C.c_void_p, C.POINTER(_abi.Params), C.POINTER(u64), u32,
                                                    C.POINTER(_abi.SampleRecord)]),
            "dtof_oracle_trace_samples_pass": (C.c_int, [C.c_void_p, C.POINTER(_abi.Params), C.POINTER(u64), u32, u32,
                                                         C.POINTER(_abi.SampleRecord)]),
            "dtof_oracle_trace_rays": (C.c_int, [C.c_void_p, C.POINTER(_abi.Ray), u32, C.c_int, C.POINTER(_abi.RayHit)]),
            "dtof_oracle_render": (C.c_int, [C.c_void_p, C.POINTER(_abi.Params), C.c_int, fp, fp]),
            "dtof_oracle_get_stats": (None, [C.c_void_p, C.POINTER(_abi.Stats)]),
        }
        _abi.bind(L, sig)
        _LIB = L
    return _LIB


class OracleScene:
    def __init__(self, flat, use_bvh: int = -1):
        self.flat = flat
        self.h = lib().dtof_oracle_scene_create(C.byref(flat.desc), use_bvh)

    def __del__(self):
        if getattr(self, "h", None):
            lib().dtof_oracle_scene_destroy(self.h)
            self.h = None

    def trace(self, params, lanes, pass_index: int = 0) -> np.ndarray:
        lanes = np.ascontiguousarray(lanes, np.uint64)
        out = np.zeros(lanes.size, _abi.SAMPLE_RECORD_DTYPE)
        rc = lib().dtof_oracle_trace_samples_pass(self.h, C.byref(params), lanes.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                  lanes.size, pass_index, out.ctypes.data_as(C.POINTER(_abi.SampleRecord)))
        if rc:
            raise RuntimeError(f"oracle trace_samples failed: {rc}")
        return out

    def trace_rays(self, rays, any_hit: bool = False) -> np.ndarray:
        rays = np.ascontiguousarray(rays, _abi.RAY_DTYPE)
        out = np.zeros(rays.size, _abi.RAY_HIT_DTYPE)
        rc = lib().dtof_oracle_trace_rays(self.h, rays.ctypes.data_as(C.POINTER(_abi.Ray)), rays.size, int(any_hit),
                                          out.ctypes.data_as(C.POINTER(_abi.RayHit)))
        if rc:
            raise RuntimeError(f"oracle trace_rays failed: {rc}")
        return out

    def render(self, params, n_threads: int = 0, develop: bool = True):
        h, w = self.flat.height, self.flat.width
        rgbw = np.zeros((h, w, 4), np.float32)
        img = np.zeros((h, w, 3), np.float32)
        rc = lib().dtof_oracle_render(self.h, C.byref(params), n_threads, _abi.as_fp(rgbw), _abi.as_fp(img))
        if rc:
            raise RuntimeError(f"oracle render failed: {rc}")
        return (img if develop else rgbw)

    def stats(self) -> _abi.Stats:
        st = _abi.Stats()
        lib().dtof_oracle_get_stats(self.h, C.byref(st))
        return st
