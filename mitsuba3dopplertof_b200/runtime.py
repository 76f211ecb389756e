"""Loader + thin object wrapper for libdtof_b200.so (the C ABI in include/dtof.h).

The product path is the CUDA library only: if the shared object is missing, fails to load, or no CUDA
device is present, every entry point raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

from . import _abi
from .integrator import DTOFError

__all__ = ["load_library", "build_library", "scene_info", "Context", "get_context", "library_path"]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB: Optional[C.CDLL] = None
_CTX: Dict[int, "Context"] = {}


def library_path() -> str:
    # DTOF_LIB selects another build of the same library (kernel experiments); there is still no non-CUDA path
    return os.environ.get("DTOF_LIB") or os.path.join(_HERE, "libdtof_b200.so")


def build_library(verbose: bool = False) -> str:
    """Compile csrc/ for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode:
        raise DTOFError("building libdtof_b200.so failed")
    return library_path()


def load_library() -> C.CDLL:
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise DTOFError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
        lib = C.CDLL(path)
        _abi.bind(lib)   # AttributeError if a symbol of include/dtof.h is not exported
        if lib.dtof_abi_version() != _abi.ABI_VERSION:
            raise DTOFError("libdtof_b200.so ABI version mismatch")
        _LIB = lib
    return _LIB


def scene_info(scene_or_flat) -> _abi.SceneInfo:
    """Host half of the upload (validation, flattening, BVH build) -- needs the library but no GPU."""
    lib = load_library()
    flat = scene_or_flat.flatten() if hasattr(scene_or_flat, "flatten") else scene_or_flat
    info, err = _abi.SceneInfo(), C.create_string_buffer(512)
    rc = lib.dtof_scene_info_for(C.byref(flat.desc), C.byref(info), err, 512)
    if rc != _abi.OK:
        msg = err.value.decode("utf-8", "replace")
        if rc == _abi.ERR_INVALID:
            raise ValueError(msg)
        raise DTOFError(f"[status {rc}] {msg}")
    return info


class Context:
    """One dtof_ctx: one GPU, or -- `devices=[0, 1, ...]` -- one context over several GPUs of the node (dtof_create_multi:
    the scene is replicated, one render is sharded over the devices and summed on devices[0]). Thread-compatible, like
    the C ABI."""

    def __init__(self, device: int = 0, devices=None):
        self.lib = load_library()
        h = C.c_void_p()
        if devices is not None:
            devices = [int(d) for d in devices]
            self.device = devices[0]
            arr = (C.c_int * len(devices))(*devices)
            rc = self.lib.dtof_create_multi(C.byref(h), arr, len(devices))
        else:
            self.device = int(device)
            rc = self.lib.dtof_create(C.byref(h), self.device)
        if rc != _abi.OK:
            raise DTOFError(f"dtof_create(device={devices or device}) failed with status {rc}: no usable CUDA device "
                            "(the product path has no CPU fallback)")
        self.h = h
        self._flat = None

    def device_count(self) -> int:
        return int(self.lib.dtof_device_count(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.dtof_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:   # noqa: BLE001
            pass

    def _check(self, rc: int):
        if rc != _abi.OK:
            msg = self.lib.dtof_last_error(self.h).decode("utf-8", "replace")
            if rc == _abi.ERR_INVALID:
                raise ValueError(msg)
            raise DTOFError(f"[status {rc}] {msg}")

    # ---- scene ---------------------------------------------------------------------------------
    def upload(self, scene_or_flat):
        flat = scene_or_flat.flatten() if hasattr(scene_or_flat, "flatten") else scene_or_flat
        self._check(self.lib.dtof_upload_scene(self.h, C.byref(flat.desc)))
        self._flat = flat
        return flat

    def update_instances(self, first: int, instances) -> None:
        arr = (_abi.Instance * len(instances))(*instances)
        self._check(self.lib.dtof_update_instances(self.h, first, len(instances), arr))

    def pass_info(self, params: _abi.Params) -> _abi.PassInfo:
        pi = _abi.PassInfo()
        self._check(self.lib.dtof_pass_info_for(self.h, C.byref(params), C.byref(pi)))
        return pi

    # ---- rendering -------------------------------------------------------------------------------
    def render(self, flat, params: _abi.Params, develop: bool = True, both: bool = False):
        """Host buffers in/out (copies inside the call): the end-to-end path a plugin uses."""
        h, w = flat.height, flat.width
        rgbw = np.empty((h, w, 4), np.float32) if (both or not develop) else None
        img = np.empty((h, w, 3), np.float32) if (both or develop) else None
        self._check(self.lib.dtof_render(self.h, C.byref(params), _abi.as_fp(rgbw) if rgbw is not None else None,
                                         _abi.as_fp(img) if img is not None else None))
        if both:
            return img, rgbw
        return img if develop else rgbw

    def render_chunked(self, flat, params: _abi.Params, chunk_lanes: int, should_stop=None, develop: bool = True, both: bool = False):
        """dtof_render_accumulate over lane chunks + dtof_read_film: what the Mitsuba plugin does so that cancel() / the
        `timeout` property can stop a render between chunks. Returns like render()."""
        pi = self.pass_info(params)
        begin = int(params.lane_begin)
        end = int(params.lane_end) or int(pi.wavefront_size)
        chunk_lanes = max(int(pi.spp_per_pass), int(chunk_lanes) // int(pi.spp_per_pass) * int(pi.spp_per_pass))   # whole pixels
        first = True
        while begin < end and not (should_stop and should_stop()):
            p = _abi.Params.from_buffer_copy(params)
            p.lane_begin, p.lane_end = begin, min(end, begin + chunk_lanes)
            self._check(self.lib.dtof_render_accumulate(self.h, C.byref(p), int(first)))
            first = False
            begin += chunk_lanes
        h, w = flat.height, flat.width
        rgbw = np.zeros((h, w, 4), np.float32) if (both or not develop) else None
        img = np.zeros((h, w, 3), np.float32) if (both or develop) else None
        if not first:
            self._check(self.lib.dtof_read_film(self.h, _abi.as_fp(rgbw) if rgbw is not None else None,
                                                _abi.as_fp(img) if img is not None else None))
        if both:
            return img, rgbw
        return img if develop else rgbw

    def render_multi_pass(self, flat, params: _abi.Params, n_renders: int) -> np.ndarray:
        """Mean of `n_renders` developed renders with seed, seed+1, ... (averaged on the device)."""
        img = np.empty((flat.height, flat.width, 3), np.float32)
        self._check(self.lib.dtof_render_multi_pass(self.h, C.byref(params), int(n_renders), _abi.as_fp(img)))
        return img

    def render_device(self, params: _abi.Params, d_rgbw_ptr: int, stream_ptr: int = 0) -> None:
        """Accumulate into a caller-owned device tensor (e.g. torch.Tensor.data_ptr()), asynchronously."""
        self._check(self.lib.dtof_render_device(self.h, C.byref(params), C.c_void_p(d_rgbw_ptr), C.c_void_p(stream_ptr)))

    def develop_device(self, d_rgbw_ptr: int, d_img_ptr: int, stream_ptr: int = 0) -> None:
        self._check(self.lib.dtof_develop_device(self.h, C.c_void_p(d_rgbw_ptr), C.c_void_p(d_img_ptr), C.c_void_p(stream_ptr)))

    def trace_samples(self, params: _abi.Params, lanes, pass_index: int = 0) -> np.ndarray:
        """Per-lane records of pass `pass_index` (the streams of a lane persist across passes: earlier ones are replayed)."""
        lanes = np.ascontiguousarray(lanes, np.uint64)
        out = np.zeros(lanes.size, _abi.SAMPLE_RECORD_DTYPE)
        self._check(self.lib.dtof_trace_samples_pass(self.h, C.byref(params), lanes.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                     lanes.size, int(pass_index), out.ctypes.data_as(C.POINTER(_abi.SampleRecord))))
        return out

    def trace_rays(self, rays: np.ndarray, any_hit: bool = False) -> np.ndarray:
        """Scene::ray_intersect_preliminary / ray_test for caller-supplied rays (structured arrays _abi.RAY_DTYPE ->
        _abi.RAY_HIT_DTYPE)."""
        rays = np.ascontiguousarray(rays, _abi.RAY_DTYPE)
        out = np.zeros(rays.size, _abi.RAY_HIT_DTYPE)
        self._check(self.lib.dtof_trace_rays(self.h, rays.ctypes.data_as(C.POINTER(_abi.Ray)), rays.size, int(any_hit),
                                             out.ctypes.data_as(C.POINTER(_abi.RayHit))))
        return out

    # ---- instrumentation -----------------------------------------------------------------------
    def set_stats(self, enabled: bool) -> None:
        self._check(self.lib.dtof_set_stats(self.h, int(enabled)))

    def stats(self) -> _abi.Stats:
        st = _abi.Stats()
        self._check(self.lib.dtof_get_stats(self.h, C.byref(st)))
        return st

    def last_traversal_mode(self) -> int:
        return int(self.lib.dtof_last_traversal_mode(self.h))

    def last_pipeline(self) -> int:
        """0 = fused kernel, 1 = wavefront pipeline (include/dtof.h: dtof_last_pipeline)"""
        return int(self.lib.dtof_last_pipeline(self.h))

    def launch_count(self) -> int:
        return int(self.lib.dtof_launch_count(self.h))

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        self._check(self.lib.dtof_last_kernel_ms(self.h, C.byref(ms)))
        return float(ms.value)


def get_context(device: Optional[int] = None) -> Context:
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _CTX:
        _CTX[device] = Context(device)
    return _CTX[device]
