"""`dopplertofpath` integrator: the reference's plugin surface over the CUDA library.

Property names, defaults and derived values follow
  src/integrators/dopplertofpath.cpp:19-57      (time, w_g, g_1, g_0, w_s, sensor_phase_offset, hetero_offset,
                                                 hetero_frequency, wave_function_type, low_frequency_component_only)
  src/render/integrator.cpp:24-27,54-100,568-585 (timeout, hide_emitters, is_doppler_integrator,
                                                 time_sampling_method, antithetic_shift,
                                                 use_stratified_sampling_for_each_interval, path_correlation_depth,
                                                 block_size, samples_per_pass, max_depth, rr_depth)
`render(scene, seed=0, spp=0, develop=True)` mirrors `Integrator::render` (src/render/integrator.cpp:104-347;
python usage doppler_tutorials/src/program_runner.py:11-31). Unknown `time_sampling_method` /
`wave_function_type` strings raise (the reference leaves the enum uninitialised).
Unknown properties raise like the XML loader's "unreferenced property" check (src/core/xml.cpp:1204-1223).

The product path is CUDA only: `render` raises if libdtof_b200.so or a GPU is missing.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np

from . import _abi

__all__ = ["DopplerToFPathIntegrator", "VelocityIntegrator", "PathIntegrator", "DTOFError"]

f32 = np.float32

_TIME = {"uniform": _abi.TIME_UNIFORM, "stratified": _abi.TIME_STRATIFIED, "antithetic": _abi.TIME_ANTITHETIC,
         "antithetic_mirror": _abi.TIME_ANTITHETIC_MIRROR}
_WAVE = {"sinusoidal": _abi.WAVE_SINUSOIDAL, "rectangular": _abi.WAVE_RECTANGULAR,
         "triangular": _abi.WAVE_TRIANGULAR, "trapezoidal": _abi.WAVE_TRAPEZOIDAL}

_KNOWN = {
    "time", "w_g", "g_1", "g_0", "w_s", "sensor_phase_offset", "hetero_offset", "hetero_frequency",
    "wave_function_type", "low_frequency_component_only", "max_depth", "rr_depth", "hide_emitters", "timeout",
    "is_doppler_integrator", "time_sampling_method", "antithetic_shift", "use_stratified_sampling_for_each_interval",
    "path_correlation_depth", "block_size", "samples_per_pass",
}


class DTOFError(RuntimeError):
    pass


class DopplerToFPathIntegrator:
    KIND = _abi.INTEGRATOR_DOPPLERTOFPATH

    def __init__(self, **props):
        unknown = set(props) - _KNOWN
        if unknown:
            raise ValueError(f"dopplertofpath: unreferenced propert{'ies' if len(unknown) > 1 else 'y'} {sorted(unknown)}")
        g = props.get
        # --- DopplerToFPathIntegrator ctor (all ScalarFloat = float32)
        self.time = f32(g("time", 0.0015))
        self.w_g = f32(g("w_g", 30.0))
        self.g_1 = f32(g("g_1", 0.5))
        self.g_0 = f32(g("g_0", 0.5))
        self.w_s = f32(g("w_s", 30.0))
        self.sensor_phase_offset = f32(g("sensor_phase_offset", 0.0))
        if "hetero_offset" in props:      # :32-34  (float * int * double -> double -> float)
            self.sensor_phase_offset = f32(float(f32(props["hetero_offset"])) * 2 * math.pi)
        if "hetero_frequency" in props:   # :35-37
            self.hetero_frequency = f32(props["hetero_frequency"])
            self.w_s = f32(float(self.w_g) + float(f32(self.hetero_frequency / self.time)) * 1e-6)
        else:                             # :38-40
            self.hetero_frequency = f32(float(f32(self.w_s - self.w_g)) * 1e6 * float(self.time))
        wave = g("wave_function_type", "sinusoidal")
        if wave not in _WAVE:
            raise ValueError(f"dopplertofpath: unknown wave_function_type '{wave}'")
        self.wave_function_type = wave
        self.low_frequency_component_only = bool(g("low_frequency_component_only", True))
        # --- SamplingIntegrator ctor
        self.is_doppler_integrator = True
        method = g("time_sampling_method", "antithetic")
        if method not in _TIME:
            raise ValueError(f"dopplertofpath: unknown time_sampling_method '{method}'")
        self.time_sampling_method = method
        default_shift = 0.5 if method == "antithetic" else 0.0
        self.antithetic_shift = f32(g("antithetic_shift", default_shift))
        self.use_stratified_sampling_for_each_interval = bool(g("use_stratified_sampling_for_each_interval", True))
        self.path_correlation_depth = int(g("path_correlation_depth", 0))
        self.block_size = int(g("block_size", 0))
        self.samples_per_pass = g("samples_per_pass", None)
        if self.samples_per_pass is not None:
            raise ValueError("'samples_per_pass' is deprecated in the reference and unsupported here")
        # --- MonteCarloIntegrator ctor (:568-585)
        max_depth = int(g("max_depth", -1))
        if max_depth < 0 and max_depth != -1:
            raise ValueError("\"max_depth\" must be set to -1 (infinite) or a value >= 0")
        self.max_depth = max_depth
        rr_depth = int(g("rr_depth", 5))
        if rr_depth <= 0:
            raise ValueError("\"rr_depth\" must be set to a value greater than zero!")
        self.rr_depth = rr_depth
        # --- Integrator ctor
        self.hide_emitters = bool(g("hide_emitters", False))
        self.timeout = float(g("timeout", -1.0))
        self._stop = False

    # ---------------------------------------------------------------------------------------
    def params(self, sampler, seed: int = 0, spp: int = 0, lane_begin: int = 0, lane_end: int = 0) -> _abi.Params:
        """Pack the C-ABI parameter block for one render() call."""
        p = _abi.Params()
        p.time, p.w_g, p.g_1, p.g_0 = float(self.time), float(self.w_g), float(self.g_1), float(self.g_0)
        p.sensor_phase_offset = float(self.sensor_phase_offset)
        p.hetero_frequency = float(self.hetero_frequency)
        p.wave_function_type = _WAVE[self.wave_function_type]
        p.low_frequency_component_only = int(self.low_frequency_component_only)
        p.max_depth, p.rr_depth, p.hide_emitters = self.max_depth, self.rr_depth, int(self.hide_emitters)
        p.time_sampling_method = _TIME[self.time_sampling_method]
        p.antithetic_shift = float(self.antithetic_shift)
        p.use_stratified_sampling_for_each_interval = int(self.use_stratified_sampling_for_each_interval)
        p.path_correlation_depth = self.path_correlation_depth
        p.sample_count = int(spp) if spp else int(sampler.sample_count)   # integrator.cpp:121-124
        p.base_seed = int(sampler.seed) & 0xFFFFFFFF
        p.time_correlate_number = int(sampler.time_correlate_number)
        p.path_correlate_number = int(sampler.path_correlate_number)
        p.seed = int(seed) & 0xFFFFFFFF
        p.lane_begin, p.lane_end = int(lane_begin), int(lane_end)
        p.integrator = self.KIND
        if self.KIND == _abi.INTEGRATOR_DOPPLERTOFPATH and getattr(sampler, "kind", "correlated") != "correlated":
            # Any other sampler answers the Doppler branch's calls with the base-class defaults (include/mitsuba/render/
            # sampler.h:131-144): next_1d_time, next_1d_correlate and next_2d_correlate all draw from the one independent
            # stream whatever strategy / correlate flag is passed. That is the correlated sampler's `rng` stream under
            # uniform time sampling with no path correlation (correlated.cpp:92-97, 156-161): same seeding, same order.
            p.time_sampling_method = _TIME["uniform"]
            p.use_stratified_sampling_for_each_interval = 0
            p.path_correlation_depth = 0
            p.time_correlate_number = p.path_correlate_number = 1
            return p
        if self.time_sampling_method == "antithetic_mirror" and p.time_correlate_number != 2:
            raise ValueError("antithetic_mirror requires time_correlate_number == 2")  # correlated.cpp:141-142
        if p.time_correlate_number < 1 or p.path_correlate_number < 1:
            raise ValueError("correlate numbers must be >= 1")
        return p

    def cancel(self) -> None:
        self._stop = True

    def should_stop(self) -> bool:
        return self._stop

    def render(self, scene, seed: int = 0, spp: int = 0, develop: bool = True, evaluate: bool = True,
               device: Optional[int] = None) -> np.ndarray:
        """Render `scene` (a scene.Scene) on the GPU; returns the developed (H, W, 3) image, or the raw
        (H, W, 4) RGBW accumulation tensor when ``develop=False``. Host buffers in, host buffers out."""
        from .runtime import get_context   # deferred: importing the package must not need a GPU
        self._stop = False
        ctx = get_context(device)
        flat = _uploaded(ctx, scene)   # flatten + BVH build + H2D once per scene STATE (Scene.fingerprint)
        p = self.params(scene.sensor.sampler, seed, spp)
        return ctx.render(flat, p, develop=develop)

    def render_multi_pass(self, scene, spp_per_pass: int, passes: int, seed: int = 0, device: Optional[int] = None) -> np.ndarray:
        """`render_image_multi_pass(scene, integrator, spp, passes)` of the tutorials
        (doppler_tutorials/src/program_runner.py:11-31): mean over `passes` renders with seed, seed+1, ...;
        the scene stays on the GPU and the mean is formed there."""
        from .runtime import get_context
        ctx = get_context(device)
        flat = _uploaded(ctx, scene)
        return ctx.render_multi_pass(flat, self.params(scene.sensor.sampler, seed, spp_per_pass), passes)

    def __repr__(self):
        return (f"DopplerToFPathIntegrator[\n  max_depth = {self.max_depth & 0xFFFFFFFF},\n"
                f"  rr_depth = {self.rr_depth}\n]")


def _uploaded(ctx, scene):
    """The flattened scene resident on `ctx`, uploaded again when the scene changed since (shapes, transforms, materials,
    emitters, the sensor, its film or sampler: Scene.fingerprint) or another scene took the context over."""
    key = scene.fingerprint()
    cached = getattr(scene, "_dtof_uploaded", None)
    if cached is None or cached[0] is not ctx or ctx._flat is not cached[1] or cached[2] != key:
        scene._dtof_uploaded = (ctx, ctx.upload(scene), key)
    return scene._dtof_uploaded[1]


class VelocityIntegrator(DopplerToFPathIntegrator):
    """`velocity`: the reference's ground-truth radial-velocity integrator (src/integrators/velocity.cpp:87-127;
    used by doppler_tutorials/src/program_runner.py:33-54). Properties: `time` (0.0015) plus the SamplingIntegrator /
    MonteCarloIntegrator ones; it is not a Doppler integrator, so pixel jitter and time come from the sampler's
    independent stream (render_sample stock branch, src/render/integrator.cpp:409-472). Every channel of the image is
    (t(time) - t(0)) / time along the camera ray, 0 where either query misses."""
    KIND = _abi.INTEGRATOR_VELOCITY
    _VELOCITY_PROPS = {"time", "max_depth", "rr_depth", "hide_emitters", "timeout", "block_size", "samples_per_pass",
                       "time_sampling_method", "antithetic_shift", "use_stratified_sampling_for_each_interval",
                       "path_correlation_depth", "is_doppler_integrator"}

    def __init__(self, **props):
        unknown = set(props) - self._VELOCITY_PROPS
        if unknown:
            raise ValueError(f"velocity: unreferenced propert{'ies' if len(unknown) > 1 else 'y'} {sorted(unknown)}")
        super().__init__(**props)
        self.is_doppler_integrator = bool(props.get("is_doppler_integrator", False))
        if self.is_doppler_integrator:
            raise ValueError("velocity with is_doppler_integrator=true is outside the hot-path scope")

    def __repr__(self):
        return (f"VelocityIntegrator[\n  max_depth = {self.max_depth & 0xFFFFFFFF},\n"
                f"  rr_depth = {self.rr_depth}\n]")


class PathIntegrator(DopplerToFPathIntegrator):
    """`path`: the stock path tracer (src/integrators/path.cpp:103-283) the tutorials render the radiance pass with
    (doppler_tutorials/src/program_runner.py:57-80). The bounce loop is the one `dopplertofpath` was derived from: no
    modulation weight, no time wrap, every draw is Sampler::next_1d / next_2d, i.e. the sampler's independent stream
    (identical for `independent` and `correlated`, src/samplers/correlated.cpp:78-90), and render_sample takes the
    stock branch (src/render/integrator.cpp:409-472). Properties: those of MonteCarloIntegrator / SamplingIntegrator;
    `time`, `w_g`, ... are not read by the plugin and therefore rejected like the XML loader does."""
    KIND = _abi.INTEGRATOR_PATH
    _PATH_PROPS = {"max_depth", "rr_depth", "hide_emitters", "timeout", "block_size", "samples_per_pass",
                   "time_sampling_method", "antithetic_shift", "use_stratified_sampling_for_each_interval",
                   "path_correlation_depth", "is_doppler_integrator"}

    def __init__(self, **props):
        unknown = set(props) - self._PATH_PROPS
        if unknown:
            raise ValueError(f"path: unreferenced propert{'ies' if len(unknown) > 1 else 'y'} {sorted(unknown)}")
        super().__init__(**props)
        self.is_doppler_integrator = bool(props.get("is_doppler_integrator", False))
        if self.is_doppler_integrator:
            raise ValueError("path with is_doppler_integrator=true is outside the hot-path scope")

    def __repr__(self):
        return (f"PathIntegrator[\n  max_depth = {self.max_depth & 0xFFFFFFFF},\n"
                f"  rr_depth = {self.rr_depth}\n]")
