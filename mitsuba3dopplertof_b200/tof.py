"""Doppler-ToF image post-processing and the tutorial drivers on top of the B200 renderer.

The reference keeps these in its tutorial scripts, not in the renderer:
  * `doppler_tutorials/src/utils/image_utils.py:20-36`   luminance / ToF scaling of a rendered image
  * `doppler_tutorials/src/utils/image_utils.py:140-199` radial velocity from a homodyne / heterodyne pair
  * `doppler_tutorials/src/program_runner.py:11-31,33-160` multi-pass rendering of the velocity ground truth, the
    radiance pass and the Doppler-ToF measurement
The functions below keep the reference's names, argument meaning and numerical behaviour (float64 numpy on the host;
`tests/golden/tof_postprocess.npz` pins them against the reference's own functions). The drivers render through
`Context.render_multi_pass`, i.e. the multi-seed mean is accumulated on the device (`dtof_render_multi_pass`).
"""
from typing import Optional, Sequence

import numpy as np

from . import runtime
from .integrator import DopplerToFPathIntegrator, PathIntegrator, VelocityIntegrator

__all__ = ["rgb2luminance", "to_tof_image", "to_tof_image_0_5", "calc_velocity_from_homo_hetero",
           "calc_velocity_from_homo_heteros", "render_image_multi_pass", "run_scene_velocity", "run_scene_radiance",
           "run_scene_doppler_tof", "doppler_velocity_map", "update_motion"]

SPEED_OF_LIGHT = 3e8   # the tutorials' constant (image_utils.py:160), not 299 792 458


def rgb2luminance(img):
    """Rec. 709 luminance of an (H, W, 3) image (image_utils.py:20-21)."""
    img = np.asarray(img)
    return (0.2126 * img[:, :, 0]) + (0.7152 * img[:, :, 1]) + (0.0722 * img[:, :, 2])


def to_tof_image(img, exposure_time: float = 0.0015):
    """Luminance scaled by the exposure time: the integrated sensor response (image_utils.py:27-31)."""
    return rgb2luminance(np.asarray(img)) * exposure_time


def to_tof_image_0_5(img):
    """Luminance of `img - 1` (image_utils.py:33-36)."""
    return rgb2luminance(np.asarray(img) - 1)


def _velocity_from_ratio(ratio, exposure_time: float, w_g: float):
    # ratio = heterodyne / homodyne = dw T / (dw T - 1)  ->  dw = ratio / T / (ratio - 1); v = c dw / (2 w_g), sign flipped
    ratio = np.clip(ratio, -1, 0.999)
    delta_w = ratio * (1 / exposure_time) / (ratio - 1)
    velocity_map = 0.5 * delta_w * SPEED_OF_LIGHT / (w_g * 1e6)
    return -velocity_map


def calc_velocity_from_homo_hetero(homodyne, heterodyne, **kwargs):
    """Radial velocity map from one homodyne / heterodyne ToF image pair (image_utils.py:140-168).
    kwargs: exposure_time (0.0015), w_g in MHz (30). Pixels with homodyne == 0 give ratio 0."""
    homodyne = np.asarray(homodyne)
    heterodyne = np.asarray(heterodyne)
    ratio = np.divide(heterodyne, homodyne, out=np.zeros_like(homodyne), where=np.abs(homodyne) > 0)
    return _velocity_from_ratio(ratio, kwargs.get("exposure_time", 0.0015), kwargs.get("w_g", 30))


def calc_velocity_from_homo_heteros(homodynes: Sequence, heterodynes: Sequence, **kwargs):
    """Several pairs (different `hetero_offset`s): the ratios are averaged with the confidence weight
    |homodyne| + 1e-5 * 0.0015 before the velocity conversion (image_utils.py:170-199)."""
    ratio_sum = 0
    weight_sum = 0
    for homodyne, heterodyne in zip(homodynes, heterodynes):
        homodyne = np.asarray(homodyne)
        heterodyne = np.asarray(heterodyne)
        ratio = np.divide(heterodyne, homodyne, out=np.zeros_like(homodyne), where=np.abs(homodyne) > 0)
        weight = np.abs(homodyne) + 1e-5 * 0.0015
        ratio_sum = ratio_sum + ratio * weight
        weight_sum = weight_sum + weight
    return _velocity_from_ratio(ratio_sum / weight_sum, kwargs.get("exposure_time", 0.0015), kwargs.get("w_g", 30))


# ---- animation (main_animation.py) -------------------------------------------------------------------------------

def update_motion(ctx: runtime.Context, scene, flat, motions) -> None:
    """Next frame of an animation without re-uploading geometry: `motions` maps an index into `scene.shapes` (an
    animated shape) to its new `AnimatedTransform`. The shape's `to_world` is replaced, the keyframes are sent with
    `dtof_update_instances` (which rebuilds the top-level BVH for the new bounds); BLASes, materials and the film stay
    resident. The reference's animation loop loads a new scene file per frame (main_animation.py:61-64)."""
    from . import _abi
    import ctypes as C
    from .transform import Transform4
    moving = [i for i, sh in enumerate(scene.shapes) if sh.animated]
    first_moving = 1 if len(moving) < len(scene.shapes) else 0        # instance 0 is the static group, if any
    for shape_index, at in motions.items():
        if shape_index not in moving:
            raise ValueError(f"shape {shape_index} is not animated: only the keyframes of animated shapes can change")
        scene.shapes[shape_index].to_world = at
        inst_index = first_moving + moving.index(shape_index)
        t0, t1 = at.get_min_time(), at.get_max_time()                  # Instance::embree_geometry, instance.cpp:295-310
        eye = np.eye(4, dtype=np.float32)
        m0, m1 = Transform4(at.eval(t0), eye).m34(), Transform4(at.eval(t1), eye).m34()
        old = flat.instances[inst_index]
        inst = _abi.Instance(old.first_mesh, old.n_meshes, 1, float(np.float32(t0)), float(np.float32(t1)),
                             (C.c_float * 12)(*m0.tolist()), (C.c_float * 12)(*m1.tolist()))
        flat.instances[inst_index] = inst
        ctx.update_instances(inst_index, [inst])
    # the resident copy now matches the edited scene: keep integrator.render()'s upload cache (keyed on
    # Scene.fingerprint) from re-uploading what was just updated in place
    cached = getattr(scene, "_dtof_uploaded", None)
    if cached is not None and cached[0] is ctx and cached[1] is flat:
        scene._dtof_uploaded = (ctx, flat, scene.fingerprint())


# ---- drivers (program_runner.py) ---------------------------------------------------------------------------------

def render_image_multi_pass(scene, integrator, single_pass_spp: int, total_pass: int,
                            ctx: Optional[runtime.Context] = None, device: int = 0) -> np.ndarray:
    """Mean of `total_pass` renders with seed = 0 .. total_pass-1 and `single_pass_spp` samples each
    (program_runner.py:11-31). The scene stays resident; the mean is accumulated on the device."""
    own = ctx is None
    if own:
        ctx = runtime.Context(device)
    try:
        flat = ctx.upload(scene)
        params = integrator.params(scene.sensor.sampler, seed=0, spp=single_pass_spp)
        return ctx.render_multi_pass(flat, params, int(total_pass))
    finally:
        if own:
            ctx.close()


def _passes(total_spp: int):
    single = min(1024, int(total_spp))
    return single, max(1, int(total_spp) // single)


def run_scene_velocity(scene, total_spp: int = 1024, exposure_time: float = 0.0015, ctx=None, **kwargs) -> np.ndarray:
    """Ground-truth radial velocity image with the `velocity` integrator (program_runner.py:33-55)."""
    single, passes = _passes(total_spp)
    return render_image_multi_pass(scene, VelocityIntegrator(time=exposure_time), single, passes, ctx=ctx)


def run_scene_radiance(scene, total_spp: int = 1024, max_depth: int = 4, ctx=None, **kwargs) -> np.ndarray:
    """Radiance pass with the stock `path` integrator (program_runner.py:57-80)."""
    single, passes = _passes(total_spp)
    return render_image_multi_pass(scene, PathIntegrator(max_depth=max_depth), single, passes, ctx=ctx)


def run_scene_doppler_tof(scene, wave_function_type: str = "sinusoidal", low_frequency_component_only: bool = True,
                          hetero_frequency: float = 1.0, hetero_offset: float = 0.0, time_sampling_method: str = "antithetic",
                          antithetic_shift: Optional[float] = None, path_correlation_depth: int = 16,
                          exposure_time: float = 0.0015, w_g: float = 30, max_depth: int = 4,
                          use_stratified_sampling_for_each_interval: bool = True, total_spp: int = 1024, ctx=None,
                          **kwargs) -> np.ndarray:
    """One Doppler-ToF measurement image (program_runner.py:82-160): the property set of the tutorial, with
    `antithetic_shift` defaulting to 0.5 for `antithetic` and 0 otherwise."""
    if antithetic_shift is None:
        antithetic_shift = 0.5 if time_sampling_method == "antithetic" else 0.0
    integrator = DopplerToFPathIntegrator(
        is_doppler_integrator=True, max_depth=max_depth, w_g=w_g, time=exposure_time, hetero_frequency=hetero_frequency,
        hetero_offset=hetero_offset, antithetic_shift=antithetic_shift, time_sampling_method=time_sampling_method,
        path_correlation_depth=path_correlation_depth, low_frequency_component_only=low_frequency_component_only,
        wave_function_type=wave_function_type,
        use_stratified_sampling_for_each_interval=use_stratified_sampling_for_each_interval)
    single, passes = _passes(total_spp)
    return render_image_multi_pass(scene, integrator, single, passes, ctx=ctx)


def doppler_velocity_map(scene, total_spp: int = 1024, hetero_offsets: Sequence[float] = (0.0,), ctx=None, **kwargs):
    """The tutorial pipeline end to end: for every `hetero_offset` render the homodyne (hetero_frequency 0) and the
    heterodyne (hetero_frequency 1) measurement, convert both to ToF images and recover the radial velocity.
    Returns (velocity_map, homodynes, heterodynes)."""
    exposure_time = kwargs.get("exposure_time", 0.0015)
    w_g = kwargs.get("w_g", 30)
    homos, heteros = [], []
    for off in hetero_offsets:
        for freq, dst in ((0.0, homos), (1.0, heteros)):
            img = run_scene_doppler_tof(scene, hetero_frequency=freq, hetero_offset=off, total_spp=total_spp, ctx=ctx, **kwargs)
            dst.append(to_tof_image(img, exposure_time))
    if len(homos) == 1:
        v = calc_velocity_from_homo_hetero(homos[0], heteros[0], exposure_time=exposure_time, w_g=w_g)
    else:
        v = calc_velocity_from_homo_heteros(homos, heteros, exposure_time=exposure_time, w_g=w_g)
    return v, homos, heteros
