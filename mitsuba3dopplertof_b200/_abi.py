"""ctypes mirror of include/dtof.h (the C ABI of libdtof_b200.so) and of the oracle's entry points.

Nothing here computes: it only lays out the plain structs the boundary takes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

ABI_VERSION = 8

# enums (include/dtof.h)
TIME_UNIFORM, TIME_STRATIFIED, TIME_ANTITHETIC, TIME_ANTITHETIC_MIRROR = range(4)
WAVE_SINUSOIDAL, WAVE_RECTANGULAR, WAVE_TRIANGULAR, WAVE_TRAPEZOIDAL = range(4)
RFILTER_BOX, RFILTER_TENT, RFILTER_GAUSSIAN, RFILTER_MITCHELL, RFILTER_CATMULLROM, RFILTER_LANCZOS = range(6)
SHAPE_MESH, SHAPE_RECTANGLE = range(2)
BSDF_DIFFUSE, BSDF_NULL_BLACK, BSDF_CONDUCTOR, BSDF_DIELECTRIC, BSDF_THINDIELECTRIC, BSDF_PLASTIC, \
    BSDF_ROUGHCONDUCTOR, BSDF_ROUGHDIELECTRIC = range(8)
EMITTER_POINT, EMITTER_AREA, EMITTER_CONSTANT, EMITTER_SPOT, EMITTER_DIRECTIONAL = range(5)
INTEGRATOR_DOPPLERTOFPATH, INTEGRATOR_VELOCITY, INTEGRATOR_PATH = range(3)

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_UNSUPPORTED, ERR_STATE = range(6)

_fp = C.POINTER(C.c_float)
_up = C.POINTER(C.c_uint32)


class Mesh(C.Structure):
    _fields_ = [
        ("n_vertices", C.c_uint32), ("n_faces", C.c_uint32),
        ("positions", _fp), ("normals", _fp), ("texcoords", _fp), ("faces", _up),
        ("bsdf", C.c_uint32), ("emitter", C.c_int32), ("flip_normals", C.c_uint32), ("kind", C.c_uint32),
        ("rect_to_world", C.c_float * 12),
    ]


class Instance(C.Structure):
    _fields_ = [
        ("first_mesh", C.c_uint32), ("n_meshes", C.c_uint32), ("animated", C.c_uint32),
        ("t0", C.c_float), ("t1", C.c_float),
        ("m0", C.c_float * 12), ("m1", C.c_float * 12),
    ]


class Bsdf(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("twosided", C.c_uint32), ("reflectance", C.c_float * 3),
                ("eta", C.c_float * 3), ("k", C.c_float * 3), ("alpha", C.c_float * 2), ("distribution", C.c_uint32),
                ("reserved", C.c_uint32)]


class Emitter(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("mesh", C.c_uint32), ("position", C.c_float * 3), ("value", C.c_float * 3),
                ("to_local", C.c_float * 9), ("cutoff_angle", C.c_float), ("beam_width", C.c_float)]


class Camera(C.Structure):
    _fields_ = [
        ("to_world", C.c_float * 12), ("sample_to_camera", C.c_float * 16),
        ("near_clip", C.c_float), ("far_clip", C.c_float),
        ("shutter_open", C.c_float), ("shutter_open_time", C.c_float),
    ]


class Film(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32),
        ("crop_offset_x", C.c_uint32), ("crop_offset_y", C.c_uint32),
        ("rfilter", C.c_uint32), ("rfilter_radius", C.c_float), ("gaussian_stddev", C.c_float),
        ("mitchell_b", C.c_float), ("mitchell_c", C.c_float),
    ]


class SceneDesc(C.Structure):
    _fields_ = [
        ("n_meshes", C.c_uint32), ("meshes", C.POINTER(Mesh)),
        ("n_instances", C.c_uint32), ("instances", C.POINTER(Instance)),
        ("n_bsdfs", C.c_uint32), ("bsdfs", C.POINTER(Bsdf)),
        ("n_emitters", C.c_uint32), ("emitters", C.POINTER(Emitter)),
        ("camera", Camera), ("film", Film),
    ]


class Params(C.Structure):
    _fields_ = [
        ("time", C.c_float), ("w_g", C.c_float), ("g_1", C.c_float), ("g_0", C.c_float),
        ("sensor_phase_offset", C.c_float), ("hetero_frequency", C.c_float),
        ("wave_function_type", C.c_uint32), ("low_frequency_component_only", C.c_uint32),
        ("max_depth", C.c_int32), ("rr_depth", C.c_int32), ("hide_emitters", C.c_uint32),
        ("time_sampling_method", C.c_uint32), ("antithetic_shift", C.c_float),
        ("use_stratified_sampling_for_each_interval", C.c_uint32), ("path_correlation_depth", C.c_uint32),
        ("sample_count", C.c_uint32), ("base_seed", C.c_uint32),
        ("time_correlate_number", C.c_uint32), ("path_correlate_number", C.c_uint32),
        ("seed", C.c_uint32),
        ("lane_begin", C.c_uint64), ("lane_end", C.c_uint64),
        ("shard_block", C.c_uint64), ("shard_count", C.c_uint32), ("shard_index", C.c_uint32),
        ("integrator", C.c_uint32), ("reserved", C.c_uint32),
    ]


class SampleRecord(C.Structure):
    _fields_ = [
        ("sample_pos", C.c_float * 2), ("time", C.c_float),
        ("ray_o", C.c_float * 3), ("ray_d", C.c_float * 3), ("ray_maxt", C.c_float),
        ("rgb", C.c_float * 3), ("path_length", C.c_float),
        ("depth", C.c_uint32), ("rng_draws", C.c_uint32),
    ]


SAMPLE_RECORD_DTYPE = np.dtype([
    ("sample_pos", np.float32, 2), ("time", np.float32),
    ("ray_o", np.float32, 3), ("ray_d", np.float32, 3), ("ray_maxt", np.float32),
    ("rgb", np.float32, 3), ("path_length", np.float32),
    ("depth", np.uint32), ("rng_draws", np.uint32),
])
assert SAMPLE_RECORD_DTYPE.itemsize == C.sizeof(SampleRecord)


class Stats(C.Structure):
    _fields_ = [
        ("samples", C.c_uint64), ("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64),
        ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64), ("inst_visits", C.c_uint64),
    ]


class Ray(C.Structure):
    _fields_ = [("o", C.c_float * 3), ("tmax", C.c_float), ("d", C.c_float * 3), ("time", C.c_float)]


class RayHit(C.Structure):
    _fields_ = [("t", C.c_float), ("u", C.c_float), ("v", C.c_float), ("prim", C.c_uint32), ("instance", C.c_int32),
                ("hit", C.c_uint32), ("nodes_visited", C.c_uint32), ("tris_tested", C.c_uint32)]


RAY_DTYPE = np.dtype([("o", "<f4", 3), ("tmax", "<f4"), ("d", "<f4", 3), ("time", "<f4")])
RAY_HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<u4"), ("instance", "<i4"), ("hit", "<u4"),
                          ("nodes_visited", "<u4"), ("tris_tested", "<u4")])


class PassInfo(C.Structure):
    _fields_ = [("spp_per_pass", C.c_uint32), ("n_passes", C.c_uint32), ("wavefront_size", C.c_uint64)]


class SceneInfo(C.Structure):
    _fields_ = [("n_triangles", C.c_uint32), ("n_nodes", C.c_uint32), ("n_instances", C.c_uint32), ("bvh_depth", C.c_uint32),
                ("traversal_bytes", C.c_uint64), ("shading_bytes", C.c_uint64), ("build_ms", C.c_float)]


# every symbol include/dtof.h declares: name -> (restype, argtypes)
_ctx = C.c_void_p
DTOF_SYMBOLS = {
    "dtof_abi_version": (C.c_uint32, []),
    "dtof_create": (C.c_int, [C.POINTER(_ctx), C.c_int]),
    "dtof_create_multi": (C.c_int, [C.POINTER(_ctx), C.POINTER(C.c_int), C.c_uint32]),
    "dtof_device_count": (C.c_uint32, [_ctx]),
    "dtof_destroy": (None, [_ctx]),
    "dtof_last_error": (C.c_char_p, [_ctx]),
    "dtof_upload_scene": (C.c_int, [_ctx, C.POINTER(SceneDesc)]),
    "dtof_scene_info_for": (C.c_int, [C.POINTER(SceneDesc), C.POINTER(SceneInfo), C.c_char_p, C.c_uint32]),
    "dtof_update_instances": (C.c_int, [_ctx, C.c_uint32, C.c_uint32, C.POINTER(Instance)]),
    "dtof_pass_info_for": (C.c_int, [_ctx, C.POINTER(Params), C.POINTER(PassInfo)]),
    "dtof_render": (C.c_int, [_ctx, C.POINTER(Params), _fp, _fp]),
    "dtof_render_accumulate": (C.c_int, [_ctx, C.POINTER(Params), C.c_int]),
    "dtof_read_film": (C.c_int, [_ctx, _fp, _fp]),
    "dtof_render_multi_pass": (C.c_int, [_ctx, C.POINTER(Params), C.c_uint32, _fp]),
    "dtof_render_device": (C.c_int, [_ctx, C.POINTER(Params), C.c_void_p, C.c_void_p]),
    "dtof_develop_device": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dtof_trace_samples": (C.c_int, [_ctx, C.POINTER(Params), C.POINTER(C.c_uint64), C.c_uint32,
                                     C.POINTER(SampleRecord)]),
    "dtof_trace_samples_pass": (C.c_int, [_ctx, C.POINTER(Params), C.POINTER(C.c_uint64), C.c_uint32, C.c_uint32,
                                          C.POINTER(SampleRecord)]),
    "dtof_trace_rays": (C.c_int, [_ctx, C.POINTER(Ray), C.c_uint32, C.c_int, C.POINTER(RayHit)]),
    "dtof_set_stats": (C.c_int, [_ctx, C.c_int]),
    "dtof_get_stats": (C.c_int, [_ctx, C.POINTER(Stats)]),
    "dtof_last_traversal_mode": (C.c_int, [_ctx]),
    "dtof_last_pipeline": (C.c_int, [_ctx]),
    "dtof_launch_count": (C.c_uint64, [_ctx]),
    "dtof_last_kernel_ms": (C.c_int, [_ctx, C.POINTER(C.c_float)]),
}


def bind(lib: C.CDLL, symbols=DTOF_SYMBOLS) -> None:
    for name, (res, args) in symbols.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args


def as_fp(a: np.ndarray):
    return a.ctypes.data_as(_fp)


def as_up(a: np.ndarray):
    return a.ctypes.data_as(_up)
