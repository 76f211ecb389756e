"""GPU parity tests of the wavefront pipeline (csrc/dtof_wavefront.cuh: generate / trace with dynamic ray fetch /
shade / shadow / splat kernels exchanging compacted queues through HBM).

It runs the same per-lane algorithm as the fused kernel, so its film must agree with the fused kernel's to float
summation order and with the CPU oracle to the film tolerance of tests/test_gpu_parity.py, for every traversal
mode, with several batches per render, with interleaved sharding, with unbounded path depth (host-side queue-length
polling) and for the stock `path` integrator."""
import os

import numpy as np
import pytest

import golden_util as gu
import mitsuba3dopplertof_b200 as dt
from mitsuba3dopplertof_b200 import procedural, runtime

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = runtime.Context(0)
    yield c
    c.close()


@pytest.fixture
def env():
    keys = ("DTOF_WAVEFRONT", "DTOF_WF_BATCH", "DTOF_WF_THRESHOLD", "DTOF_MODE")
    saved = {k: os.environ.get(k) for k in keys}
    yield os.environ
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _render(ctx, flat, params, env, wavefront, **extra):
    env["DTOF_WAVEFRONT"] = str(wavefront)
    for k, v in extra.items():
        env[k] = str(v)
    rgbw = ctx.render(flat, params, develop=False)
    assert ctx.last_pipeline() == wavefront
    for k in extra:
        env.pop(k, None)
    return rgbw


CASES = [
    ("c1_example", dict(resx=48, resy=40, spp=64)),
    ("c2_arealight", dict(resx=40, resy=40, spp=32)),
    ("c2b_two_emitters", dict(resx=32, resy=32, spp=32)),
    ("c3_rotor", dict(resx=32, resy=32, spp=32, wave="rectangular", tsm="stratified", pcn=4)),
    ("c4_domino", dict(resx=64, resy=32, spp=32, wave="trapezoidal", tsm="antithetic_mirror", shift=0.0, w_g=150)),
    ("c5_slabroom", dict(resx=32, resy=32, spp=36, tcn=3, pcn=6, tsm="antithetic")),   # spp not a multiple of 32
    ("c7_constant", dict(resx=40, resy=32, spp=32, hetero_frequency=0.0, max_depth=6)),  # constant environment emitter
    ("c8_conductor", dict(resx=40, resy=40, spp=32, hetero_frequency=0.0, max_depth=8)),  # mirrors: delta BSDF samples
    ("c9_dielectric", dict(resx=40, resy=40, spp=32, hetero_frequency=0.0, max_depth=10, rr_depth=4)),  # glass: eta tracking
    ("c10_thinglass", dict(resx=40, resy=40, spp=32, hetero_frequency=0.0, max_depth=8)),  # thin pane: Null transmission
    ("c11_plastic", dict(resx=40, resy=40, spp=32, hetero_frequency=0.0, max_depth=6)),  # plastic: smooth + delta lobe
    ("c12_roughconductor", dict(resx=40, resy=40, spp=32, hetero_frequency=0.0, max_depth=6)),  # microfacets
    ("c14_spot", dict(resx=40, resy=40, spp=32, hetero_frequency=0.0)),                         # spot light falloff
    ("c15_roughdielectric", dict(resx=40, resy=40, spp=32, hetero_frequency=0.0, max_depth=8)),  # frosted glass
    ("c16_directional", dict(resx=40, resy=40, spp=32, hetero_frequency=0.0)),                  # distant lights
]


@pytest.mark.parametrize("scene_name,kw", CASES)
@pytest.mark.parametrize("mode", [0, 1])   # BVH walked from HBM / staged in shared memory
def test_wavefront_film_matches_fused_and_oracle(ctx, env, scene_name, kw, mode):
    import oracle_lib
    scene = dt.load_file(os.path.join(gu.SCENES, scene_name + ".xml"), **kw)
    params = scene.integrator.params(scene.sensor.sampler, seed=3)
    flat = ctx.upload(scene)
    env["DTOF_MODE"] = str(mode)
    fused = _render(ctx, flat, params, env, 0)
    wave = _render(ctx, flat, params, env, 1)
    assert ctx.last_traversal_mode() == mode
    scale = np.abs(fused[..., :3]).max()
    # identical per-lane values, different order of the film atomics
    assert np.abs(wave[..., 3] - fused[..., 3]).max() <= 5e-6 * fused[..., 3].max()   # float atomics arrive in any order: ~2 ulp of the sum
    assert np.abs(wave[..., :3] - fused[..., :3]).max() <= 2e-5 * scale
    ref = oracle_lib.OracleScene(flat).render(params, develop=False)
    assert np.abs(wave[..., 3] - ref[..., 3]).max() <= 1e-4 * ref[..., 3].max()
    err = np.abs(wave[..., :3] - ref[..., :3]).max(axis=2)
    allowed = 0.005 * err.size if scene_name in ("c12_roughconductor", "c15_roughdielectric") else 0   # see test_gpu_parity
    assert (err > 2e-4 * np.abs(ref[..., :3]).max()).sum() <= allowed


@pytest.mark.parametrize("threshold", [0, 8, 32])
def test_batches_and_fetch_threshold_do_not_change_the_film(ctx, env, threshold):
    scene = dt.load_file(os.path.join(gu.SCENES, "c2_arealight.xml"), resx=40, resy=24, spp=32)
    params = scene.integrator.params(scene.sensor.sampler, seed=1)
    flat = ctx.upload(scene)
    one = _render(ctx, flat, params, env, 1)
    # 40 * 24 * 32 = 30720 lanes in batches of 4099 (ragged last batch, batch borders inside a pixel)
    many = _render(ctx, flat, params, env, 1, DTOF_WF_BATCH=4099, DTOF_WF_THRESHOLD=threshold)
    scale = np.abs(one[..., :3]).max()
    assert np.abs(many[..., :3] - one[..., :3]).max() <= 2e-5 * scale
    assert np.abs(many[..., 3] - one[..., 3]).max() <= 5e-6 * one[..., 3].max()


def test_unbounded_depth_with_russian_roulette(ctx, env):
    """max_depth = -1: the driver polls the queue length from the 8th bounce on; paths end by Russian roulette."""
    import oracle_lib
    scene = dt.load_file(os.path.join(gu.SCENES, "c1_example.xml"), resx=24, resy=24, spp=16, max_depth=-1, rr_depth=2)
    params = scene.integrator.params(scene.sensor.sampler, seed=9)
    flat = ctx.upload(scene)
    fused = _render(ctx, flat, params, env, 0)
    wave = _render(ctx, flat, params, env, 1)
    scale = np.abs(fused[..., :3]).max()
    assert np.abs(wave[..., :3] - fused[..., :3]).max() <= 2e-5 * scale
    ref = oracle_lib.OracleScene(flat).render(params, develop=False)
    assert np.abs(wave[..., :3] - ref[..., :3]).max() <= 2e-4 * np.abs(ref[..., :3]).max()


def test_sharded_wavefront_renders_sum_to_the_full_film(ctx, env):
    scene = dt.load_file(os.path.join(gu.SCENES, "c2_arealight.xml"), resx=32, resy=32, spp=32)
    sampler = scene.sensor.sampler
    flat = ctx.upload(scene)
    full = _render(ctx, flat, scene.integrator.params(sampler, seed=2), env, 1)
    parts = []
    for r in range(3):   # sample-slot sharding on correlate-group boundaries, 3 shards (ragged)
        p = scene.integrator.params(sampler, seed=2)
        p.shard_block, p.shard_count, p.shard_index = 4, 3, r
        parts.append(_render(ctx, flat, p, env, 1, DTOF_WF_BATCH=5000))
    total = np.sum(parts, axis=0)
    assert np.abs(total[..., 3] - full[..., 3]).max() <= 1e-5 * full[..., 3].max()
    assert np.abs(total[..., :3] - full[..., :3]).max() <= 2e-5 * np.abs(full[..., :3]).max()


def test_path_integrator_through_the_wavefront(ctx, env):
    scene = dt.load_file(os.path.join(gu.SCENES, "c2_arealight.xml"), resx=32, resy=32, spp=32)
    integ = dt.PathIntegrator(max_depth=5, rr_depth=3)
    params = integ.params(scene.sensor.sampler, seed=6)
    flat = ctx.upload(scene)
    fused = _render(ctx, flat, params, env, 0)
    wave = _render(ctx, flat, params, env, 1)
    assert np.abs(fused[..., :3]).max() > 0
    assert np.abs(wave[..., :3] - fused[..., :3]).max() <= 2e-5 * np.abs(fused[..., :3]).max()


def test_large_mesh_defaults_to_the_wavefront(ctx, env):
    """A BVH that does not fit shared memory is walked from HBM; that is where the wavefront pipeline is the default."""
    env.pop("DTOF_WAVEFRONT", None)
    env.pop("DTOF_MODE", None)
    sc = dt.load_file(os.path.join(gu.SCENES, "c5_slabroom.xml"), resx=64, resy=64, spp=16)
    scene = procedural.large_scene(sc, n=80, seed=1234)
    params = scene.integrator.params(scene.sensor.sampler, seed=4)
    flat = ctx.upload(scene)
    wave = ctx.render(flat, params, develop=False)
    assert ctx.last_traversal_mode() == 0 and ctx.last_pipeline() == 1
    fused = _render(ctx, flat, params, env, 0)
    assert np.abs(wave[..., :3] - fused[..., :3]).max() <= 2e-5 * np.abs(fused[..., :3]).max()
    # small scenes keep the fused kernel
    small = dt.load_file(os.path.join(gu.SCENES, "c1_example.xml"), resx=16, resy=16, spp=16)
    env.pop("DTOF_WAVEFRONT", None)
    ctx.render(ctx.upload(small), small.integrator.params(small.sensor.sampler), develop=False)
    assert ctx.last_pipeline() == 0


def test_two_pass_render_keeps_the_streams_across_passes(ctx, env):
    """2048 x 2048 @ 1024 spp is 2 passes x 512 spp (wavefront of 2^32 lanes > 0xffffffff, integrator.cpp:231-238); a lane range keeps
    the test small. The random streams continue from pass 0 into pass 1 (integrator.cpp:299-308): the wavefront pipeline
    carries them in its per-lane state, the fused kernel in registers."""
    scene = dt.load_file(os.path.join(gu.SCENES, "c2_arealight.xml"), resx=2048, resy=2048, spp=1024)
    flat = ctx.upload(scene)
    first = (1000 * 2048 + 1000) * 512           # pixel (1000, 1000), spp_per_pass = 512
    params = scene.integrator.params(scene.sensor.sampler, seed=11, lane_begin=first, lane_end=first + 3 * 512 + 100)
    pi = ctx.pass_info(params)
    assert (pi.n_passes, pi.spp_per_pass) == (2, 512)
    fused = _render(ctx, flat, params, env, 0)
    wave = _render(ctx, flat, params, env, 1, DTOF_WF_BATCH=1024)   # 2 batches x 2 passes, ragged
    assert fused[..., 3].sum() > 0
    assert np.abs(wave[..., 3] - fused[..., 3]).max() <= 5e-6 * fused[..., 3].max()   # float atomics arrive in any order: ~2 ulp of the sum
    assert np.abs(wave[..., :3] - fused[..., :3]).max() <= 2e-5 * np.abs(fused[..., :3]).max()


def test_full_size_c2_both_pipelines_agree(ctx, env):
    """BASELINE config 2 at full size (512x512 @ 1024 spp, 268 M lanes, 16 batches of 16 Mi lanes on four streams):
    the two pipelines accumulate the same per-lane values, so the films agree to float32 summation order."""
    scene = dt.load_file(os.path.join(gu.SCENES, "c2_arealight.xml"), resx=512, resy=512)
    params = scene.integrator.params(scene.sensor.sampler)
    flat = ctx.upload(scene)
    fused = _render(ctx, flat, params, env, 0)
    wave = _render(ctx, flat, params, env, 1)
    scale = np.abs(fused[..., :3]).max()
    assert np.abs(wave[..., :3] - fused[..., :3]).max() <= 1e-4 * scale
    assert np.abs(wave[..., 3] - fused[..., 3]).max() <= 1e-4 * fused[..., 3].max()
    # tent of radius 1 is a partition of unity: interior pixels collect spp of weight
    assert abs(wave[2:-2, 2:-2, 3].mean() - 1024.0) < 0.5
    # heterodyne + antithetic pairs: the image is a small difference of large per-sample values
    img = wave[..., :3] / wave[..., 3:]
    assert np.isfinite(img).all() and np.abs(img).max() < 1.0
