"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path through the C ABI against
(1) the reference's own per-lane radiance (tests/golden/lanes_*.json), (2) the CPU oracle on the same inputs,
(3) size-independent properties at the BASELINE configuration sizes."""
import os

import numpy as np
import pytest

import golden_util as gu
import mitsuba3dopplertof_b200 as dt

pytestmark = pytest.mark.gpu

LOWPASS = [n for n in gu.case_names() if "_full" not in n]
FULL = [n for n in gu.case_names() if "_full" in n]


@pytest.fixture(scope="module")
def ctx():
    from mitsuba3dopplertof_b200 import runtime
    c = runtime.Context(0)   # raises if the .so or the GPU is missing: no CPU fallback
    yield c
    c.close()


@pytest.mark.parametrize("name", LOWPASS)
def test_cuda_lanes_match_reference_and_oracle(ctx, name):
    import oracle_lib
    scene, params, ref = gu.load_case(name)
    flat = ctx.upload(scene)
    rec = ctx.trace_samples(params, ref["lanes"])
    orc = oracle_lib.OracleScene(flat, 0).trace(params, ref["lanes"])
    # a lane where CUDA and the oracle agree with each other but not with the reference is a triangle-edge / tie
    # decision on which Embree differs from any other BVH; every other lane must stay within OUTLIER_TOL
    same_as_oracle = (np.abs(rec["rgb"].astype(np.float64) - orc["rgb"]) <=
                      gu.REL_TOL * np.maximum(np.abs(orc["rgb"]), gu.ABS_FLOOR)).all(axis=1)
    frac, worst, bad = gu.compare(rec, ref, label=f"cuda:{name}", explained=same_as_oracle)
    assert frac >= gu.min_fraction(name, 0.99), f"{name}: {len(bad)} lanes differ from the reference run: {ref['lanes'][bad][:8]}"
    assert worst <= gu.OUTLIER_TOL, f"{name}: an unexplained lane is {worst:.2e} away from the reference"
    # against the oracle: identical streams and hit arithmetic; the shading arithmetic after the hit uses the SFU
    # approximations the reference's CUDA variants use (<= 2 ulp per op, csrc/dtof_device.cuh), so the comparison is
    # statistical: the same 1e-4 bound as against the reference on >= 99 % of lanes, a median at rounding level and
    # p90 <= 1e-5 (measured: p50 ~1e-7, p90 ~1e-6, p99 1e-5 .. 8e-5 for the 150 MHz cases -- one ulp of a 10 m path
    # length is 3e-6 rad of phase there)
    d = np.abs(rec["rgb"].astype(np.float64) - orc["rgb"])
    scale = np.maximum(np.abs(orc["rgb"]), gu.ABS_FLOOR)
    ok = (d <= gu.REL_TOL * scale).all(axis=1)
    assert ok.mean() >= gu.min_fraction(name, 0.99), f"{name}: CUDA vs oracle mismatch on lanes {ref['lanes'][~ok][:8]}"
    rel = (d / scale).max(axis=1)
    # lanes outside 1e-4 of the ORACLE are bounded too, unless the two took a different decision (depth / draw count)
    flipped = (rec["depth"] != orc["depth"]) | (rec["rng_draws"] != orc["rng_draws"])
    assert rel[~flipped].max() <= gu.OUTLIER_TOL, f"{name}: same decisions, but {rel[~flipped].max():.2e} away from the oracle"
    gu.REPORT[f"cuda-vs-oracle:{name}"] = {"fraction_within_1e-4": float(ok.mean()), "worst_rel_same_decisions": float(rel[~flipped].max()),
                                           "decision_flips": int(flipped.sum()), "median_rel": float(np.median(rel))}
    assert np.median(rel) <= 1e-6 and np.quantile(rel, 0.9) <= 1e-5, (np.median(rel), np.quantile(rel, 0.9))
    assert np.array_equal(rec["depth"][ok], orc["depth"][ok])
    assert np.array_equal(rec["rng_draws"][ok], orc["rng_draws"][ok])   # identical stream consumption
    np.testing.assert_array_equal(rec["time"], orc["time"])               # sampler + camera are bit-exact
    np.testing.assert_array_equal(rec["sample_pos"], orc["sample_pos"])
    np.testing.assert_array_equal(rec["ray_d"], orc["ray_d"])


@pytest.mark.parametrize("name", gu.pass_case_names())
def test_cuda_every_pass_matches_reference_and_oracle(ctx, name):
    """SURVEY 8(a) row a7 beyond pass 0: the C4 (2 x 2048 spp) and C5 (2 x 512 spp) pass structures and a 4-pass case, all
    four time-sampling modes. dtof_trace_samples_pass replays a lane's earlier passes (streams persist, src/render/
    integrator.cpp:299-308; sample index / dimension reset, src/render/sampler.cpp:52-55,94-103) -- the same code path
    (LaneSampler::next_time(pass), csrc/dtof_device.cuh) the production kernels run for pass >= 1."""
    import oracle_lib
    scene, params, ref = gu.load_case(name, multipass=True)
    flat = ctx.upload(scene)
    osc = oracle_lib.OracleScene(flat, 0)
    for k in range(int(ref["pass"].max()) + 1):
        sub = gu.select(ref, ref["pass"] == k)
        rec = ctx.trace_samples(params, sub["lanes"], k)
        orc = osc.trace(params, sub["lanes"], k)
        same = (np.abs(rec["rgb"].astype(np.float64) - orc["rgb"]) <= gu.REL_TOL * np.maximum(np.abs(orc["rgb"]), gu.ABS_FLOOR)).all(axis=1)
        frac, worst, bad = gu.compare(rec, sub, label=f"cuda:2p:{name}:pass{k}", explained=same)
        assert frac >= 0.99 and worst <= gu.OUTLIER_TOL, f"{name} pass {k}: lanes {sub['lanes'][bad][:8]}"
        assert same.mean() >= 0.99
        np.testing.assert_array_equal(rec["time"], orc["time"])            # sampler of pass k: bit-exact
        np.testing.assert_array_equal(rec["sample_pos"], orc["sample_pos"])
        np.testing.assert_array_equal(rec["ray_d"], orc["ray_d"])
        assert np.array_equal(rec["rng_draws"][same], orc["rng_draws"][same])


@pytest.mark.parametrize("scene_name,kw", [
    ("c4_domino", dict(resx=1024, resy=1024, spp=4096, wave="trapezoidal", tsm="antithetic_mirror", shift=0.0, w_g=150)),
    ("c5_slabroom", dict(resx=2048, resy=2048, spp=1024)),
    ("c5_slabroom", dict(resx=2048, resy=2048, spp=1024, tsm="stratified", shift=0.0, pcn=4)),
])
def test_cuda_two_pass_film_window_matches_oracle(ctx, scene_name, kw):
    """The PRODUCTION render (all passes, film splat) over a window of lanes of the full-size two-pass wavefront against
    the oracle's render of the same window: pass >= 1 through the very kernels the bench runs."""
    import oracle_lib
    scene = dt.load_file(os.path.join(gu.SCENES, scene_name + ".xml"), **kw)
    flat = ctx.upload(scene)
    base = scene.integrator.params(scene.sensor.sampler, seed=3)
    pi = ctx.pass_info(base)
    assert pi.n_passes == 2
    w = flat.width
    first_px = (flat.height // 2) * w + w // 2 - 24      # 48 pixels of one row in the middle of the frame
    lo, hi = first_px * pi.spp_per_pass, (first_px + 48) * pi.spp_per_pass
    p = scene.integrator.params(scene.sensor.sampler, seed=3, lane_begin=lo, lane_end=hi)
    rgbw = ctx.render(flat, p, develop=False)
    ref = oracle_lib.OracleScene(flat).render(p, 4, develop=False)   # 4 threads: each holds a full-frame accumulator
    assert abs(float(rgbw[..., 3].sum()) - 48 * pi.spp_per_pass * 2) < 1e-3 * 48 * pi.spp_per_pass * 2   # both passes landed
    scale = np.abs(ref[..., :3]).max()
    assert np.abs(rgbw[..., :3] - ref[..., :3]).max() <= 2e-4 * scale
    assert np.abs(rgbw[..., 3] - ref[..., 3]).max() <= 1e-4 * ref[..., 3].max()


@pytest.mark.parametrize("name", FULL)
def test_cuda_full_waveform_mode_matches_oracle(ctx, name):
    import oracle_lib
    scene, params, ref = gu.load_case(name)
    flat = ctx.upload(scene)
    rec = ctx.trace_samples(params, ref["lanes"])
    orc = oracle_lib.OracleScene(flat, 0).trace(params, ref["lanes"])
    d = np.abs(rec["rgb"].astype(np.float64) - orc["rgb"]).max(axis=1) / np.maximum(np.abs(orc["rgb"]).max(axis=1), 1e-2)
    assert (d <= 1e-5).mean() >= 0.98


@pytest.mark.parametrize("scene_name,kw", [
    ("c1_example", dict(resx=48, resy=40, spp=64)),
    ("c2_arealight", dict(resx=40, resy=40, spp=32)),
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
])
def test_cuda_film_matches_oracle(ctx, scene_name, kw):
    import oracle_lib
    scene = dt.load_file(os.path.join(gu.SCENES, scene_name + ".xml"), **kw)
    params = scene.integrator.params(scene.sensor.sampler, seed=5)
    flat = ctx.upload(scene)
    img, rgbw = ctx.render(flat, params, both=True)
    ref = oracle_lib.OracleScene(flat).render(params, develop=False)
    scale = np.abs(ref[..., :3]).max()
    assert np.abs(rgbw[..., 3] - ref[..., 3]).max() <= 1e-4 * ref[..., 3].max()
    err = np.abs(rgbw[..., :3] - ref[..., :3]).max(axis=2)
    # A glossy bounce turns a last-bit difference of the sampled microfacet normal (SFU arithmetic on the GPU, IEEE in
    # the oracle) into a different hit at a triangle edge or at the `cos_theta(wo) > 0` test once in ~20 000 samples
    # (measured: 3 of 51 200 lanes, all others agree to p99.9 = 2e-5): allow that many pixels to be off.
    allowed = 0.005 * err.size if scene_name in ("c12_roughconductor", "c15_roughdielectric") else 0
    assert (err > 2e-4 * scale).sum() <= allowed, f"{(err > 2e-4 * scale).sum()} pixels differ (max {err.max():.3e}, scale {scale:.3e})"
    w = np.where(rgbw[..., 3:] == 0, 1, rgbw[..., 3:])
    np.testing.assert_allclose(img, rgbw[..., :3] / w, rtol=1e-6, atol=1e-9)   # develop = RGB / W


@pytest.mark.parametrize("rfilter", ["box", "gaussian", "tent"])
def test_cuda_film_filters(ctx, rfilter):
    import oracle_lib
    scene = dt.load_file(os.path.join(gu.SCENES, "c1_example.xml"), resx=24, resy=24, spp=32)
    scene.sensor.film.rfilter = rfilter
    params = scene.integrator.params(scene.sensor.sampler)
    flat = ctx.upload(scene)
    rgbw = ctx.render(flat, params, develop=False)
    ref = oracle_lib.OracleScene(flat).render(params, develop=False)
    assert np.abs(rgbw - ref).max() <= 2e-4 * np.abs(ref).max()
    if rfilter == "box":
        assert np.array_equal(rgbw[..., 3], np.full((24, 24), 32.0, np.float32))   # exact sample counts


def test_full_size_properties_c1(ctx):
    """BASELINE config 1 at full size (256x256 @ 1024 spp): properties that need no CPU reference."""
    scene = dt.load_file(os.path.join(gu.SCENES, "c1_example.xml"))
    params = scene.integrator.params(scene.sensor.sampler)
    flat = ctx.upload(scene)
    n = 256 * 256 * 1024
    full = ctx.render(flat, params, develop=False)
    # (a) linearity over lane shards: two halves (split on a correlate-group boundary) sum to the whole
    pa = scene.integrator.params(scene.sensor.sampler, lane_begin=0, lane_end=n // 2)
    pb = scene.integrator.params(scene.sensor.sampler, lane_begin=n // 2, lane_end=n)
    halves = ctx.render(flat, pa, develop=False) + ctx.render(flat, pb, develop=False)
    assert np.abs(halves - full).max() <= 1e-4 * np.abs(full[..., :3]).max() + 1e-3   # W ~ 1e3 -> 1e-3 abs
    # (b) film weight: tent of radius 1 is a partition of unity -> interior pixels collect spp of weight
    inner = full[2:-2, 2:-2, 3]
    assert abs(inner.mean() - 1024.0) < 0.5
    # (c) the reference's committed render of this very scene (cuda_rgb, fp16 EXR): same sample streams
    gold = np.load(os.path.join(gu.GOLDEN, "scene_exr.npy")).astype(np.float32)
    img = full[..., :3] / full[..., 3:]
    num = ((img - gold) ** 2).mean()
    den = (gold ** 2).mean()
    assert num / den < 5e-3, f"relative MSE vs configs_example/scene.exr = {num / den:.3e}"


def test_c_abi_error_behaviour(ctx):
    scene = dt.load_file(os.path.join(gu.SCENES, "c1_example.xml"), resx=16, resy=16, spp=16)
    flat = ctx.upload(scene)
    p = scene.integrator.params(scene.sensor.sampler)
    p.wave_function_type = 9
    with pytest.raises(ValueError):
        ctx.render(flat, p)
    p = scene.integrator.params(scene.sensor.sampler)
    p.lane_end = 10 ** 12
    with pytest.raises(ValueError):
        ctx.render(flat, p)
    p = scene.integrator.params(scene.sensor.sampler)
    with pytest.raises(ValueError):
        ctx.trace_samples(p, [16 * 16 * 16])   # one past the wavefront


@pytest.mark.parametrize("mode,world", [("slots", 4), ("tiles", 3)])
def test_interleaved_shards_sum_to_whole(ctx, mode, world):
    """The shards the multi-GPU path hands to each rank (distributed.shard_params) partition the wavefront."""
    from mitsuba3dopplertof_b200.distributed import shard_params
    scene = dt.load_file(os.path.join(gu.SCENES, "c3_rotor.xml"), resx=40, resy=24, spp=64, pcn=4)
    params = scene.integrator.params(scene.sensor.sampler, seed=2)
    flat = ctx.upload(scene)
    pi = ctx.pass_info(params)
    whole = ctx.render(flat, params, develop=False)
    parts = sum(ctx.render(flat, shard_params(params, pi, world, r, mode, tile_pixels=7), develop=False) for r in range(world))
    assert np.abs(parts - whole).max() <= 1e-5 * np.abs(whole).max()
    assert np.abs(parts[..., 3] - whole[..., 3]).max() <= 1e-4


def test_multi_pass_driver_equals_mean_of_renders(ctx):
    """dtof_render_multi_pass == render_image_multi_pass of the tutorials: mean of developed renders, seed = 0, 1, 2."""
    scene = dt.load_file(os.path.join(gu.SCENES, "c2_arealight.xml"), resx=48, resy=32, spp=64)
    flat = ctx.upload(scene)
    base = scene.integrator.params(scene.sensor.sampler, seed=10, spp=32)
    got = ctx.render_multi_pass(flat, base, 3)
    want = np.mean([ctx.render(flat, scene.integrator.params(scene.sensor.sampler, seed=10 + i, spp=32)) for i in range(3)], axis=0)
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    with pytest.raises(ValueError):
        ctx.render_multi_pass(flat, base, 0)


@pytest.mark.parametrize("hide", [False, True])
def test_environment_emitter_and_hide_emitters(ctx, hide):
    """`constant` environment emitter (src/emitters/constant.cpp): escaping paths collect its radiance; with
    hide_emitters the camera does not see it directly (valid_ray starts false, dopplertofpath.cpp:102,279)."""
    import oracle_lib
    scene = dt.load_file(os.path.join(gu.SCENES, "c7_constant.xml"), resx=32, resy=32, spp=32, hetero_frequency=0.0)
    scene.integrator.hide_emitters = hide
    params = scene.integrator.params(scene.sensor.sampler, seed=1)
    assert params.hide_emitters == int(hide)
    flat = ctx.upload(scene)
    rgbw = ctx.render(flat, params, develop=False)
    ref = oracle_lib.OracleScene(flat).render(params, develop=False)
    assert np.abs(rgbw[..., :3] - ref[..., :3]).max() <= 2e-4 * np.abs(ref[..., :3]).max()
    sky = rgbw[0, 0, :3] / rgbw[0, 0, 3]            # top-left pixel looks past the geometry
    if hide:
        assert np.all(sky == 0)
    else:
        assert np.all(np.abs(sky) > 0)
    # two environment emitters are rejected like Scene's constructor does (scene.cpp:53-55)
    scene.emitters.append(dt.ConstantEmitter((1, 1, 1)))
    scene.scene_order = None
    with pytest.raises(ValueError):
        scene.flatten()
