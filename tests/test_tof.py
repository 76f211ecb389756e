"""Doppler-ToF post-processing (mitsuba3dopplertof_b200/tof.py) and the tutorial pipeline on top of the renderer.

CPU: the host functions against vectors produced by the reference's own `doppler_tutorials/src/utils/image_utils.py`
(tests/golden/make_tof_golden.py -> tests/golden/tof_postprocess.npz).
GPU: the pipeline the tutorials run (program_runner.py): homodyne + heterodyne Doppler-ToF measurements -> radial
velocity map, compared with the ground truth of the `velocity` integrator on the example scene."""
import os

import numpy as np
import pytest

import golden_util as gu
import mitsuba3dopplertof_b200 as dt
from mitsuba3dopplertof_b200 import tof

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tof_postprocess.npz"))


def test_luminance_and_tof_scaling_match_the_reference():
    img = G["img"]
    np.testing.assert_array_equal(tof.rgb2luminance(img), G["luminance"])
    np.testing.assert_array_equal(tof.to_tof_image(img), G["tof"])
    np.testing.assert_array_equal(tof.to_tof_image(img, 0.002), G["tof_T2"])
    np.testing.assert_array_equal(tof.to_tof_image_0_5(img), G["tof_0_5"])


def test_velocity_from_one_pair_matches_the_reference():
    ho, he = G["homos"], G["heteros"]
    np.testing.assert_allclose(tof.calc_velocity_from_homo_hetero(ho[0], he[0]), G["v_single"], rtol=1e-13, atol=0)
    np.testing.assert_allclose(tof.calc_velocity_from_homo_hetero(ho[1], he[1], exposure_time=0.002, w_g=150),
                               G["v_single_kw"], rtol=1e-13, atol=0)
    # pixels without homodyne signal have ratio 0 -> velocity 0; ratios are clipped to [-1, 0.999]
    v = tof.calc_velocity_from_homo_hetero(ho[0], he[0])
    assert np.all(v[ho[0] == 0] == 0)
    vmax = 0.5 * (0.999 / 0.0015 / 0.001) * 3e8 / 30e6
    assert np.abs(v).max() <= vmax * (1 + 1e-12)


def test_velocity_from_several_pairs_matches_the_reference():
    ho, he = G["homos"], G["heteros"]
    np.testing.assert_allclose(tof.calc_velocity_from_homo_heteros(list(ho), list(he)), G["v_multi"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(tof.calc_velocity_from_homo_heteros(list(ho[:2]), list(he[:2]), exposure_time=0.001, w_g=60),
                               G["v_multi_kw"], rtol=1e-12, atol=0)


def test_inputs_are_not_modified():
    ho, he = G["homos"][0].copy(), G["heteros"][0].copy()
    tof.calc_velocity_from_homo_hetero(ho, he)
    np.testing.assert_array_equal(ho, G["homos"][0])
    np.testing.assert_array_equal(he, G["heteros"][0])


@pytest.mark.gpu
def test_doppler_pipeline_recovers_the_ground_truth_velocity():
    """Example scene (two cubes moving along z at -/+10 m/s): the velocity recovered from the homodyne / heterodyne
    measurements agrees with the `velocity` integrator where the homodyne signal is strong."""
    from mitsuba3dopplertof_b200 import runtime
    scene = dt.load_file(os.path.join(gu.SCENES, "c1_example.xml"), resx=96, resy=96, spp=1024)
    ctx = runtime.Context(0)
    try:
        gt = tof.run_scene_velocity(scene, total_spp=64, ctx=ctx)[:, :, 0]
        v, homos, heteros = tof.doppler_velocity_map(scene, total_spp=4096, hetero_offsets=(0.0, 0.25), ctx=ctx,
                                                     max_depth=2, path_correlation_depth=2)
    finally:
        ctx.close()
    assert v.shape == gt.shape == (96, 96)
    moving = np.abs(gt) > 5.0                       # interior of the two cubes (|v| = 10 m/s along the view axis)
    assert moving.sum() > 200
    strong = moving & (np.maximum(np.abs(homos[0]), np.abs(homos[1])) > 0.5 * np.median(np.abs(homos[0][moving])))
    err = np.abs(v[strong] - gt[strong])
    print("pixels", strong.sum(), "median |err|", np.median(err), "p90", np.quantile(err, 0.9),
          "sign agreement", np.mean(np.sign(v[strong]) == np.sign(gt[strong])))
    assert np.mean(np.sign(v[strong]) == np.sign(gt[strong])) > 0.9
    assert np.median(err) < 2.5                     # m/s, of 10
    static = (np.abs(gt) < 1e-3) & (np.abs(homos[0]) > np.median(np.abs(homos[0])))
    assert np.median(np.abs(v[static])) < 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("big_scene", [False, True])
def test_animation_frames_reuse_the_uploaded_scene(big_scene):
    """update_motion (dtof_update_instances + TLAS rebuild) against a fresh upload of the same frame: the moved cube
    leaves the bounds it was uploaded with, so a stale TLAS would cull it."""
    import copy
    from mitsuba3dopplertof_b200 import procedural, runtime
    from mitsuba3dopplertof_b200.transform import AnimatedTransform, Transform4

    def load():
        sc = dt.load_file(os.path.join(gu.SCENES, "c5_slabroom.xml" if big_scene else "c1_example.xml"), resx=48, resy=48, spp=32)
        return procedural.large_scene(sc, n=80, seed=1234) if big_scene else sc   # 76 800 triangles: BVH walked from HBM

    scene = load()
    moving = [i for i, sh in enumerate(scene.shapes) if sh.animated]
    assert moving
    target = moving[-1]
    old = scene.shapes[target].to_world
    shift = np.eye(4, dtype=np.float32)
    shift[0, 3], shift[1, 3] = 0.35, 0.2                         # well outside the uploaded bounds
    at = AnimatedTransform()
    for t, tr in zip(old.times, old.transforms):
        m = (shift @ tr.matrix).astype(np.float32)
        at.append(t, Transform4.from_matrix(m, np.float32))

    ctx = runtime.Context(0)
    try:
        flat = ctx.upload(scene)
        params = scene.integrator.params(scene.sensor.sampler, seed=2)
        frame0 = ctx.render(flat, params, develop=False)
        tof.update_motion(ctx, scene, flat, {target: at})
        frame1 = ctx.render(flat, params, develop=False)
        fresh_scene = load()
        fresh_scene.shapes[target].to_world = copy.deepcopy(at)
        fresh = ctx.render(ctx.upload(fresh_scene), params, develop=False)
    finally:
        ctx.close()
    scale = np.abs(fresh[..., :3]).max()
    assert np.abs(frame1[..., :3] - frame0[..., :3]).max() > 1e-2 * scale      # the frame did change
    assert np.abs(frame1[..., :3] - fresh[..., :3]).max() <= 2e-5 * scale      # and equals the fresh upload
    assert np.abs(frame1[..., 3] - fresh[..., 3]).max() <= 5e-6 * fresh[..., 3].max()
