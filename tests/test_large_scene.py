"""Large-mesh path (BASELINE config 5 stand-in): procedural displaced cube-sphere as an animated instance.
CPU: the host half of the upload (flattening + parallel subtree BVH build) through dtof_scene_info_for.
GPU: the BVH-from-HBM traversal mode against the oracle (its own, independent BVH) on identical sample streams."""
import os

import numpy as np
import pytest

import golden_util as gu
import mitsuba3dopplertof_b200 as dt
from mitsuba3dopplertof_b200 import procedural, runtime


def _scene(n, **kw):
    sc = dt.load_file(os.path.join(gu.SCENES, "c5_slabroom.xml"), **kw)
    return procedural.large_scene(sc, n=n, seed=1234)


def test_displaced_sphere_is_deterministic_and_closed():
    p1, f1 = procedural.displaced_cube_sphere(12, seed=7)
    p2, f2 = procedural.displaced_cube_sphere(12, seed=7)
    assert np.array_equal(p1, p2) and np.array_equal(f1, f2)
    assert f1.shape == (12 * 12 * 12, 3) and p1.shape == (6 * 13 * 13, 3)
    r = np.linalg.norm(p1, axis=1)
    assert 0.8 < r.min() and r.max() < 1.2
    # consistent outward orientation: signed volume is positive and close to the unit sphere's
    v = np.einsum("ij,ij->i", p1[f1[:, 0]], np.cross(p1[f1[:, 1]], p1[f1[:, 2]])).sum() / 6.0
    assert abs(v - 4.0 / 3.0 * np.pi) < 0.4 and v > 0


@pytest.mark.parametrize("n", [16, 80])   # 80 -> 76 800 triangles: the parallel subtree builder (>= 65 536)
def test_host_upload_half_builds_the_bvh(n):
    scene = _scene(n, resx=32, resy=32, spp=4)
    flat = scene.flatten()
    info = runtime.scene_info(flat)
    assert info.n_triangles == 12 * n * n + 92 - 0 or info.n_triangles == flat.n_triangles
    assert info.n_triangles == flat.n_triangles
    assert info.n_instances == 4
    # binary tree with leaves of <= 4 (large) / <= 2 (small) triangles: inner nodes < triangles
    assert info.n_triangles / 8 < info.n_nodes < info.n_triangles
    assert info.bvh_depth < 64
    assert info.traversal_bytes == info.n_nodes * 64 + info.n_triangles * 48 + info.n_instances * 128


def test_scene_info_reports_reference_errors():
    scene = _scene(4, resx=8, resy=8, spp=4)
    scene.shapes[-1].radiance = (1.0, 1.0, 1.0)          # an emitter inside an animated instance
    with pytest.raises(ValueError):                      # python host: shapegroup.cpp:27-30
        scene.flatten()


@pytest.mark.gpu
def test_large_mesh_lanes_match_oracle():
    import oracle_lib
    scene = _scene(80, resx=96, resy=96, spp=16)
    params = scene.integrator.params(scene.sensor.sampler, seed=4)
    ctx = runtime.Context(0)
    flat = ctx.upload(scene)
    lanes = np.arange(5, 96 * 96 * 16, 61, dtype=np.uint64)
    rec = ctx.trace_samples(params, lanes)
    assert ctx.last_traversal_mode() == 0                 # BVH read from HBM through L1/L2
    orc = oracle_lib.OracleScene(flat).trace(params, lanes)
    np.testing.assert_array_equal(rec["time"], orc["time"])
    same = (rec["depth"] == orc["depth"]) & (rec["rng_draws"] == orc["rng_draws"])
    assert same.mean() >= 0.995
    d = np.abs(rec["rgb"].astype(np.float64) - orc["rgb"]).max(axis=1) / np.maximum(np.abs(orc["rgb"]).max(axis=1), gu.ABS_FLOOR)
    assert (d <= gu.REL_TOL).mean() >= 0.99 and np.median(d) <= 1e-6
    # film level
    rgbw = ctx.render(flat, params, develop=False)
    ref = oracle_lib.OracleScene(flat).render(params, develop=False)
    assert np.abs(rgbw[..., :3] - ref[..., :3]).max() <= 5e-4 * np.abs(ref[..., :3]).max()
    assert np.abs(rgbw[..., 3] - ref[..., 3]).max() <= 1e-4 * ref[..., 3].max()


@pytest.mark.gpu
def test_c5_full_size_mesh_both_pipelines_match_oracle():
    """BASELINE config 5 at its REAL size: the 4.2 M-triangle mesh (n = 592: parallel subtree builder, BVH depth ~30, a
    working set of ~330 MB > the 126 MB L2), 2048 x 2048 @ 1024 spp = 2 passes x 512. (1) 2 600 lanes spread over the
    frame, both passes, through the record kernel (BVH walked from HBM) against the oracle and its own, independent BVH;
    (2) a window of the production render through BOTH pipelines (wavefront = the default here, fused) against the
    oracle's film of the same lanes."""
    import oracle_lib
    scene = _scene(592, resx=2048, resy=2048, spp=1024)
    params = scene.integrator.params(scene.sensor.sampler, seed=4)
    ctx = runtime.Context(0)
    flat = ctx.upload(scene)
    assert flat.n_triangles > 4_000_000
    pi = ctx.pass_info(params)
    assert (pi.n_passes, pi.spp_per_pass) == (2, 512)
    osc = oracle_lib.OracleScene(flat)
    n_lanes = 2048 * 2048 * 512
    lanes = np.arange(7, n_lanes, n_lanes // 2600, dtype=np.uint64)
    assert lanes.size >= 2000
    for k in (0, 1):
        rec = ctx.trace_samples(params, lanes, k)
        assert ctx.last_traversal_mode() == 0
        orc = osc.trace(params, lanes, k)
        np.testing.assert_array_equal(rec["time"], orc["time"])
        np.testing.assert_array_equal(rec["ray_d"], orc["ray_d"])
        same = (rec["depth"] == orc["depth"]) & (rec["rng_draws"] == orc["rng_draws"])
        d = np.abs(rec["rgb"].astype(np.float64) - orc["rgb"]).max(axis=1) / np.maximum(np.abs(orc["rgb"]).max(axis=1), gu.ABS_FLOOR)
        gu.REPORT[f"cuda-vs-oracle:c5_full_mesh:pass{k}"] = {
            "lanes": int(lanes.size), "fraction_within_1e-4": float((d <= gu.REL_TOL).mean()), "same_decisions": float(same.mean()),
            "worst_rel_same_decisions": float(d[same].max()), "median_rel": float(np.median(d))}
        assert same.mean() >= 0.995 and (d <= gu.REL_TOL).mean() >= 0.99 and np.median(d) <= 1e-6
        assert d[same].max() <= gu.OUTLIER_TOL
    # (2) 24 pixels in the middle of the frame (they look at the mesh), all 2 x 512 samples each
    first = (1024 * 2048 + 1012) * 512
    p = scene.integrator.params(scene.sensor.sampler, seed=4, lane_begin=first, lane_end=first + 24 * 512)
    ref = osc.render(p, 4, develop=False)   # 4 threads: every oracle thread holds a full-frame double accumulator
    scale = np.abs(ref[..., :3]).max()
    old = os.environ.get("DTOF_WAVEFRONT")
    try:
        for wf in ("1", "0"):
            os.environ["DTOF_WAVEFRONT"] = wf
            rgbw = ctx.render(flat, p, develop=False)
            assert ctx.last_pipeline() == int(wf) and ctx.last_traversal_mode() == 0
            assert abs(float(rgbw[..., 3].sum()) - 24 * 1024) < 1.0
            assert np.abs(rgbw[..., :3] - ref[..., :3]).max() <= 5e-4 * scale, f"pipeline {wf}"
            assert np.abs(rgbw[..., 3] - ref[..., 3]).max() <= 1e-4 * ref[..., 3].max()
    finally:
        if old is None:
            os.environ.pop("DTOF_WAVEFRONT", None)
        else:
            os.environ["DTOF_WAVEFRONT"] = old
    ctx.close()
