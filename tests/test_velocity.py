"""`velocity` integrator (reference: src/integrators/velocity.cpp:87-127).

The reference has no test or artefact for it, so parity is pinned by (0) per-lane values of the reference's own
VelocityIntegrator::sample driven by the JIT-variant sample streams (tests/golden/lanes_velocity_*.json,
tests/golden/make_golden.py), (a) an analytic property of the example scene (its two cubes translate by -/+0.015 along z in 1.5 ms while the
camera looks down -z, so the radial velocity of every visible box face is +/-10 up to the cosine of the pixel's
viewing angle; static walls give exactly 0) and (b) CUDA == oracle on identical sample streams."""
import os

import numpy as np
import pytest

import golden_util as gu
import mitsuba3dopplertof_b200 as dt
import oracle_lib

VEL_XML = None


def _velocity_scene(tmp_path_factory=None, **kw):
    """c1_example.xml with the integrator swapped for <integrator type="velocity">."""
    global VEL_XML
    if VEL_XML is None:
        import re
        xml = open(os.path.join(gu.SCENES, "c1_example.xml")).read()
        VEL_XML = re.sub(r'<integrator type="dopplertofpath">.*?</integrator>',
                         '<integrator type="velocity">\n<float name="time" value="$T" />\n<integer name="max_depth" value="$max_depth" />\n</integrator>',
                         xml, flags=re.S)
        # parameters that only the Doppler integrator referenced must not stay as unused <default>s ... they are
        # harmless: unused defaults are allowed, only unused -D parameters throw (xml.cpp:1069)
    return dt.load_string(VEL_XML, base_dir=gu.SCENES, **kw)


def test_velocity_property_surface():
    scene = _velocity_scene(resx=8, resy=8, spp=4)
    assert isinstance(scene.integrator, dt.VelocityIntegrator)
    p = scene.integrator.params(scene.sensor.sampler)
    assert p.integrator == 1 and abs(p.time - 0.0015) < 1e-9
    with pytest.raises(ValueError):
        dt.VelocityIntegrator(w_g=30.0)          # not a property of `velocity`
    with pytest.raises(ValueError):
        dt.VelocityIntegrator(max_depth=-2)


def test_velocity_oracle_analytic():
    scene = _velocity_scene(resx=96, resy=96, spp=4)
    scene.sensor.film.rfilter = "box"
    flat = scene.flatten()
    img = oracle_lib.OracleScene(flat).render(scene.integrator.params(scene.sensor.sampler), develop=True)
    assert np.array_equal(img[..., 0], img[..., 1]) and np.array_equal(img[..., 0], img[..., 2])
    v = img[..., 0]
    moving = np.abs(v) > 1.0
    assert 0.05 < moving.mean() < 0.6                       # the two boxes cover part of the view
    assert np.all(v[~moving & (np.abs(v) > 0)] < 1.0)       # box silhouettes: mixed pixels only
    inner = v[moving]
    # |v| = 10 / cos(angle to the optical axis) for faces translating along z: fov 19.5 deg -> within 2 %
    core = inner[(np.abs(np.abs(inner) - 10.0) < 0.3)]
    assert core.size > 0.7 * inner.size
    assert (core > 0).any() and (core < 0).any()            # one box approaches, the other recedes
    # static geometry: exactly zero (same hit distance at t = 0 and t = T)
    assert (v == 0).mean() > 0.3


@pytest.mark.gpu
def test_velocity_cuda_matches_oracle():
    from mitsuba3dopplertof_b200 import runtime
    scene = _velocity_scene(resx=64, resy=48, spp=8)
    params = scene.integrator.params(scene.sensor.sampler, seed=2)
    ctx = runtime.Context(0)
    flat = ctx.upload(scene)
    lanes = np.arange(0, 64 * 48 * 8, 37, dtype=np.uint64)
    rec, orc = ctx.trace_samples(params, lanes), oracle_lib.OracleScene(flat).trace(params, lanes)
    np.testing.assert_array_equal(rec["sample_pos"], orc["sample_pos"])
    np.testing.assert_array_equal(rec["time"], orc["time"])
    np.testing.assert_array_equal(rec["rng_draws"], orc["rng_draws"])
    ok = np.abs(rec["rgb"] - orc["rgb"]).max(axis=1) <= 1e-3      # (t2 - t1) / 1.5e-3 amplifies 1-ulp differences of t
    assert ok.mean() >= 0.99
    rgbw = ctx.render(flat, params, develop=False)
    ref = oracle_lib.OracleScene(flat).render(params, develop=False)
    assert np.abs(rgbw[..., 3] - ref[..., 3]).max() <= 1e-4 * ref[..., 3].max()
    assert np.abs(rgbw[..., :3] - ref[..., :3]).max() <= 2e-3 * np.abs(ref[..., :3]).max()


# ---- reference-generated fixtures -----------------------------------------------------------------------------------
# (t2 - t1) / time amplifies a 1-ulp difference of a hit distance (t ~ 5-7 -> ulp 4.8e-7) to 3e-4; Embree's
# watertight triangle test and the Moeller-Trumbore of the path agree to ~2 ulp of t, so |dv| <= 1.5e-3 (v ~ 10).
VEL_TOL = 1.5e-3


def _check_velocity(rec, ref):
    np.testing.assert_allclose(rec["sample_pos"], ref["sample_pos"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(rec["time"], ref["time"], rtol=2e-6, atol=1e-12)
    np.testing.assert_allclose(rec["ray_d"], ref["ray_d"], rtol=0, atol=2e-6)
    d = np.abs(rec["rgb"].astype(np.float64) - ref["rgb"]).max(axis=1)
    # silhouette samples (a hit at one time, a miss at the other) are edge decisions: allow 1 %
    assert (d <= VEL_TOL).mean() >= 0.99, np.sort(d)[-5:]


@pytest.mark.parametrize("name", gu.case_names("velocity"))
def test_velocity_oracle_matches_reference_lanes(name):
    scene, params, ref = gu.load_case(name)
    assert params.integrator == 1
    rec = oracle_lib.OracleScene(scene.flatten(), 0).trace(params, ref["lanes"])
    _check_velocity(rec, ref)
    assert (np.abs(ref["rgb"]) > 0.05).any()           # the fixture does see moving geometry


@pytest.mark.gpu
@pytest.mark.parametrize("name", gu.case_names("velocity"))
def test_velocity_cuda_matches_reference_lanes(name):
    from mitsuba3dopplertof_b200 import runtime
    scene, params, ref = gu.load_case(name)
    ctx = runtime.Context(0)
    ctx.upload(scene)
    _check_velocity(ctx.trace_samples(params, ref["lanes"]), ref)
