"""`path` integrator: the stock path tracer the tutorials render the radiance pass with (reference:
src/integrators/path.cpp:103-283, doppler_tutorials/src/program_runner.py:57-80).

Per-lane parity against the reference's own PathIntegrator::sample (tests/golden/lanes_path_*.json) is covered by the
generic fixture tests (test_oracle_lanes.py on the CPU, test_gpu_parity.py on the GPU: gu.case_names() includes the
path fixtures). Here: the property surface of both hosts, stream semantics, and the film."""
import os
import subprocess

import numpy as np
import pytest

import golden_util as gu
import mitsuba3dopplertof_b200 as dt
import oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _path_xml(name="c1_example.xml", sampler=None):
    xml = gu.swap_integrator(open(os.path.join(gu.SCENES, name)).read(), "path")
    if sampler == "independent":
        import re
        xml = re.sub(r'<sampler type="correlated">.*?</sampler>',
                     '<sampler type="independent">\n<integer name="sample_count" value="$spp" />\n</sampler>', xml, flags=re.S)
    return xml


def test_path_property_surface():
    scene = dt.load_string(_path_xml(), base_dir=gu.SCENES, resx=8, resy=8, spp=4)
    assert isinstance(scene.integrator, dt.PathIntegrator)
    p = scene.integrator.params(scene.sensor.sampler)
    assert p.integrator == 2 and p.max_depth == 4 and p.rr_depth == 5
    for bad in ({"time": 0.0015}, {"w_g": 30.0}, {"wave_function_type": "sinusoidal"}):
        with pytest.raises(ValueError):
            dt.PathIntegrator(**bad)              # PathIntegrator reads none of them: "unreferenced property"
    with pytest.raises(ValueError):
        dt.PathIntegrator(rr_depth=0)
    assert "PathIntegrator" in repr(dt.PathIntegrator(max_depth=3))


def test_path_and_dopplertofpath_accept_the_independent_sampler():
    """Sampler::next_1d of `correlated` IS the independent sampler's stream (src/samplers/correlated.cpp:78-90,
    src/render/sampler.cpp:115-134): the same lanes give the same values under either sampler. Under `dopplertofpath`
    the independent sampler answers with the base-class defaults (sampler.h:131-144), which is the correlated sampler
    with uniform time sampling and no path correlation -- lane for lane."""
    a = dt.load_string(_path_xml(), base_dir=gu.SCENES, resx=16, resy=16, spp=8)
    b = dt.load_string(_path_xml(sampler="independent"), base_dir=gu.SCENES, resx=16, resy=16, spp=8)
    assert b.sensor.sampler.kind == "independent"
    lanes = np.arange(0, 16 * 16 * 8, 7, dtype=np.uint64)
    ra = oracle_lib.OracleScene(a.flatten()).trace(a.integrator.params(a.sensor.sampler, seed=1), lanes)
    rb = oracle_lib.OracleScene(b.flatten()).trace(b.integrator.params(b.sensor.sampler, seed=1), lanes)
    assert np.array_equal(ra["rgb"], rb["rgb"]) and np.array_equal(ra["time"], rb["time"])
    xml = open(os.path.join(gu.SCENES, "c1_example.xml")).read()
    import re
    xml = re.sub(r'<sampler type="correlated">.*?</sampler>', '<sampler type="independent">\n<integer name="sample_count" value="$spp" />\n</sampler>',
                 xml, flags=re.S)
    c = dt.load_string(xml, base_dir=gu.SCENES, resx=16, resy=16, spp=8)                      # antithetic, pcd 4 in the file
    d = dt.load_file(os.path.join(gu.SCENES, "c1_example.xml"), resx=16, resy=16, spp=8, tsm="uniform", pcd=0, strat="false")
    rc = oracle_lib.OracleScene(c.flatten()).trace(c.integrator.params(c.sensor.sampler, seed=1), lanes)
    rd = oracle_lib.OracleScene(d.flatten()).trace(d.integrator.params(d.sensor.sampler, seed=1), lanes)
    assert np.array_equal(rc["rgb"], rd["rgb"]) and np.array_equal(rc["time"], rd["time"]) and np.abs(rc["rgb"]).max() > 0


def test_path_is_dopplertofpath_without_modulation():
    """With a constant modulation weight the two integrators must agree sample by sample when fed the same numbers:
    pcd = 0 makes dopplertofpath draw every path sample from the independent stream too; what is left differs only by
    the weight 0.5 * g_1 * cos(phase) -- homodyne (w_d = 0) and w_g -> 0 make it 0.25 for every path."""
    d = dt.load_file(os.path.join(gu.SCENES, "c1_example.xml"), resx=16, resy=16, spp=8, pcd=0, hetero_frequency=0.0, w_g=1e-12,
                     tsm="uniform", shift=0.0)
    p = dt.load_string(_path_xml(), base_dir=gu.SCENES, resx=16, resy=16, spp=8)
    lanes = np.arange(0, 16 * 16 * 8, 5, dtype=np.uint64)
    rd = oracle_lib.OracleScene(d.flatten()).trace(d.integrator.params(d.sensor.sampler, seed=3), lanes)
    rp = oracle_lib.OracleScene(p.flatten()).trace(p.integrator.params(p.sensor.sampler, seed=3), lanes)
    # same jitter (next_2d_correlate with correlate = false == next_2d), same uniform time draw
    np.testing.assert_array_equal(rd["sample_pos"], rp["sample_pos"])
    np.testing.assert_array_equal(rd["time"], rp["time"])
    # the path stream of `correlated` advanced too, but its values were never selected
    np.testing.assert_allclose(rd["rgb"], 0.25 * rp["rgb"], rtol=2e-6, atol=1e-9)
    assert np.abs(rp["rgb"]).max() > 0


def test_cpp_host_path_property_surface(tmp_path):
    cli = os.path.join(ROOT, "host", "dtof_render")
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "host")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]

    def run(xml, *extra):
        f = tmp_path / "s.xml"
        f.write_text(xml)
        return subprocess.run([cli, "--dump-desc", str(tmp_path / "d.bin"), *extra, str(f)], capture_output=True, text=True)
    assert run(_path_xml()).returncode == 0
    assert run(_path_xml(sampler="independent")).returncode == 0
    bad = _path_xml().replace('<integrator type="path">', '<integrator type="path">\n<float name="time" value="0.0015" />')
    r = run(bad)
    assert r.returncode == 1 and "unreferenced property" in r.stderr
    xml = open(os.path.join(gu.SCENES, "c1_example.xml")).read().replace('<sampler type="correlated">', '<sampler type="independent">')
    xml = xml.replace('<integer name="time_correlate_number" value="$tcn" />', '').replace('<integer name="path_correlate_number" value="$pcn" />', '')
    assert run(xml).returncode == 0   # dopplertofpath + independent: the uniform, uncorrelated stream (sampler.h:131-144)


@pytest.mark.gpu
@pytest.mark.parametrize("scene_name,kw", [
    ("c1_example", dict(resx=48, resy=40, spp=64)),
    ("c2_arealight", dict(resx=40, resy=40, spp=32)),
    ("c5_slabroom", dict(resx=32, resy=32, spp=36)),
])
def test_path_cuda_film_matches_oracle(scene_name, kw):
    from mitsuba3dopplertof_b200 import runtime
    scene = dt.load_string(_path_xml(scene_name + ".xml"), base_dir=gu.SCENES, **kw)
    params = scene.integrator.params(scene.sensor.sampler, seed=4)
    ctx = runtime.Context(0)
    flat = ctx.upload(scene)
    img, rgbw = ctx.render(flat, params, both=True)
    ref = oracle_lib.OracleScene(flat).render(params, develop=False)
    scale = np.abs(ref[..., :3]).max()
    assert np.abs(rgbw[..., 3] - ref[..., 3]).max() <= 1e-4 * ref[..., 3].max()
    assert np.abs(rgbw[..., :3] - ref[..., :3]).max() <= 2e-4 * scale
    assert (img >= 0).all() and img.max() > 0          # radiance, not a signed correlation image
