"""CPU tests: the oracle against per-lane radiance produced by the reference's own compiled integrator
(tests/golden/lanes_*.json, see tests/golden/make_golden.py), brute force and BVH traversal."""
import numpy as np
import pytest

import golden_util as gu
import oracle_lib

LOWPASS = [n for n in gu.case_names() if "_full" not in n]
FULL = [n for n in gu.case_names() if "_full" in n]


@pytest.mark.parametrize("name", LOWPASS)
def test_oracle_matches_reference_lanes(name):
    scene, params, ref = gu.load_case(name)
    flat = scene.flatten()
    rec = oracle_lib.OracleScene(flat, 0).trace(params, ref["lanes"])
    frac, worst, bad = gu.compare(rec, ref, label=f"oracle:{name}")
    assert frac >= gu.min_fraction(name, 1.0), f"{name}: lanes {ref['lanes'][bad][:8]} differ from the reference"
    assert worst <= (gu.REL_TOL if gu.min_fraction(name, 1.0) == 1.0 else 3e-4)   # worst over ALL lanes
    # the oracle's own BVH must not change a single bit
    rec_bvh = oracle_lib.OracleScene(flat, 1).trace(params, ref["lanes"])
    assert np.array_equal(rec_bvh["rgb"], rec["rgb"])


@pytest.mark.parametrize("name", FULL)
def test_oracle_full_waveform_mode(name):
    # low_frequency_component_only=false evaluates cos(w_g * t + ...) with w_g * t ~ 1.4e5 rad in float32
    # (ulp 0.016 rad; the scalar reference additionally reduces the argument in double): per-sample agreement is
    # limited to ~1e-2 by the reference's own conditioning, and square waves flip sign at edges.
    scene, params, ref = gu.load_case(name)
    rec = oracle_lib.OracleScene(scene.flatten(), 0).trace(params, ref["lanes"])
    d = np.abs(rec["rgb"].astype(np.float64) - ref["rgb"]).max(axis=1) / np.maximum(np.abs(ref["rgb"]).max(axis=1), 1e-2)
    assert np.median(d) <= 1e-2 and np.percentile(d, 95) <= 6e-2


def test_antithetic_pairs_cancel_on_static_paths():
    # heterodyne hf=1 with shift 0.5: W(t + T/2) = -W(t); pairs that only touch static geometry negate exactly
    scene, params, ref = gu.load_case("c1_hetero")
    rec = oracle_lib.OracleScene(scene.flatten(), 0).trace(params, ref["lanes"])
    rgb = rec["rgb"].reshape(-1, 2, 3)
    s = np.abs(rgb[:, 0] + rgb[:, 1]).max(axis=1) / np.abs(rgb[:, 0]).max(axis=1)
    assert np.median(s) < 2e-2 and (s < 1e-5).any()


def test_pass_split_follows_reference():
    import ctypes as C
    from mitsuba3dopplertof_b200 import _abi
    scene, params, _ = gu.load_case("c4_domino")          # 1024^2 x 4096 -> 2 passes x 2048
    flat = scene.flatten()
    pi = _abi.PassInfo()
    assert oracle_lib.lib().dtof_oracle_pass_info(C.byref(flat.desc), C.byref(params), C.byref(pi)) == 0
    assert (pi.spp_per_pass, pi.n_passes, pi.wavefront_size) == (2048, 2, 2 ** 31)
    # 2048^2 x 16384: divisor 17 -> spp_per_pass 963 does not divide 16384 -> the reference throws (sampler.cpp:81-82)
    flat.desc.film.width = flat.desc.film.height = 2048
    params.sample_count = 16384
    assert oracle_lib.lib().dtof_oracle_pass_info(C.byref(flat.desc), C.byref(params), C.byref(pi)) != 0


@pytest.mark.parametrize("name", gu.pass_case_names())
def test_oracle_matches_reference_in_every_pass(name):
    """Multi-pass renders (C4: 2 x 2048 spp, C5: 2 x 512 spp, a 4-pass case; all four time-sampling modes): the streams
    continue from pass to pass, sample_index = pass * spp_per_pass + idx % spp_per_pass, the dimension index restarts
    (src/render/integrator.cpp:299-308, src/render/sampler.cpp:52-55,94-103, src/samplers/correlated.cpp:92-153)."""
    scene, params, ref = gu.load_case(name, multipass=True)
    osc = oracle_lib.OracleScene(scene.flatten(), 0)
    assert ref["pass"].max() >= 1
    for k in range(int(ref["pass"].max()) + 1):
        sub = gu.select(ref, ref["pass"] == k)
        rec = osc.trace(params, sub["lanes"], k)
        frac, worst, bad = gu.compare(rec, sub, label=f"oracle:2p:{name}:pass{k}")
        # the 150 MHz case (c4): one ulp of a 10 m path length is already 3e-6 rad of phase, a lane in 200 lands at 1.3e-4
        assert frac >= 0.99 and worst <= 3e-4, f"{name} pass {k}: lanes {sub['lanes'][bad][:8]} differ"
    # the time sample of pass k is not the time sample of pass 0 (the streams really moved on)
    t0 = osc.trace(params, ref["lanes"][ref["pass"] == 0], 0)["time"]
    t1 = osc.trace(params, ref["lanes"][ref["pass"] == 0], 1)["time"]
    assert (t0 != t1).mean() > 0.9
