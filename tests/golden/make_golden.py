#!/usr/bin/env python
"""Fixture generator (runs only in the build container; NOT a test).

Produces tests/golden/lanes_<case>.json by running the REFERENCE's own compiled code
(oracle/_ref: libmitsuba.so + plugins built from /root/reference, scalar_rgb + Embree) through
oracle/ref_harness/replay_harness.cpp, which feeds `DopplerToFPathIntegrator::sample()` the sample
streams of the reference's JIT variants lane by lane.

Time scaling. The scalar Embree glue stores ray.time in the ray's *tnear* slot
(src/render/scene_embree.inl:226,369) and the user-geometry callback (rectangles) returns a hit
distance shortened by tnear (src/render/shape.cpp:141-147). The JIT variants pass mint = 0
(scene_embree.inl:287,405), so this is a scalar-only artefact. To get JIT-faithful numbers out of
the scalar build the reference is run with ALL times (integrator `time`, shutter, keyframes)
multiplied by 2^-20: every quantity the path computes is invariant under a power-of-two time scale
(w_d*t, the keyframe fraction, ...) bit for bit, while the artefact shrinks below half an ulp of t.
Full-waveform mode (low_frequency_component_only=false) is not scale invariant (w_g*t) and is
therefore generated unscaled on `c5_slabroom.xml`, whose walls are triangle meshes (no user geometry).

Also writes header_vectors.json (reference header templates) and scene_exr.npy (the reference's
committed configs_example/scene.exr) when the inputs are available.
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_RT = os.path.join(ROOT, "oracle", "_ref")
SCENES = os.path.join(ROOT, "tests", "scenes")
T = np.float32(0.0015)
SCALE = np.float32(2.0 ** -20)

# name -> (scene, xml params, harness seed, time-scaled?)   [tcn/pcn/spp/res live in the xml params]
CASES = {
    "c1_hetero": ("c1_example", {}, 0, True),
    "c1_homodyne": ("c1_example", {"hetero_frequency": 0.0}, 0, True),
    "c1_rect_strat_pcn4": ("c1_example", {"wave": "rectangular", "tsm": "stratified", "shift": 0.0, "pcn": 4}, 0, True),
    "c1_trap_mirror": ("c1_example", {"wave": "trapezoidal", "tsm": "antithetic_mirror", "shift": 0.0, "w_g": 150}, 0, True),
    "c1_tri_uniform": ("c1_example", {"wave": "triangular", "tsm": "uniform", "shift": 0.0, "strat": "false"}, 3, True),
    "c1_rr": ("c1_example", {"max_depth": 16, "pcd": 16, "rr_depth": 2}, 0, True),
    "c1_seed7_spp64_tcn4": ("c1_example", {"spp": 64, "tcn": 4, "pcn": 4, "resx": 128, "resy": 128}, 7, True),
    "c1_strat_nointerval": ("c1_example", {"tsm": "stratified", "shift": 0.0, "strat": "false", "wave": "triangular"}, 1, True),
    "c1_pcd0_offset": ("c1_example", {"pcd": 0, "hetero_offset": 0.25, "tsm": "antithetic", "shift": 0.25}, 0, True),
    "c2_arealight": ("c2_arealight", {"resx": 512, "resy": 512}, 0, True),
    "c2b_two_emitters": ("c2b_two_emitters", {"max_depth": 6, "pcd": 6}, 0, True),
    "c3_rotor": ("c3_rotor", {"wave": "rectangular", "tsm": "stratified", "shift": 0.0, "pcn": 4, "resx": 512, "resy": 512}, 0, True),
    "c4_domino": ("c4_domino", {"wave": "trapezoidal", "tsm": "antithetic_mirror", "shift": 0.0, "w_g": 150,
                                "resx": 1024, "resy": 1024, "spp": 4096}, 0, True),
    "c5_slabroom_full": ("c5_slabroom", {"lowpass": "false", "wave": "rectangular"}, 0, False),
    "c5_slabroom_full_sin": ("c5_slabroom", {"lowpass": "false", "wave": "sinusoidal", "hetero_frequency": 0.0}, 2, False),
    "c5_slabroom_lowpass": ("c5_slabroom", {}, 0, True),
    # the gem as sub-mesh 1 of a double-precision v4 `.serialized` file: the reference's SerializedMesh loader
    "c6_serialized": ("c6_serialized", {}, 4, True),
    # open scene (floor + the two moving boxes) under a constant environment emitter + the point light (2 emitters)
    "c7_constant": ("c7_constant", {"max_depth": 5, "pcd": 5}, 0, True),
    # smooth conductors: the tall box (explicit eta / k, one-sided) and the back wall (two-sided perfect mirror)
    "c8_conductor": ("c8_conductor", {"max_depth": 6, "pcd": 6}, 0, True),
    "c8_conductor_homodyne": ("c8_conductor", {"hetero_frequency": 0.0, "max_depth": 8, "rr_depth": 3, "tsm": "stratified", "shift": 0.0}, 2, True),
    # smooth dielectrics: the short box is bk7 glass in air, the tall box water with tinted lobes (eta tracking)
    "c9_dielectric": ("c9_dielectric", {"max_depth": 8, "pcd": 8}, 0, True),
    "c9_dielectric_homodyne": ("c9_dielectric", {"hetero_frequency": 0.0, "max_depth": 12, "rr_depth": 3, "pcd": 2}, 5, True),
    # c9 + a tinted thin glass pane (thindielectric: Null transmission, valid_ray semantics)
    "c10_thinglass": ("c10_thinglass", {"max_depth": 8, "pcd": 8, "hetero_frequency": 0.0}, 1, True),
    # smooth plastic: the floor (nonlinear, tinted coat, one-sided) and the short box (two-sided, int_ior 1.9)
    "c11_plastic": ("c11_plastic", {"max_depth": 6, "pcd": 6}, 0, True),
    "c11_plastic_homodyne": ("c11_plastic", {"hetero_frequency": 0.0, "max_depth": 8, "rr_depth": 3}, 3, True),
    # rough conductors: GGX (two-sided, copper-like), anisotropic Beckmann (default distribution), Beckmann back wall
    "c12_roughconductor": ("c12_roughconductor", {"max_depth": 6, "pcd": 6}, 0, True),
    "c12_roughconductor_homodyne": ("c12_roughconductor", {"hetero_frequency": 0.0, "max_depth": 8, "rr_depth": 3}, 7, True),
    # named conductor materials (measured spectra -> RGB by the reference): gold mirror box, rough aluminium box
    "c13_named_metals": ("c13_named_metals", {"max_depth": 6, "pcd": 6, "hetero_frequency": 0.0}, 2, True),
    # the ToF illuminator as a spot light next to the camera (falloff between beam_width and cutoff_angle)
    "c14_spot": ("c14_spot", {"max_depth": 4}, 0, True),
    "c14_spot_homodyne": ("c14_spot", {"hetero_frequency": 0.0, "max_depth": 5, "pcd": 5}, 8, True),
    # rough dielectrics (frosted glass): GGX bk7 box, anisotropic Beckmann water box with tinted lobes
    "c15_roughdielectric": ("c15_roughdielectric", {"max_depth": 8, "pcd": 8}, 0, True),
    "c15_roughdielectric_homodyne": ("c15_roughdielectric", {"hetero_frequency": 0.0, "max_depth": 10, "rr_depth": 4}, 9, True),
    # two directional (distant) lights: `direction` property and a lookat to_world; shadows through the open front
    "c16_directional": ("c16_directional", {"max_depth": 4}, 0, True),
    "c16_directional_homodyne": ("c16_directional", {"hetero_frequency": 0.0, "max_depth": 6, "pcd": 6, "rr_depth": 3}, 4, True),
    "c7_constant_homodyne": ("c7_constant", {"hetero_frequency": 0.0, "tsm": "uniform", "shift": 0.0, "rr_depth": 2, "max_depth": 8}, 6, True),
}
# the stock path tracer (src/integrators/path.cpp) on the same scenes: the integrator element is swapped (golden_util.swap_integrator)
PATH_CASES = {
    "path_c1": ("c1_example", {}, 0, True),
    "path_c2_arealight": ("c2_arealight", {"resx": 512, "resy": 512}, 0, True),
    "path_c2b_rr": ("c2b_two_emitters", {"max_depth": 16, "rr_depth": 2}, 5, True),
    "path_c3_rotor": ("c3_rotor", {"resx": 512, "resy": 512, "spp": 64}, 2, True),
    "path_c5_slabroom": ("c5_slabroom", {}, 1, True),
    "path_c7_constant": ("c7_constant", {}, 3, True),
    "path_c8_conductor": ("c8_conductor", {"max_depth": 6}, 1, True),
    "path_c9_dielectric": ("c9_dielectric", {"max_depth": 8}, 2, True),
    "path_c10_thinglass": ("c10_thinglass", {"max_depth": 6}, 0, True),
    "path_c11_plastic": ("c11_plastic", {"max_depth": 6}, 4, True),
    "path_c12_roughconductor": ("c12_roughconductor", {"max_depth": 6}, 5, True),
    "path_c14_spot": ("c14_spot", {}, 6, True),
    "path_c16_directional": ("c16_directional", {}, 8, True),
    "path_c15_roughdielectric": ("c15_roughdielectric", {"max_depth": 8}, 7, True),
}
# the ground-truth radial velocity integrator (src/integrators/velocity.cpp). Its value (t2 - t1) / time scales by exactly
# 2^20 under the time scaling; golden_util.load_case undoes it.
VELOCITY_CASES = {
    "velocity_c1": ("c1_example", {}, 0, True),
    "velocity_c3_rotor": ("c3_rotor", {"resx": 512, "resy": 512, "spp": 64}, 1, True),
    "velocity_c4_domino": ("c4_domino", {"resx": 512, "resy": 512, "spp": 64}, 2, True),
}
# Multi-pass renders (W * H * spp > 2^32 - 1 lanes, src/render/integrator.cpp:231-238): EVERY pass of the chosen lanes is
# kept, written to lanes2p_<name>.json. The sampler streams continue from pass to pass (integrator.cpp:299-308), the sample
# index becomes pass * spp_per_pass + idx % spp_per_pass and the dimension index restarts (src/render/sampler.cpp:52-55,94-103).
PASS_CASES = {
    "c4_domino": ("c4_domino", {"wave": "trapezoidal", "tsm": "antithetic_mirror", "shift": 0.0, "w_g": 150,
                                "resx": 1024, "resy": 1024, "spp": 4096}, 1, True),                       # C4: 2 x 2048
    "c5_antithetic": ("c5_slabroom", {"resx": 2048, "resy": 2048, "spp": 1024}, 0, True),               # C5: 2 x 512
    "c5_uniform": ("c5_slabroom", {"resx": 2048, "resy": 2048, "spp": 1024, "tsm": "uniform", "shift": 0.0, "strat": "false"}, 2, True),
    "c5_stratified": ("c5_slabroom", {"resx": 2048, "resy": 2048, "spp": 1024, "tsm": "stratified", "shift": 0.0, "pcn": 4,
                                      "wave": "rectangular"}, 3, True),
    "c5_mirror_4pass": ("c5_slabroom", {"resx": 2048, "resy": 2048, "spp": 3072, "tsm": "antithetic_mirror", "shift": 0.0,
                                        "wave": "triangular"}, 4, True),                                   # 4 x 768
}
CASES.update(PATH_CASES)
CASES.update(VELOCITY_CASES)
SWAPPED = dict({k: "path" for k in PATH_CASES}, **{k: "velocity" for k in VELOCITY_CASES})
DEFAULTS = {"spp": 1024, "resx": 256, "resy": 256, "max_depth": 4, "tcn": 2, "pcn": 2}
N_PIXELS = 48


def lanes_for(case, params, rng):
    p = dict(DEFAULTS, **params)
    spp, w, h = int(p["spp"]), int(p["resx"]), int(p["resy"])
    spp_pp = spp
    wave = w * h * spp_pp
    if wave > 0xFFFFFFFF:   # integrator.cpp:231-238
        spp_pp //= (wave + 0xFFFFFFFF - 1) // 0xFFFFFFFF
    group = 4
    px = rng.integers(w // 8, w - w // 8, N_PIXELS)
    py = rng.integers(h // 8, h - h // 8, N_PIXELS)
    slot = rng.integers(0, spp_pp // group, N_PIXELS) * group
    lanes = []
    for x, y, s in zip(px, py, slot):
        base = (int(y) * w + int(x)) * spp_pp + int(s)
        lanes += [base + k for k in range(group)]
    return lanes


def run_case(name, table=None):
    all_passes = table is not None
    scene, params, seed, scaled = (table or CASES)[name]
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    lanes = lanes_for(name, params, rng)
    lanes_file = f"/tmp/_dtof_lanes_{name}.txt"
    with open(lanes_file, "w") as f:
        f.write("\n".join(map(str, lanes)) + "\n")
    p = dict(DEFAULTS, **params)
    tval = T * SCALE if scaled else T
    scene_file = os.path.join(SCENES, scene + ".xml")
    if name in SWAPPED:
        sys.path[:0] = [os.path.dirname(HERE), ROOT]
        import golden_util
        scene_file = os.path.join(SCENES, f"_tmp_{name}.xml")   # next to the .ply files it references
        with open(scene_file, "w") as f:
            f.write(golden_util.swap_integrator(open(os.path.join(SCENES, scene + ".xml")).read(), SWAPPED[name]))
    cmd = [os.path.join(REF_RT, "replay_harness"), scene_file, "--seed", str(seed),
           "--tcn", str(p["tcn"]), "--pcn", str(p["pcn"]), "--lanes", lanes_file, f"-DT={tval:.9g}"]
    if name in VELOCITY_CASES:   # the integrator's own `time`: 0.75 T (golden_util.swap_integrator), scaled like T
        cmd.append(f"-DTvel={np.float32(0.001125) * (SCALE if scaled else np.float32(1)):.9g}")
    for k, v in params.items():
        cmd.append(f"-D{k}={v}")
    env = dict(os.environ, LD_LIBRARY_PATH=REF_RT, DTOF_REF_DIR=REF_RT)
    out = subprocess.run(cmd, check=True, capture_output=True, text=True, env=env).stdout.splitlines()
    if name in SWAPPED:
        os.remove(scene_file)
    rows = [l.split() for l in out if l and l[0].isdigit()]
    if all_passes:
        assert len(rows) % len(lanes) == 0 and len(rows) > len(lanes), (len(rows), len(lanes))
    else:
        rows = [r for r in rows if int(r[1]) == 0]    # single-pass fixtures: pass 0 (the multi-pass ones are PASS_CASES)
        assert len(rows) == len(lanes), (len(rows), len(lanes))
    rec = {
        "scene": scene + ".xml", "integrator": SWAPPED.get(name, "dopplertofpath"), "xml_params": params, "seed": seed, "time_scale": float(SCALE) if scaled else 1.0,
        "header": [l for l in out if l.startswith("#")][0],
        "columns": "idx px py sample_pos.x sample_pos.y time ray_o(3) ray_d(3) ray_maxt R G B",
        "lanes": [int(r[0]) for r in rows],
        **({"pass": [int(r[1]) for r in rows]} if all_passes else {}),
        "rows": [[int(r[2]), int(r[3])] + [float(np.float32(v)) for v in r[4:]] for r in rows],
    }
    with open(os.path.join(HERE, f"lanes2p_{name}.json" if all_passes else f"lanes_{name}.json"), "w") as f:
        json.dump(rec, f, separators=(",", ":"))
    nz = sum(1 for r in rec["rows"] if any(abs(v) > 0 for v in r[-3:]))
    print(f"{name}: {len(rows)} lanes, {nz} non-zero")


def main():
    if not os.path.exists(os.path.join(REF_RT, "replay_harness")):
        sys.exit("oracle/_ref/replay_harness missing: run `make -C oracle/ref_harness` in the build container")
    names = sys.argv[1:] or (list(CASES) + ["2p:" + n for n in PASS_CASES])
    for n in names:
        if n.startswith("2p:"):
            run_case(n[3:], PASS_CASES)
        else:
            run_case(n)
    hv = os.path.join(REF_RT, "header_vectors")
    if not sys.argv[1:] and os.path.exists(hv):
        with open(os.path.join(HERE, "header_vectors.json"), "w") as f:
            f.write(subprocess.run([hv], check=True, capture_output=True, text=True).stdout)
        print("header_vectors.json written")
    exr = "/root/reference/configs_example/scene.exr"
    if not sys.argv[1:] and os.path.exists(exr):
        try:
            os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
            import cv2
            img = cv2.imread(exr, cv2.IMREAD_UNCHANGED)[..., ::-1]   # BGR -> RGB
            np.save(os.path.join(HERE, "scene_exr.npy"), np.ascontiguousarray(img.astype(np.float16)))
            print("scene_exr.npy written", img.shape, img.dtype)
        except Exception as e:   # noqa: BLE001
            print("scene.exr not converted:", e)


if __name__ == "__main__":
    main()
