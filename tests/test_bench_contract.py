"""bench.py's contract, as far as it can be checked without a GPU: every workload of BASELINE.json loads and flattens, the
plan of the default run names the configurations the driver is meant to see, the reference arm leaves the non-zero ranks
without work, and the committed bench records carry the keys the contract asks for."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_every_workload_loads_on_the_host(monkeypatch):
    import bench
    monkeypatch.setenv("DTOF_BENCH_MESH_N", "24")          # C5 with a 6 912-triangle sphere: same scene file, small mesh
    from mitsuba3dopplertof_b200 import runtime
    for name in sorted(bench.WORKLOADS):
        scene, desc = bench.load_workload(name)
        flat = scene.flatten()
        info = runtime.scene_info(flat)                    # host half of the upload: validation + BVH build, no GPU
        assert info.n_triangles == flat.n_triangles > 0 and desc.startswith(name.upper())
        s = scene.sensor.sampler
        assert s.kind == "correlated" and s.sample_count >= 1024
    assert bench.HEADLINE == "c1"


def test_reference_arm_gives_other_ranks_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_bench_without_a_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "c1", "--steps", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_committed_bench_records_follow_the_contract(n):
    path = os.path.join(ROOT, "profiles", f"r02_bench_default_n{n}.json")
    if not os.path.exists(path):
        pytest.skip("no record for this GPU count")
    d = json.loads(open(path).read().strip().split("\n")[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "workloads"):
        assert k in d, k
    assert d["n_gpus"] == n and d["metric"] == "Msamples/s" and d["vs_baseline"] is None and d["gpu_launches"] > 0
    assert d["config"]["workload"].startswith("C1 ") and d["warmup"] >= 3
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert 0 < d["e2e"]["value"] <= 1.02 * d["value"] and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0
    for r in [d["roofline"]] + [w["roofline"] for w in d["workloads"]]:
        assert r["bound"] in ("issue", "hbm") and 0 < r["frac"] and {"achieved", "peak", "unit", "traffic", "issue", "hbm"} <= set(r)
    names = [w["workload"] for w in d["workloads"]]
    assert names == (["c1", "c2", "c3", "c4", "c5"] if n == 1 else ["c1", "c2", "c4", "c5"])
    if n > 1:
        by = {w["workload"]: w for w in d["workloads"]}
        assert by["c4"]["scaling"] == "strong" and by["c4"]["sharding"] == "tiles"
        assert by["c5"]["scaling"] == "strong" and by["c5"]["sharding"] == "slots"
    if n == 1:
        assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
