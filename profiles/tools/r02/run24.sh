#!/bin/bash
# Round 2, GPU run 24: reciprocal direction clamped to +-1e18 in the ray-level API (axis-parallel probe rays cull): whole suite, speed check
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_run24_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run24_pytest.log
tail -6 gpurun_out/r02_run24_pytest.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/parity_report_gpu.json"))
for k, v in d.items():
    if k.startswith("cuda-rays"): print(k, v["nodes_max"], v["nodes_median"], v["bvh_nodes"])
PY
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[2], round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "kernel_ms", round(d["roofline"]["kernel_ms"], 3))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
for wl in c1 c2; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_exp24_${wl}.json 2> gpurun_out/r02_exp24_${wl}.err
  show gpurun_out/r02_exp24_${wl}.json "robust-rays-only $wl"
done
