#!/bin/bash
# Round 2, GPU run 27: ray binning in the wavefront pipeline (class A = crosses an animated instance's box, queue filled from both ends)
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_wavefront.py tests/test_large_scene.py tests/test_tof.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
run() { # tag env...
  tag=$1; shift
  env "$@" timeout 600 python bench.py --workload c5 --spp 128 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_exp27_$tag.json 2> gpurun_out/r02_exp27_$tag.err
  python - "gpurun_out/r02_exp27_$tag.json" "$tag" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[2], round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
run bins1 DTOF_WF_BINS=1
run bins0 DTOF_WF_BINS=0
run bins1_b DTOF_WF_BINS=1
run bins0_b DTOF_WF_BINS=0
