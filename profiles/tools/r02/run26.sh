#!/bin/bash
# Round 2, GPU run 26: C5 wavefront scheduling thresholds re-swept on the round-2 node test (fetch threshold, inner threshold, double step)
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
run() { # tag env...
  tag=$1; shift
  env "$@" timeout 600 python bench.py --workload c5 --spp 128 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_exp26_$tag.json 2> gpurun_out/r02_exp26_$tag.err
  python - "gpurun_out/r02_exp26_$tag.json" "$tag" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[2], round(d["value"], 1))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
run base DTOF_WF_THRESHOLD=24
run thr16 DTOF_WF_THRESHOLD=16
run thr20 DTOF_WF_THRESHOLD=20
run thr28 DTOF_WF_THRESHOLD=28
run inner8 DTOF_WF_INNER=8
run inner12 DTOF_WF_INNER=12
run inner20 DTOF_WF_INNER=20
run inner24 DTOF_WF_INNER=24
run dbl16 DTOF_WF_DOUBLE=16
run dbl24 DTOF_WF_DOUBLE=24
run dbl28 DTOF_WF_DOUBLE=28
run b32M DTOF_WF_BATCH=33554432
