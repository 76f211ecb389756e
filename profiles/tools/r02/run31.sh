#!/bin/bash
# Round 2, GPU run 31: entries of a CTA's queue share ordered by direction octant (-DDTOF_WF_OCTANTS) vs in-tree
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
DTOF_LIB=$PWD/exp_build/octants.so timeout 600 python -m pytest tests/test_wavefront.py tests/test_large_scene.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
run() { # tag lib
  DTOF_LIB=$PWD/$2 timeout 600 python bench.py --workload c5 --spp 128 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_exp31_$1.json 2> gpurun_out/r02_exp31_$1.err
  python - "gpurun_out/r02_exp31_$1.json" "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[2], round(d["value"], 1))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
run base mitsuba3dopplertof_b200/libdtof_b200.so
run octants exp_build/octants.so
run base_b mitsuba3dopplertof_b200/libdtof_b200.so
run octants_b exp_build/octants.so
