#!/bin/bash
# Round 2, GPU run 19: C5 wavefront, triangle prefetch at leaf discovery (-DDTOF_WF_PREFETCH) vs in-tree
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[2], round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "kernel_ms", round(d["roofline"]["kernel_ms"], 3), d["roofline"]["pipeline"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
run() { # tag lib env...
  tag=$1; lib=$2; shift; shift
  env DTOF_LIB=$PWD/$lib "$@" timeout 600 python bench.py --workload c5 --spp 128 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_exp19_$tag.json 2> gpurun_out/r02_exp19_$tag.err
  show gpurun_out/r02_exp19_$tag.json "$tag"
}
run base mitsuba3dopplertof_b200/libdtof_b200.so
run prefetch exp_build/prefetch.so
run base_b mitsuba3dopplertof_b200/libdtof_b200.so
run prefetch_b exp_build/prefetch.so
