#!/bin/bash
# Round 2, GPU run 16: C5 wavefront: L2 access-policy window over the BVH nodes (DTOF_L2_PERSIST), trace CTAs per SM
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
python - <<'PY'
import torch
from cuda import cudart
for a in ("cudaDevAttrMaxPersistingL2CacheSize", "cudaDevAttrMaxAccessPolicyWindowSize", "cudaDevAttrL2CacheSize"):
    print(a, cudart.cudaDeviceGetAttribute(getattr(cudart.cudaDeviceAttr, a), 0))
PY
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python bench.py --workload c5 --spp 128 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_exp16_$tag.json 2> gpurun_out/r02_exp16_$tag.err
  show gpurun_out/r02_exp16_$tag.json "$tag"
}
run persist0 DTOF_L2_PERSIST=0
run persist1 DTOF_L2_PERSIST=1
run persist0_b DTOF_L2_PERSIST=0
run persist1_b DTOF_L2_PERSIST=1
run persist1_grid3 DTOF_L2_PERSIST=1 DTOF_WF_TRACE_GRID=3
run persist1_grid4 DTOF_L2_PERSIST=1 DTOF_WF_TRACE_GRID=4
run persist1_b8M DTOF_L2_PERSIST=1 DTOF_WF_BATCH=8388608
