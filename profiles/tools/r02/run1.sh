#!/bin/bash
# Round 2, GPU run 1: (a) whole GPU suite on the new shared-memory traversal (SMEM stack, fma slab, ABI v8 pass replay),
# (b) A/B of the fused kernel against the round-1 library on the shared-memory workloads.
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_run1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run1_pytest.log
tail -5 gpurun_out/r02_run1_pytest.log
for v in base new new3 new5; do
  lib=exp_build/$v.so
  [ "$v" = new ] && lib=mitsuba3dopplertof_b200/libdtof_b200.so
  for wl in c1 c2 c3 c4; do
    spp=0; [ "$wl" = c4 ] && spp=512
    DTOF_LIB=$PWD/$lib timeout 300 python bench.py --workload $wl --spp $spp --steps 5 --warmup 3 --no-cpu-baseline \
      > gpurun_out/r02_exp1_${v}_${wl}.json 2> gpurun_out/r02_exp1_${v}_${wl}.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_exp1_${v}_${wl}.json")); print("${v} ${wl}", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "kernel_ms", round(d["roofline"]["kernel_ms"],3))
except Exception as e: print("${v} ${wl} FAILED", e)
PY
  done
done
