#!/bin/bash
# Round 2, GPU run 21: compute_si as a real call (-DDTOF_NOINLINE_SI: fewer spills in the caller, SI through local memory)
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[2], round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "kernel_ms", round(d["roofline"]["kernel_ms"], 3))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
for v in new noinl_si; do
  lib=exp_build/$v.so
  [ "$v" = new ] && lib=mitsuba3dopplertof_b200/libdtof_b200.so
  for wl in c1 c2 c4; do
    spp=0; [ "$wl" = c4 ] && spp=512
    DTOF_LIB=$PWD/$lib timeout 400 python bench.py --workload $wl --spp $spp --steps 5 --warmup 3 --no-cpu-baseline \
      > gpurun_out/r02_exp21_${v}_${wl}.json 2> gpurun_out/r02_exp21_${v}_${wl}.err
    show gpurun_out/r02_exp21_${v}_${wl}.json "$v $wl"
  done
done
