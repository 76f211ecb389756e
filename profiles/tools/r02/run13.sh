#!/bin/bash
# Round 2, GPU run 13: CTA size 512 x 2 / 1024 x 1 vs 256 x 4 on C1-C4; leaf size 1 / 4 (default 2) on the SMEM scenes
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
for v in new blk512 blk1024; do
  lib=exp_build/$v.so
  [ "$v" = new ] && lib=mitsuba3dopplertof_b200/libdtof_b200.so
  for wl in c1 c2 c3 c4; do
    spp=0; [ "$wl" = c4 ] && spp=512
    DTOF_LIB=$PWD/$lib timeout 400 python bench.py --workload $wl --spp $spp --steps 5 --warmup 3 --no-cpu-baseline \
      > gpurun_out/r02_exp13_${v}_${wl}.json 2> gpurun_out/r02_exp13_${v}_${wl}.err
    show gpurun_out/r02_exp13_${v}_${wl}.json "$v $wl"
  done
done
for leaf in 1 4; do
  for wl in c2 c4; do
    spp=0; [ "$wl" = c4 ] && spp=512
    DTOF_MAX_LEAF=$leaf timeout 400 python bench.py --workload $wl --spp $spp --steps 5 --warmup 3 --no-cpu-baseline \
      > gpurun_out/r02_exp13_leaf${leaf}_${wl}.json 2> gpurun_out/r02_exp13_leaf${leaf}_${wl}.err
    show gpurun_out/r02_exp13_leaf${leaf}_${wl}.json "leaf$leaf $wl"
  done
done
