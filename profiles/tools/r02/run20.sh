#!/bin/bash
# Round 2, GPU run 20: per-thread instance-transform cache in the shared-memory walk vs -DDTOF_NO_INST_CACHE; parity subset first
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_wavefront.py tests/test_velocity.py tests/test_tof.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[2], round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "kernel_ms", round(d["roofline"]["kernel_ms"], 3), d["roofline"]["pipeline"], d["roofline"]["traversal_mode"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
for v in new nocache; do
  lib=exp_build/$v.so
  [ "$v" = new ] && lib=mitsuba3dopplertof_b200/libdtof_b200.so
  for wl in c1 c2 c3 c4; do
    spp=0; [ "$wl" = c4 ] && spp=512
    DTOF_LIB=$PWD/$lib timeout 400 python bench.py --workload $wl --spp $spp --steps 5 --warmup 3 --no-cpu-baseline \
      > gpurun_out/r02_exp20_${v}_${wl}.json 2> gpurun_out/r02_exp20_${v}_${wl}.err
    show gpurun_out/r02_exp20_${v}_${wl}.json "$v $wl"
  done
done
