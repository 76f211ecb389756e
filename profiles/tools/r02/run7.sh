#!/bin/bash
# Round 2, GPU run 7: session traversal with vote-based refill (threshold 8 = in-tree, 4 / 16 = exp_build) vs base
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_rays.py tests/test_wavefront.py tests/test_path.py -m gpu -q -rs -p no:cacheprovider > gpurun_out/r02_run7_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run7_pytest.log
tail -12 gpurun_out/r02_run7_pytest.log
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[2], round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "kernel_ms", round(d["roofline"]["kernel_ms"], 3))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
for v in new refill4 refill16 base; do
  lib=exp_build/$v.so
  [ "$v" = new ] && lib=mitsuba3dopplertof_b200/libdtof_b200.so
  for wl in c1 c2 c4; do
    spp=0; [ "$wl" = c4 ] && spp=512
    DTOF_LIB=$PWD/$lib timeout 400 python bench.py --workload $wl --spp $spp --steps 5 --warmup 3 --no-cpu-baseline \
      > gpurun_out/r02_exp7_${v}_${wl}.json 2> gpurun_out/r02_exp7_${v}_${wl}.err
    show gpurun_out/r02_exp7_${v}_${wl}.json "$v $wl"
  done
done
