#!/bin/bash
# Round 2, GPU run 14: 1024 x 1 as the default CTA shape: whole GPU suite, C1-C4 + C5 (fused record/stats launches use it too),
# SMEM stack vs local stack at this shape
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/r02_run14_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run14_pytest.log
tail -6 gpurun_out/r02_run14_pytest.log
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[2], round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "kernel_ms", round(d["roofline"]["kernel_ms"], 3), d["roofline"]["pipeline"], d["roofline"]["traversal_mode"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
for v in new lstack1024; do
  lib=exp_build/$v.so
  [ "$v" = new ] && lib=mitsuba3dopplertof_b200/libdtof_b200.so
  for wl in c1 c2 c3 c4; do
    spp=0; [ "$wl" = c4 ] && spp=512
    DTOF_LIB=$PWD/$lib timeout 400 python bench.py --workload $wl --spp $spp --steps 5 --warmup 3 --no-cpu-baseline \
      > gpurun_out/r02_exp14_${v}_${wl}.json 2> gpurun_out/r02_exp14_${v}_${wl}.err
    show gpurun_out/r02_exp14_${v}_${wl}.json "$v $wl"
  done
done
DTOF_WAVEFRONT=0 timeout 600 python bench.py --workload c5 --spp 128 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_exp14_new_c5_fused.json 2> gpurun_out/r02_exp14_new_c5_fused.err
show gpurun_out/r02_exp14_new_c5_fused.json "new c5 fused (128 spp)"
