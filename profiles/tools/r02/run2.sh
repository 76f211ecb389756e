#!/bin/bash
# Round 2, GPU run 2: centre/half-extent node test in all three walks. (a) GPU suite, (b) A/B against the round-1 walk,
# (c) ncu --set full of the fused kernel on C2 (64 spp), (d) C5 (wavefront) A/B at 128 spp.
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_run2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run2_pytest.log
tail -5 gpurun_out/r02_run2_pytest.log
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[2], round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "kernel_ms", round(d["roofline"]["kernel_ms"], 3))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
for v in base new; do
  lib=exp_build/$v.so
  [ "$v" = new ] && lib=mitsuba3dopplertof_b200/libdtof_b200.so
  for wl in c1 c2 c3 c4 c5; do
    spp=0; [ "$wl" = c4 ] && spp=512; [ "$wl" = c5 ] && spp=128
    DTOF_LIB=$PWD/$lib timeout 400 python bench.py --workload $wl --spp $spp --steps 5 --warmup 3 --no-cpu-baseline \
      > gpurun_out/r02_exp2_${v}_${wl}.json 2> gpurun_out/r02_exp2_${v}_${wl}.err
    show gpurun_out/r02_exp2_${v}_${wl}.json "$v $wl"
  done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -f -o gpurun_out/r02_render_c2_v9 \
  python bench.py --workload c2 --spp 64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_c2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
