#!/bin/bash
# Round 2, GPU run 12: CTA size A/B (128 x 8, 512 x 2 vs 256 x 4), wavefront pipeline on the SMEM scenes, ncu counters of the
# fused kernel on C1 / C3 / C4 (for profiles/r02_counters.json), compute-sanitizer memcheck + racecheck (result kept)
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
for v in new blk128 blk512; do
  lib=exp_build/$v.so
  [ "$v" = new ] && lib=mitsuba3dopplertof_b200/libdtof_b200.so
  for wl in c1 c2 c4; do
    spp=0; [ "$wl" = c4 ] && spp=512
    DTOF_LIB=$PWD/$lib timeout 400 python bench.py --workload $wl --spp $spp --steps 5 --warmup 3 --no-cpu-baseline \
      > gpurun_out/r02_exp12_${v}_${wl}.json 2> gpurun_out/r02_exp12_${v}_${wl}.err
    show gpurun_out/r02_exp12_${v}_${wl}.json "$v $wl"
  done
done
for wl in c1 c2 c4; do
  spp=0; [ "$wl" = c4 ] && spp=512
  DTOF_WAVEFRONT=1 timeout 400 python bench.py --workload $wl --spp $spp --steps 5 --warmup 3 --no-cpu-baseline \
      > gpurun_out/r02_exp12_wf_${wl}.json 2> gpurun_out/r02_exp12_wf_${wl}.err
  show gpurun_out/r02_exp12_wf_${wl}.json "wavefront $wl"
done
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
for wl in c1 c2 c3 c4; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:render_kernel -s 3 -c 1 --csv --log-file gpurun_out/r02_counters_${wl}.csv \
    python bench.py --workload $wl --spp 64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_counters_${wl}.log 2>&1
  grep -c render_kernel gpurun_out/r02_counters_${wl}.csv
done
export PATH=/usr/local/cuda/bin:$PATH
{
echo "== compute-sanitizer memcheck: wavefront pipeline (c4, c5 slab room, two passes, sharded), fused kernel parity subset"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_wavefront.py -x -q -p no:cacheprovider -k "c4_domino or c5_slabroom or two_pass or sharded" 2>&1 | tail -6
echo "memcheck wavefront rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "c1_ or c4_" 2>&1 | tail -6
echo "memcheck fused rc=$?"
echo "== compute-sanitizer racecheck: fused kernel with the shared-memory stack (c2)"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "c2_arealight" 2>&1 | tail -6
echo "racecheck rc=$?"
} > gpurun_out/r02_compute_sanitizer.txt 2>&1
cat gpurun_out/r02_compute_sanitizer.txt
