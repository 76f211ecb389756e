#!/bin/bash
# Round 2, GPU run 28: thresholds with ray binning on (C5, 128 spp)
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
run() { # tag env...
  tag=$1; shift
  env "$@" timeout 600 python bench.py --workload c5 --spp 128 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_exp28_$tag.json 2> gpurun_out/r02_exp28_$tag.err
  python - "gpurun_out/r02_exp28_$tag.json" "$tag" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[2], round(d["value"], 1))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
run base DTOF_WF_BINS=1
run inner8 DTOF_WF_INNER=8
run inner12 DTOF_WF_INNER=12
run inner20 DTOF_WF_INNER=20
run inner24 DTOF_WF_INNER=24
run thr20 DTOF_WF_THRESHOLD=20
run thr28 DTOF_WF_THRESHOLD=28
run dbl16 DTOF_WF_DOUBLE=16
run dbl24 DTOF_WF_DOUBLE=24
run grid3 DTOF_WF_TRACE_GRID=3
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio \
  --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_c5_bins.csv \
  python bench.py --workload c5 --spp 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_c5_bins.log 2>&1
python profiles/tools/launch_table.py gpurun_out/r02_launches_c5_bins.csv | grep -A3 "wf_trace\|wf_shade"
