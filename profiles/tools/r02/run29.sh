#!/bin/bash
# Round 2, GPU run 29: ray binning + fetch threshold 20 as defaults: whole suite, C5 at full size, launch list for the counters
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_run29_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run29_pytest.log
tail -4 gpurun_out/r02_run29_pytest.log
timeout 900 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c5_bins.json 2> gpurun_out/r02_bench_c5_bins.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c5_bins.json')); print('c5 full', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
timeout 1200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_c5_wavefront.csv \
  python bench.py --workload c5 --spp 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_c5_ncu.log 2>&1
python profiles/tools/launch_table.py gpurun_out/r02_launches_c5_wavefront.csv > gpurun_out/r02_launches_c5_wavefront_table.txt 2>&1
grep -c wf_ gpurun_out/r02_launches_c5_wavefront.csv
