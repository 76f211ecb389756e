#!/bin/bash
# Round 2, GPU run 10: plugin tests after the first-chunk fix; ncu launch list of the C5 wavefront pipeline (current build, 64 spp)
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests/test_mitsuba_plugin.py -m gpu -q -rs -p no:cacheprovider > gpurun_out/r02_run10_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run10_pytest.log
tail -4 gpurun_out/r02_run10_pytest.log
timeout 1200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_c5_wavefront.csv \
  python bench.py --workload c5 --spp 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_c5_ncu.log 2>&1
python profiles/tools/launch_table.py gpurun_out/r02_launches_c5_wavefront.csv > gpurun_out/r02_launches_c5_wavefront_table.txt 2>&1
tail -60 gpurun_out/r02_launches_c5_wavefront_table.txt
