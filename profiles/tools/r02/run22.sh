#!/bin/bash
# Round 2, GPU run 22: evidence for the shipped library (1 x 1024 CTA): ncu counters of the fused kernel on C1-C4, ncu --set full
# of the headline kernel (C1) and of C2, launch list of the headline command
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
for wl in c1 c2 c3 c4; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:render_kernel -s 3 -c 1 --csv --log-file gpurun_out/r02_counters_${wl}.csv \
    python bench.py --workload $wl --spp 64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_counters_${wl}.log 2>&1
  grep -c render_kernel gpurun_out/r02_counters_${wl}.csv
done
for wl in c1 c2; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -f -o gpurun_out/r02_render_${wl}_final \
    python bench.py --workload $wl --spp 64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_${wl}_final.log 2>&1
done
ls -la gpurun_out/r02_render_c*_final.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_headline.csv \
  python bench.py --workload c1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_headline.log 2>&1
python profiles/tools/launch_table.py gpurun_out/r02_launches_headline.csv | head -20
