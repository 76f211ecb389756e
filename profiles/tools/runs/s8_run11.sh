# full GPU suite with the final library
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
# launch list of the wavefront pipeline on C5 (same command as the bench line below, reduced spp)
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_c5_wavefront.csv python bench.py --workload c5 --spp 64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s8_ncu11a.log 2>&1
tail -c 150 gpurun_out/s8_ncu11a.log; echo
# full captures: closest trace (bounce 2), shade (bounce 1), shadow trace (bounce 1)
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:wf_trace_kernel<.*0, .*0>' -s 6 -c 1 -o gpurun_out/r01_wf_trace_closest_c5 python bench.py --workload c5 --spp 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/s8_ncu11b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wf_shade_kernel -s 5 -c 1 -o gpurun_out/r01_wf_shade_c5 python bench.py --workload c5 --spp 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/s8_ncu11c.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:wf_trace_kernel<.*0, .*1>' -s 5 -c 1 -o gpurun_out/r01_wf_trace_shadow_c5 python bench.py --workload c5 --spp 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/s8_ncu11d.log 2>&1
ls -la gpurun_out/*.ncu-rep
# bench lines
python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_c5_v7.json 2> gpurun_out/s8_c5_v7.err; tail -c 300 gpurun_out/s8_c5_v7.err; cut -c1-400 gpurun_out/r01_bench_c5_v7.json
DTOF_WAVEFRONT=0 python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_c5_v7_fused.json 2> gpurun_out/s8_c5_v7f.err; cut -c1-200 gpurun_out/r01_bench_c5_v7_fused.json
