DTOF_WAVEFRONT=1 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:wf_trace_kernel<.*0, .*0>' -s 6 -c 1 -o gpurun_out/s8_wf_trace_closest python bench.py --workload c5 --spp 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/s8_ncu5.log 2>&1
tail -c 200 gpurun_out/s8_ncu5.log
ls -la gpurun_out/
