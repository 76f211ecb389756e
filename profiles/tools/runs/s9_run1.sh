ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_c2_v8.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s9_ncu1.log 2>&1
ls -la gpurun_out/r01_launches_c2_v8.csv; tail -c 300 gpurun_out/s9_ncu1.log
