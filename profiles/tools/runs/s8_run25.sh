python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > gpurun_out/r01_bench_c2_final.json 2> gpurun_out/s8_c2_final.err; tail -c 200 gpurun_out/s8_c2_final.err; cut -c1-300 gpurun_out/r01_bench_c2_final.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_ref_c2_final.json 2>/dev/null; cut -c1-200 gpurun_out/r01_ref_c2_final.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_c2_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s8_ncu25.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -o gpurun_out/r01_render_c2_final python bench.py --workload c2 --spp 64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s8_ncu25b.log 2>&1
ls -la gpurun_out/r01_render_c2_final.ncu-rep gpurun_out/r01_launches_c2_final.csv
