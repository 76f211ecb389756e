python bench.py > gpurun_out/r01_bench_c2_v7.json 2> gpurun_out/s8_c2_v7.err; tail -c 300 gpurun_out/s8_c2_v7.err; cut -c1-330 gpurun_out/r01_bench_c2_v7.json; echo
for w in c1 c3 c4; do python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_${w}_v7.json 2> gpurun_out/s8_${w}_v7.err; cut -c1-200 gpurun_out/r01_bench_${w}_v7.json; echo; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_ref_c2_reference.json 2> gpurun_out/s8_ref.err; cat gpurun_out/r01_ref_c2_reference.json
python -c "import __graft_entry__ as g; g.smoke()"
