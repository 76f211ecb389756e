cd "$GRAFT_REPO_ROOT"
start=$(date +%s)
python bench.py --gpus 1 --steps 25 --warmup 5 > gpurun_out/r02_bench_driverlike_n1.json 2> gpurun_out/r02_bench_driverlike_n1.err
echo "rc=$? wall=$(( $(date +%s) - start )) s"
grep "^\[bench\]" gpurun_out/r02_bench_driverlike_n1.err
wc -l gpurun_out/r02_bench_driverlike_n1.json
start=$(date +%s)
python bench.py --impl reference --gpus 1 --steps 25 --warmup 5 > gpurun_out/r02_bench_driverlike_ref.json 2> gpurun_out/r02_bench_driverlike_ref.err
echo "ref rc=$? wall=$(( $(date +%s) - start )) s"; cut -c1-160 gpurun_out/r02_bench_driverlike_ref.json
