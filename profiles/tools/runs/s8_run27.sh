N=$1
for sc in strong-slots strong-tiles; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3 --scaling $sc --no-cpu-baseline > gpurun_out/r01_bench_c2_${sc}_n$N.json 2> gpurun_out/s8_${sc}_n$N.err; tail -c 200 gpurun_out/s8_${sc}_n$N.err; cut -c1-200 gpurun_out/r01_bench_c2_${sc}_n$N.json; echo
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --workload c5 --spp 256 --steps 2 --warmup 3 --scaling strong-slots --no-cpu-baseline > gpurun_out/r01_bench_c5_strong-slots_n$N.json 2> gpurun_out/s8_c5s_n$N.err; tail -c 200 gpurun_out/s8_c5s_n$N.err; cut -c1-200 gpurun_out/r01_bench_c5_strong-slots_n$N.json; echo
