N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r01_bench_c2_final_n$N.json 2> gpurun_out/s8_n$N.err; tail -c 200 gpurun_out/s8_n$N.err; cut -c1-220 gpurun_out/r01_bench_c2_final_n$N.json; echo; wc -l gpurun_out/r01_bench_c2_final_n$N.json
