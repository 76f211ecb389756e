export PATH=/usr/local/cuda/bin:$PATH
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_wavefront.py -x -q -k "c4_domino or c5_slabroom or threshold or unbounded or two_pass or sharded" 2>&1 | tail -8
echo "memcheck rc=$?"
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_wavefront.py -x -q -k "c2_arealight and 1" 2>&1 | tail -6
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_tof.py -x -q -k "animation" 2>&1 | tail -4
# stdout purity of the bench under torchrun is checked on the 1-GPU box with world size 1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 1 --steps 1 --warmup 3 --no-cpu-baseline 2>/dev/null | wc -l
