#!/bin/bash
# Round 2, GPU run 9 (2 GPUs): multi-device context behind the C ABI + the drop-in plugin on all GPUs; default bench at N = 2
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_multi_device.py tests/test_mitsuba_plugin.py -m gpu -q -rs -p no:cacheprovider > gpurun_out/r02_run9_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run9_pytest.log
tail -8 gpurun_out/r02_run9_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 \
  > gpurun_out/r02_bench_default_n2.json 2> gpurun_out/r02_bench_default_n2.err
grep -E "^\[bench\]" gpurun_out/r02_bench_default_n2.err; tail -3 gpurun_out/r02_bench_default_n2.err
