#!/bin/bash
# Round 2, multi-GPU run: bash run_multi.sh N -- multi-device tests (C ABI context over N GPUs, plugin on all GPUs) and the
# default bench line at N ranks (C1 / C2 weak, C4 strong by tiles, C5 strong by sample slots), launched as the driver does
set -u
N=${1:-2}
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m pytest tests/test_multi_device.py tests/test_mitsuba_plugin.py -m gpu -q -rs -p no:cacheprovider -k "multi or all_gpus or devices" > gpurun_out/r02_multi_n${N}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_multi_n${N}_pytest.log
tail -4 gpurun_out/r02_multi_n${N}_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N \
  > gpurun_out/r02_bench_default_n${N}.json 2> gpurun_out/r02_bench_default_n${N}.err
grep -E "^\[bench\]" gpurun_out/r02_bench_default_n${N}.err | sort -u
