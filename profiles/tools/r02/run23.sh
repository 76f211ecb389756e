#!/bin/bash
# Round 2, GPU run 23 (8 GPUs): one render through the single-process multi-device context, C4 and C5 at full size
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python profiles/tools/bench_multi_device.py c1 c4 c5 > gpurun_out/r02_multi_device_e2e.jsonl 2> gpurun_out/r02_multi_device_e2e.err
cat gpurun_out/r02_multi_device_e2e.jsonl | cut -c1-700; tail -3 gpurun_out/r02_multi_device_e2e.err
