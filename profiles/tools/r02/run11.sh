#!/bin/bash
# Round 2, GPU run 11: ncu --set full of one closest-hit and one shadow wf_trace launch of C5 (bounce >= 1)
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:wf_trace_kernel<\(int\)0, \(bool\)0>' -s 5 -c 1 -f -o gpurun_out/r02_wf_trace_closest_c5 \
  python bench.py --workload c5 --spp 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_wf_closest.log 2>&1
tail -3 gpurun_out/r02_ncu_wf_closest.log
ls -la gpurun_out/*.ncu-rep
