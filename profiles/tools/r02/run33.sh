#!/bin/bash
# Round 2, GPU run 33: ncu --set full of one closest-hit and one shading launch of the C5 wavefront pipeline (shipped library, ray binning on)
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:wf_trace_kernel<\(int\)0, \(bool\)0>' -s 5 -c 1 -f -o gpurun_out/r02_wf_trace_closest_c5_final \
  python bench.py --workload c5 --spp 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_wf_closest_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:wf_shade_kernel' -s 5 -c 1 -f -o gpurun_out/r02_wf_shade_c5_final \
  python bench.py --workload c5 --spp 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_wf_shade_final.log 2>&1
ls -la gpurun_out/r02_wf_*_final.ncu-rep
