#!/bin/bash
# Round 2, GPU run 5: whole GPU suite after the drop-in fix; fresh ncu --set full capture of the headline kernel (C2, 64 spp)
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_run5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run5_pytest.log
tail -8 gpurun_out/r02_run5_pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -f -o gpurun_out/r02_render_c2_v10 \
  python bench.py --workload c2 --spp 64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_c2_v10.log 2>&1
ls -la gpurun_out/r02_render_c2_v10.ncu-rep
