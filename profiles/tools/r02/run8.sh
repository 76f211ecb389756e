#!/bin/bash
# Round 2, GPU run 8: test_rays diagnostics + drop-in plugin tests; the default bench line (C1 headline + workloads matrix)
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rs -p no:cacheprovider > gpurun_out/r02_run8_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run8_pytest.log
tail -5 gpurun_out/r02_run8_pytest.log
python bench.py > gpurun_out/r02_bench_default_n1.json 2> gpurun_out/r02_bench_default_n1.err
grep -E "^\[bench\]" gpurun_out/r02_bench_default_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r02_bench_reference_n1.err
cat gpurun_out/r02_bench_reference_n1.json | cut -c1-600
