#!/bin/bash
# Round 2, GPU run 18: 4-wide tree for the HBM-resident walk: parity (every mode-0 test), C5 wavefront / fused, trace kernels at 6 / 5 CTAs
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests/test_wavefront.py tests/test_large_scene.py tests/test_rays.py tests/test_tof.py tests/test_multi_device.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_run18_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run18_pytest.log
tail -5 gpurun_out/r02_run18_pytest.log
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[2], round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "kernel_ms", round(d["roofline"]["kernel_ms"], 3), d["roofline"]["pipeline"], d["roofline"]["per_sample"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
run() { # tag lib env...
  tag=$1; lib=$2; shift; shift
  env DTOF_LIB=$PWD/$lib "$@" timeout 600 python bench.py --workload c5 --spp 128 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_exp18_$tag.json 2> gpurun_out/r02_exp18_$tag.err
  show gpurun_out/r02_exp18_$tag.json "$tag"
}
run wide6 mitsuba3dopplertof_b200/libdtof_b200.so
run wide5 exp_build/wide5.so
run wide6_fused mitsuba3dopplertof_b200/libdtof_b200.so DTOF_WAVEFRONT=0
run wide6_inner12 mitsuba3dopplertof_b200/libdtof_b200.so DTOF_WF_INNER=12
run wide6_nodouble mitsuba3dopplertof_b200/libdtof_b200.so DTOF_WF_DOUBLE=33
run wide6_grid3 mitsuba3dopplertof_b200/libdtof_b200.so DTOF_WF_TRACE_GRID=3
