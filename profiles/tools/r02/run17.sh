#!/bin/bash
# Round 2, GPU run 17: C5 wavefront with evict-first (ld/st.global.cs) queue + state traffic vs plain loads / stores
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[2], round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "kernel_ms", round(d["roofline"]["kernel_ms"], 3), d["roofline"]["pipeline"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
python -m pytest tests/test_wavefront.py tests/test_large_scene.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
run() { # tag lib
  DTOF_LIB=$PWD/$2 timeout 600 python bench.py --workload c5 --spp 128 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_exp17_$1.json 2> gpurun_out/r02_exp17_$1.err
  show gpurun_out/r02_exp17_$1.json "$1"
}
run cs mitsuba3dopplertof_b200/libdtof_b200.so
run plain exp_build/wfplain.so
run cs_b mitsuba3dopplertof_b200/libdtof_b200.so
run plain_b exp_build/wfplain.so
