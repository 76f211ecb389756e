#!/bin/bash
# Round 2, GPU run 30: wavefront pipeline on the shared-memory scenes with ray binning forced on (DTOF_WF_BINS=2) vs fused
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
run() { # tag wl spp env...
  tag=$1; wl=$2; spp=$3; shift; shift; shift
  env "$@" timeout 600 python bench.py --workload $wl --spp $spp --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_exp30_$tag.json 2> gpurun_out/r02_exp30_$tag.err
  python - "gpurun_out/r02_exp30_$tag.json" "$tag" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[2], round(d["value"], 1), d["roofline"]["pipeline"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
for wl in c1 c2 c4; do
  spp=0; [ "$wl" = c4 ] && spp=512
  run wf_bins2_$wl $wl $spp DTOF_WAVEFRONT=1 DTOF_WF_BINS=2
  run wf_bins0_$wl $wl $spp DTOF_WAVEFRONT=1 DTOF_WF_BINS=0
done
