#!/bin/bash
# diagnostic: C5 regression (38 Msamples/s instead of ~840): which kernel?
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
DTOF_WAVEFRONT=0 timeout 300 python bench.py --workload c5 --spp 32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_diag_c5_fused.json 2> gpurun_out/r02_diag_c5_fused.err
python -c "
import json; d=json.load(open('gpurun_out/r02_diag_c5_fused.json')); print('fused c5', d['value'], d['roofline']['kernel_ms'])"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_diag_c5_launches.csv \
  python bench.py --workload c5 --spp 32 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_diag_c5_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_diag_c5_launches.csv")) if len(r) > 10]
hdr = rows[0]; ik = hdr.index("Kernel Name"); im = hdr.index("Metric Name"); iv = hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    if r[im] == "gpu__time_duration.sum":
        agg[r[ik][:60]].append(float(r[iv].replace(",", "")))
for k, v in agg.items():
    print(k, len(v), "mean", sum(v) / len(v), "max", max(v))
PY
