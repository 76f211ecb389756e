#!/usr/bin/env python
"""Condenses an ncu report into the JSON summary kept under profiles/ (usage: ncu_summary.py rep out.json "kernel desc" "capture cmd")."""
import csv, io, json, subprocess, sys
rep, out, desc, cmd = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sass__inst_executed_shared_loads", "sass__inst_executed_global_loads",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__inst_executed_op_global_red.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic"]
m = {}
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            m[w] = {"value": vals[i], "unit": units[i]}
st = {}
for i, h in enumerate(hdr):
    if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
        st[h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = round(float(vals[i]), 3)
json.dump({"capture": cmd, "kernel": desc, "metrics": m, "warps_stalled_per_issue": dict(sorted(st.items(), key=lambda kv: -kv[1]))},
          open(out, "w"), indent=1)
for k, v in m.items():
    print(k, v["value"], v["unit"])
print({k: v for k, v in list(sorted(st.items(), key=lambda kv: -kv[1]))[:8]})
