#!/usr/bin/env python
"""Writes profiles/r02_counters.json: per workload the counters ncu measured on the shipped library, per sample.
Inputs (gpurun_out/, produced by profiles/tools/r02/run22.sh and run10.sh):
  r02_counters_<wl>.csv            one render_kernel launch of `bench.py --workload <wl> --spp 64` (fused kernel)
  r02_launches_c5_wavefront.csv    launch list of `bench.py --workload c5 --spp 64` (wavefront pipeline, every wf_* launch)
bench.py reads the file for roofline.traffic / issue.executed_* (numbers measured under ncu are never bench values; these are
per-sample COUNTS, which do not depend on the clock or on serialisation)."""
import collections, csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
G = os.path.join(ROOT, "gpurun_out")
PIX = {"c1": 256 * 256, "c2": 512 * 512, "c3": 512 * 512, "c4": 1024 * 1024, "c5": 2048 * 2048}
SPP = 64


def rows(path):
    r = list(csv.reader(open(path)))
    h = [i for i, x in enumerate(r) if x and x[0] == "ID"][0]
    ix = {n: i for i, n in enumerate(r[h])}
    for x in r[h + 1:]:
        if len(x) >= len(r[h]):
            yield x[ix["ID"]], x[ix["Kernel Name"]], x[ix["Metric Name"]], float(x[ix["Metric Value"]].replace(",", ""))


out = {"_doc": __doc__.split("\n")[0] + " ncu: --clock-control none; samples = width x height x 64 spp of the captured render."}
for wl in ("c1", "c2", "c3", "c4"):
    p = os.path.join(G, f"r02_counters_{wl}.csv")
    if not os.path.exists(p):
        continue
    m = {n: v for _, k, n, v in rows(p)}
    n = PIX[wl] * SPP
    out[wl] = {
        "source": f"ncu --metrics ... -k regex:render_kernel -s 3 -c 1 python bench.py --workload {wl} --spp 64 (profiles/tools/r02/run22.sh)",
        "samples": n, "kernel_ms_under_ncu": m["gpu__time_duration.sum"] / 1e6,
        "bytes_per_sample": (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / n,
        "dram_read_bytes_per_sample": m["dram__bytes_read.sum"] / n, "dram_write_bytes_per_sample": m["dram__bytes_write.sum"] / n,
        "warp_inst_per_sample": m["smsp__inst_executed.sum"] / n, "thread_inst_per_sample": m["smsp__thread_inst_executed.sum"] / n,
        "lanes_per_inst": m["smsp__thread_inst_executed.sum"] / m["smsp__inst_executed.sum"],
        "issue_active_pct": m["smsp__issue_active.avg.pct_of_peak_sustained_active"],
        "warps_active_pct": m["sm__warps_active.avg.pct_of_peak_sustained_active"],
        "note": "DRAM bytes are almost all write-back of local memory (register spills at 64 registers / thread); scene data is shared-memory resident",
    }
p = os.path.join(G, "r02_launches_c5_wavefront.csv")
if os.path.exists(p):
    per = collections.defaultdict(lambda: collections.defaultdict(list))
    for _id, k, n, v in rows(p):
        per[k.split("(")[0].replace("void ", "")][n].append(v)
    # the capture holds the warm-up render and most of a second one: count whole renders by the splat launches (one per batch and pass)
    batches = PIX["c5"] * SPP / (16 * 1024 * 1024)
    n_splat = len(per["dtof::wf_splat_kernel"]["gpu__time_duration.sum"])
    renders = n_splat / batches
    stages = {}
    tot = collections.Counter()
    for k, m in per.items():
        if not k.startswith("dtof::wf_"):
            continue
        launches = len(m["gpu__time_duration.sum"])
        # per-launch means x launches of ONE render (launch counts per render: generate / splat = batches, per-bounce kernels = the rest)
        per_render = launches / renders
        st = {"launches_per_render": per_render}
        for name, key in (("dram_read_bytes_per_sample", "dram__bytes_read.sum"), ("dram_write_bytes_per_sample", "dram__bytes_write.sum"),
                          ("warp_inst_per_sample", "smsp__inst_executed.sum")):
            st[name] = sum(m[key]) / launches * per_render / (PIX["c5"] * SPP)
            tot[name] += st[name]
        st["lanes_per_inst"] = sum(m["smsp__thread_inst_executed_per_inst_executed.ratio"]) / launches
        st["issue_active_pct_alone"] = sum(m["smsp__issue_active.avg.pct_of_peak_sustained_active"]) / launches
        st["ms_alone_per_render"] = sum(m["gpu__time_duration.sum"]) / launches * per_render / 1e6
        stages[k] = st
    thread = sum(s["warp_inst_per_sample"] * s["lanes_per_inst"] for s in stages.values())
    out["c5_wavefront"] = {
        "source": "ncu launch list of python bench.py --workload c5 --spp 64 (profiles/r02_launches_c5_wavefront.csv, profiles/tools/r02/run29.sh)",
        "samples": PIX["c5"] * SPP, "renders_in_capture": renders,
        "bytes_per_sample": tot["dram_read_bytes_per_sample"] + tot["dram_write_bytes_per_sample"],
        "dram_read_bytes_per_sample": tot["dram_read_bytes_per_sample"], "dram_write_bytes_per_sample": tot["dram_write_bytes_per_sample"],
        "warp_inst_per_sample": tot["warp_inst_per_sample"], "thread_inst_per_sample": thread,
        "lanes_per_inst": thread / tot["warp_inst_per_sample"], "issue_active_pct": None,
        "stages": stages,
        "note": "issue_active of a stage is measured with the kernel ALONE (ncu serialises); live, four batches overlap: "
                "warp_inst_per_sample x samples/s / (148 x 4 x SM clock) is the pipeline's issue-slot utilisation",
    }
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_counters.json"), "w"), indent=1)
for k, v in out.items():
    if isinstance(v, dict):
        print(k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items() if a not in ("stages", "source", "note")})
