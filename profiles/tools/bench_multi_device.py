#!/usr/bin/env python
"""End-to-end throughput of ONE render through the single-process multi-device context (dtof_create_multi: what the Mitsuba
plugin uses with devices=all): host buffers in / out, one process, one host thread, n GPUs. Prints one JSON line per workload.
usage: bench_multi_device.py [c4 c5 ...]   (all visible GPUs, then GPU 0 alone)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from mitsuba3dopplertof_b200 import runtime

n_all = torch.cuda.device_count()
for wl in (sys.argv[1:] or ["c4", "c5"]):
    scene, desc = bench.load_workload(wl)
    out = {"workload": desc, "metric": "Msamples/s, e2e through host buffers, ONE render (dtof_render on a dtof_create_multi context)"}
    for n in sorted({n_all, 1}, reverse=True):
        ctx = runtime.Context(devices=list(range(n))) if n > 1 else runtime.Context(0)
        flat = ctx.upload(scene)
        params = scene.integrator.params(scene.sensor.sampler, seed=0)
        samples = flat.height * flat.width * scene.sensor.sampler.sample_count
        for _ in range(2):
            ctx.render(flat, params, both=True)
        steps = 3
        t0 = time.perf_counter()
        for _ in range(steps):
            ctx.render(flat, params, both=True)
        dt = (time.perf_counter() - t0) / steps
        out[f"n{n}"] = {"devices": ctx.device_count(), "ms_per_render": dt * 1e3, "value": samples / dt / 1e6}
        ctx.close()
    if n_all > 1:
        out["speedup"] = out[f"n{n_all}"]["value"] / out["n1"]["value"]
    print(json.dumps(out), flush=True)
