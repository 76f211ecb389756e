import os, sys
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import mitsuba3dopplertof_b200 as dt
from mitsuba3dopplertof_b200 import runtime
import oracle_lib
scene = dt.load_file("tests/scenes/c12_roughconductor.xml", resx=40, resy=40, spp=32, hetero_frequency=0.0, max_depth=6)
params = scene.integrator.params(scene.sensor.sampler, seed=5)
ctx = runtime.Context(0)
flat = ctx.upload(scene)
lanes = np.arange(0, 40 * 40 * 32, dtype=np.uint64)
g = ctx.trace_samples(params, lanes)
o = oracle_lib.OracleScene(flat).trace(params, lanes)
d = np.abs(g["rgb"].astype(np.float64) - o["rgb"]).max(axis=1)
s = np.maximum(np.abs(o["rgb"]).max(axis=1), 1e-2)
rel = d / s
print("lanes", len(lanes), "median", np.median(rel), "p99", np.quantile(rel, 0.99), "p999", np.quantile(rel, 0.999), "max", rel.max())
bad = np.argsort(-rel)[:8]
for b in bad:
    print(int(lanes[b]), rel[b], g["rgb"][b], o["rgb"][b], g["depth"][b], o["depth"][b], g["rng_draws"][b], o["rng_draws"][b])
print("scale of sample values", np.abs(o["rgb"]).max(), "depth mismatch", (g["depth"] != o["depth"]).sum())
