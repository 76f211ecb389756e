"""Per-category node counts of dtof_trace_rays on the test_rays.py ray set (diagnostic)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_rays as tr
from mitsuba3dopplertof_b200 import runtime
scene = tr._scene("c1_example.xml", resx=16, resy=16, spp=4)
ctx = runtime.Context(0)
flat = ctx.upload(scene)
rays = tr._rays(np.random.default_rng(11), 1 << 15, -1.2, 2.2, 0.0015)
got = ctx.trace_rays(rays)
k = rays.size // 8
n = got["nodes_visited"].astype(float)
print("lib", os.environ.get("DTOF_LIB"), "hit frac", got["hit"].mean(), "median all", np.median(n), "mean", n.mean())
for c in range(8):
    s = slice(c * k, (c + 1) * k)
    print(" category", c, "median", np.median(n[s]), "mean", round(n[s].mean(), 2), "max", n[s].max(), "hit", round(got["hit"][s].mean(), 3), "tris", round(got["tris_tested"][s].mean(), 2))
print(np.bincount(n.astype(int))[:14])
