"""How skewed is the per-ray traversal cost inside a small scene? Rays with origins inside the scene's bounds and uniform
directions (a stand-in for bounce rays), cost proxy = 50 x nodes + 45 x triangles (+ 135 per instance entry, seen as instance
hits only). Prints the mean, the mean of the per-32-group maximum (what a warp of a fused kernel pays) and their ratio, split by
whether the closest hit is on an animated instance."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util as gu
import mitsuba3dopplertof_b200 as dt
from mitsuba3dopplertof_b200 import _abi, runtime
for name, kw in (("c1_example.xml", {}), ("c2_arealight.xml", {}), ("c4_domino.xml", {"w_g": 150})):
    scene = dt.load_file(os.path.join(gu.SCENES, name), resx=16, resy=16, spp=4, **kw)
    ctx = runtime.Context(0)
    flat = ctx.upload(scene)
    info = runtime.scene_info(flat)
    # scene bounds from the flattened static meshes
    pts = np.concatenate([np.ctypeslib.as_array(flat.meshes[i].positions, (flat.meshes[i].n_vertices * 3,)).reshape(-1, 3)
                          for i in range(flat.desc.n_meshes)])
    lo, hi = pts.min(0), pts.max(0)
    rng = np.random.default_rng(5)
    n = 1 << 16
    r = np.zeros(n, _abi.RAY_DTYPE)
    r["o"] = (lo + (hi - lo) * rng.uniform(0.05, 0.95, (n, 3))).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    r["d"] = d; r["tmax"] = np.float32(3.4e38); r["time"] = rng.uniform(0, 0.0015, n).astype(np.float32)
    g = ctx.trace_rays(r)
    cost = 50.0 * g["nodes_visited"] + 45.0 * g["tris_tested"]
    grp = cost.reshape(-1, 32)
    inst = g["instance"] >= 0
    print(name, "nodes", info.n_nodes, "| mean nodes", g["nodes_visited"].mean().round(2), "tris", g["tris_tested"].mean().round(2),
          "| cost mean", cost.mean().round(0), "mean of max over 32", grp.max(1).mean().round(0), "ratio", (grp.max(1).mean() / cost.mean()).round(2),
          "| instance-hit rays", inst.mean().round(3), "their mean cost", cost[inst].mean().round(0), "others", cost[~inst].mean().round(0),
          "| p50 / p90 / p99 / max", np.percentile(cost, [50, 90, 99, 100]).round(0))
    ctx.close()
