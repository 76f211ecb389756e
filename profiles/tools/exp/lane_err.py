import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
import golden_util as gu, oracle_lib
from mitsuba3dopplertof_b200 import runtime
ctx = runtime.Context(0)
for name in gu.case_names():
    scene, params, ref = gu.load_case(name)
    flat = ctx.upload(scene)
    rec = ctx.trace_samples(params, ref["lanes"])
    orc = oracle_lib.OracleScene(flat, 0).trace(params, ref["lanes"])
    d = np.abs(rec["rgb"].astype(np.float64) - orc["rgb"]).max(axis=1)
    scale = np.maximum(np.abs(orc["rgb"]).max(axis=1), gu.ABS_FLOOR)
    rel = d / scale
    dref = np.abs(rec["rgb"].astype(np.float64) - ref["rgb"]).max(axis=1) / np.maximum(np.abs(ref["rgb"]).max(axis=1), gu.ABS_FLOOR)
    q = np.quantile(rel, [0.5, 0.9, 0.99, 1.0])
    qr = np.quantile(dref, [0.5, 0.9, 0.99, 1.0])
    print(f"{name:28s} vs oracle  p50 {q[0]:.1e} p90 {q[1]:.1e} p99 {q[2]:.1e} max {q[3]:.1e} | vs reference p50 {qr[0]:.1e} p90 {qr[1]:.1e} p99 {qr[2]:.1e} max {qr[3]:.1e}  depth== {np.mean(rec['depth']==orc['depth']):.4f}")
    if name in ("c1_trap_mirror", "c4_domino"):
        w = np.argsort(-rel)[:6]
        for i in w:
            print("   lane", ref["lanes"][i], "rel", f"{rel[i]:.2e}", "gpu", rec["rgb"][i], "orc", orc["rgb"][i], "L", rec["path_length"][i], orc["path_length"][i],
                  "dL", float(rec["path_length"][i]) - float(orc["path_length"][i]), "depth", rec["depth"][i])
        dl = np.abs(rec["path_length"].astype(np.float64) - orc["path_length"])
        print("   path_length abs diff quantiles", np.quantile(dl, [0.5, 0.9, 0.99, 1.0]))
