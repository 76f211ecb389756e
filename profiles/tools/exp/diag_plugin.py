import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
import test_mitsuba_plugin as t
import mitsuba3dopplertof_b200 as dt
import tempfile
d = tempfile.mkdtemp()
for name, defs in [("c4_domino.xml", {"resx": 128, "resy": 96, "spp": 64, "wave": "trapezoidal", "tsm": "antithetic_mirror", "shift": 0.0, "w_g": 150}),
                   ("c4_domino.xml", {"resx": 128, "resy": 96, "spp": 64}),
                   ("c3_rotor.xml", {"resx": 128, "resy": 96, "spp": 64})]:
    out = os.path.join(d, "out.pfm")
    r = t._run([f"-D{k}={v}" for k, v in defs.items()] + ["-o", out, t._plugin_scene(d, name)])
    img = t._read_pfm(out)
    scene = dt.load_file(os.path.join(t.SCENES, name), **defs)
    ref = scene.integrator.render(scene, seed=0)
    ref2 = scene.integrator.render(scene, seed=0)
    diff = np.abs(img - ref).max(axis=2)
    print(name, defs, "scale", np.abs(ref).max(), "max diff", diff.max(), "self diff", np.abs(ref - ref2).max(),
          "n>2e-5*scale", (diff > 2e-5 * np.abs(ref).max()).sum(), "of", diff.size)
    ys, xs = np.nonzero(diff > 2e-5 * np.abs(ref).max())
    print("   pixels:", list(zip(ys.tolist(), xs.tolist()))[:20])
