#!/usr/bin/env python
"""Generates tests/golden/tof_postprocess.npz from the REFERENCE's own post-processing functions
(/root/reference/doppler_tutorials/src/utils/image_utils.py). The module imports matplotlib and cv2 for its plotting
helpers; neither is installed here and neither is touched by the numeric functions, so both are stubbed.
Run in the build container only (the GPU box has no /root/reference); the .npz is the committed artefact."""
import os
import sys
import types

import numpy as np

REF = "/root/reference/doppler_tutorials/src"
for name in ("matplotlib", "matplotlib.pyplot", "cv2"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, REF)
from utils import image_utils as ref   # noqa: E402

rng = np.random.default_rng(20261017)
H, W = 24, 32
out = {}
# rendered images are float32 (H, W, 3); ToF images are their float64 luminance x exposure time
img = (rng.standard_normal((H, W, 3)) * 1e-3).astype(np.float32)
out["img"] = img
out["luminance"] = ref.rgb2luminance(img)
out["tof"] = ref.to_tof_image(img)
out["tof_T2"] = ref.to_tof_image(img, 0.002)
out["tof_0_5"] = ref.to_tof_image_0_5(img)
homos, heteros = [], []
for i in range(3):
    ho = rng.standard_normal((H, W)) * 1e-6
    he = ho * rng.uniform(-1.5, 1.2, (H, W))       # ratios on both sides of the clip range [-1, 0.999]
    ho[rng.random((H, W)) < 0.05] = 0.0            # pixels without signal
    homos.append(ho)
    heteros.append(he)
out["homos"] = np.stack(homos)
out["heteros"] = np.stack(heteros)
out["v_single"] = ref.calc_velocity_from_homo_hetero(homos[0].copy(), heteros[0].copy())
out["v_single_kw"] = ref.calc_velocity_from_homo_hetero(homos[1].copy(), heteros[1].copy(), exposure_time=0.002, w_g=150)
out["v_multi"] = ref.calc_velocity_from_homo_heteros([h.copy() for h in homos], [h.copy() for h in heteros])
out["v_multi_kw"] = ref.calc_velocity_from_homo_heteros([h.copy() for h in homos[:2]], [h.copy() for h in heteros[:2]],
                                                          exposure_time=0.001, w_g=60)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tof_postprocess.npz"), **out)
print({k: v.shape for k, v in out.items()})
