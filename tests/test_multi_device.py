"""One context over several GPUs behind the C ABI (dtof_create_multi): the path a plugin inside the reference's single
`mitsuba` process uses to reach the whole box. CPU: the symbol surface and argument checks. GPU (needs >= 2 devices, run
with `gpurun --gpus 2` or more): a sharded render equals the one-device render to summation order, for both shardings,
the multi-pass driver and after a keyframe update."""
import ctypes as C
import os

import numpy as np
import pytest

import golden_util as gu
import mitsuba3dopplertof_b200 as dt
from mitsuba3dopplertof_b200 import _abi, runtime


def test_create_multi_rejects_bad_arguments():
    lib = runtime.load_library()
    h = C.c_void_p()
    assert lib.dtof_create_multi(C.byref(h), None, 2) == _abi.ERR_INVALID
    assert lib.dtof_create_multi(C.byref(h), (C.c_int * 2)(0, 0), 2) in (_abi.ERR_INVALID, 2)   # the same device twice
    assert lib.dtof_create_multi(C.byref(h), (C.c_int * 1)(0), 0) == _abi.ERR_INVALID
    assert lib.dtof_device_count(None) == 0


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:   # noqa: BLE001
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("scene_name,kw,mode", [
    ("c2_arealight", dict(resx=96, resy=64, spp=64), 0),                              # 64 spp over D devices: sample slots
    ("c4_domino", dict(resx=64, resy=64, spp=6, tcn=2, pcn=2, wave="trapezoidal", tsm="antithetic_mirror", shift=0.0), 1),
])
def test_sharded_render_equals_single_device(scene_name, kw, mode):
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    devices = list(range(min(n, 8)))
    if mode == 1 and 6 % (2 * len(devices)) == 0:
        devices = devices[:2] if len(devices) != 2 else devices   # keep spp_per_pass indivisible -> tile sharding
    scene = dt.load_file(os.path.join(gu.SCENES, scene_name + ".xml"), **kw)
    params = scene.integrator.params(scene.sensor.sampler, seed=9)
    one = runtime.Context(0)
    flat = one.upload(scene)
    want_img, want = one.render(flat, params, both=True)
    multi = runtime.Context(devices=devices)
    assert multi.device_count() == len(devices)
    flat_m = multi.upload(scene)
    got_img, got = multi.render(flat_m, params, both=True)
    scale = np.abs(want[..., :3]).max()
    assert np.abs(got[..., 3] - want[..., 3]).max() <= 1e-4 * want[..., 3].max()
    assert np.abs(got[..., :3] - want[..., :3]).max() <= 1e-5 * scale
    np.testing.assert_allclose(got_img, want_img, rtol=0, atol=1e-5 * np.abs(want_img).max())
    # a caller-side shard on a multi-device context is refused: the context shards by itself
    p = scene.integrator.params(scene.sensor.sampler, seed=9)
    p.shard_block, p.shard_count, p.shard_index = 64, 2, 0
    with pytest.raises(ValueError):
        multi.render(flat_m, p)
    # multi-pass driver: seeds dealt to the devices, mean of developed images
    base = scene.integrator.params(scene.sensor.sampler, seed=20)
    a = one.render_multi_pass(flat, base, 3)
    b = multi.render_multi_pass(flat_m, base, 3)
    assert np.abs(a - b).max() <= 1e-5 * np.abs(a).max()
    # keyframe update reaches every device
    anim = [(i, flat.instances[i]) for i in range(flat.desc.n_instances) if flat.instances[i].animated]
    if anim:
        i, inst = anim[0]
        moved = _abi.Instance.from_buffer_copy(inst)
        moved.m1[3] += 0.05
        one.update_instances(i, [moved])
        multi.update_instances(i, [moved])
        w2 = one.render(flat, params, develop=False)
        g2 = multi.render(flat_m, params, develop=False)
        assert np.abs(w2[..., :3] - want[..., :3]).max() > 0       # the motion changed the image
        assert np.abs(g2[..., :3] - w2[..., :3]).max() <= 1e-5 * np.abs(w2[..., :3]).max()
    multi.close()
    one.close()


@pytest.mark.gpu
def test_single_device_multi_context_is_the_plain_context():
    scene = dt.load_file(os.path.join(gu.SCENES, "c1_example.xml"), resx=32, resy=32, spp=32)
    params = scene.integrator.params(scene.sensor.sampler, seed=1)
    a, b = runtime.Context(0), runtime.Context(devices=[0])
    fa, fb = a.upload(scene), b.upload(scene)
    ra, rb = a.render(fa, params, develop=False), b.render(fb, params, develop=False)
    assert b.device_count() == 1
    assert np.abs(ra - rb).max() <= 5e-6 * np.abs(ra).max()   # two renders: float atomics arrive in any order
    a.close()
    b.close()
