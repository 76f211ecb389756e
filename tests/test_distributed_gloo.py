"""CPU test of the multi-GPU host logic: 2 ranks over gloo, the oracle injected as the per-rank renderer.
The sharded + all-reduced film must equal the unsharded film for every sharding mode."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_util as gu

WORLD = 2


def _worker(rank, port, mode, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        import mitsuba3dopplertof_b200 as dt
        from mitsuba3dopplertof_b200.distributed import render_distributed
        import oracle_lib
        scene = dt.load_file(os.path.join(gu.SCENES, "c1_example.xml"), resx=24, resy=16, spp=16, pcn=4)

        def oracle_render(flat, p):
            return torch.from_numpy(oracle_lib.OracleScene(flat).render(p, n_threads=2, develop=False))

        film = render_distributed(scene, seed=3, mode=mode, develop=False, render_fn=oracle_render, tile_pixels=5)
        if rank == 0:
            np.save(out, film.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["slots", "tiles", "seeds"])
def test_two_rank_sharding_matches_single(tmp_path, mode):
    import mitsuba3dopplertof_b200 as dt
    import oracle_lib
    out = str(tmp_path / "film.npy")
    port = 29500 + (os.getpid() + hash(mode)) % 2000
    mp.spawn(_worker, args=(port, mode, out), nprocs=WORLD, join=True)
    film = np.load(out)
    scene = dt.load_file(os.path.join(gu.SCENES, "c1_example.xml"), resx=24, resy=16, spp=16, pcn=4)
    flat = scene.flatten()
    if mode == "seeds":
        refs = [oracle_lib.OracleScene(flat).render(scene.integrator.params(scene.sensor.sampler, seed=3 + r), 2, develop=False)
                for r in range(WORLD)]
        ref = sum(refs) / WORLD
    else:
        ref = oracle_lib.OracleScene(flat).render(scene.integrator.params(scene.sensor.sampler, seed=3), 2, develop=False)
    np.testing.assert_allclose(film, ref, rtol=0, atol=2e-6 * max(1.0, np.abs(ref).max()))


def _worker_developed(rank, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        import mitsuba3dopplertof_b200 as dt
        from mitsuba3dopplertof_b200.distributed import render_distributed
        import oracle_lib
        scene = dt.load_file(os.path.join(gu.SCENES, "c1_example.xml"), resx=24, resy=16, spp=16, pcn=4)

        def oracle_render(flat, p):
            return torch.from_numpy(oracle_lib.OracleScene(flat).render(p, n_threads=2, develop=False))

        img = render_distributed(scene, seed=3, mode="seeds", develop=True, render_fn=oracle_render)
        if rank == 0:
            np.save(out, img.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_seeds_image_is_the_mean_of_the_developed_images(tmp_path):
    """The tutorials' multi-seed estimator (doppler_tutorials/src/program_runner.py:11-31) and dtof_render_multi_pass:
    mean over seeds of rgb_r / w_r. With the tent filter the weights of two seeds differ, so this is NOT
    sum(rgb) / sum(w)."""
    import mitsuba3dopplertof_b200 as dt
    import oracle_lib
    out = str(tmp_path / "img.npy")
    mp.spawn(_worker_developed, args=(29500 + (os.getpid() + 77) % 2000, out), nprocs=WORLD, join=True)
    img = np.load(out)
    scene = dt.load_file(os.path.join(gu.SCENES, "c1_example.xml"), resx=24, resy=16, spp=16, pcn=4)
    flat = scene.flatten()
    films = [oracle_lib.OracleScene(flat).render(scene.integrator.params(scene.sensor.sampler, seed=3 + r), 2, develop=False)
             for r in range(WORLD)]
    dev = [f[..., :3] / np.where(f[..., 3:4] == 0, 1, f[..., 3:4]) for f in films]
    ref = sum(dev) / WORLD
    np.testing.assert_allclose(img, ref, rtol=0, atol=2e-6 * max(1.0, np.abs(ref).max()))
    ratio_of_sums = sum(films)[..., :3] / np.where(sum(films)[..., 3:4] == 0, 1, sum(films)[..., 3:4])
    assert np.abs(ratio_of_sums - ref).max() > 1e-5 * np.abs(ref).max()    # the two estimators do differ on this film


def test_shard_params_keep_correlate_groups_together():
    import mitsuba3dopplertof_b200 as dt
    from mitsuba3dopplertof_b200 import _abi
    from mitsuba3dopplertof_b200.distributed import shard_params
    scene = dt.load_file(os.path.join(gu.SCENES, "c1_example.xml"), spp=1024, tcn=2, pcn=4)
    p = scene.integrator.params(scene.sensor.sampler)
    pi = _abi.PassInfo(1024, 1, 256 * 256 * 1024)
    for world in (2, 4, 8):
        covered = np.zeros(1024, int)
        for r in range(world):
            q = shard_params(p, pi, world, r, "slots")
            assert q.shard_block % 4 == 0 and q.shard_count == world and q.shard_index == r
            covered[r * q.shard_block:(r + 1) * q.shard_block] += 1
        assert (covered == 1).all()
    with pytest.raises(ValueError):
        shard_params(p, _abi.PassInfo(12, 1, 12), 8, 0, "slots")
