"""Multi-GPU rendering: one process per GPU, shards of the wavefront, one film reduction.

The path shards without any data-path exchange: every wavefront lane is independent and
its three PCG32 streams are a pure function of (seed, lane index). Three sharding modes:

  "slots"  each rank renders a contiguous range of the sample slots of EVERY pixel (spp sharding, C5); blocks are
           multiples of lcm(time_correlate_number, path_correlate_number) so antithetic / correlated groups stay together
  "tiles"  interleaved pixel tiles (C4): block = tile_pixels consecutive pixels
  "seeds"  rank r renders the full wavefront with seed + r (the tutorials' multi-seed averaging,
           doppler_tutorials/src/program_runner.py:11-31); the result is the mean over seeds

All passes of a lane stay on one rank (the RNG state persists across passes, src/render/integrator.cpp:299-308).
The only collective is one all-reduce (sum) of the (H, W, 4) RGBW film per render -- NCCL over NVLink on GPUs; the
host-side logic runs unchanged on gloo for CPU tests with an injected render function.
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import numpy as np

from . import _abi

__all__ = ["shard_params", "render_distributed"]


def shard_params(params: _abi.Params, pass_info, world: int, rank: int, mode: str = "slots",
                 tile_pixels: int = 64) -> _abi.Params:
    """Returns a copy of `params` restricted to rank's share of the wavefront."""
    p = _abi.Params.from_buffer_copy(params)
    if world == 1:
        return p
    spp_pp = int(pass_info.spp_per_pass)
    if mode == "slots":
        group = math.lcm(int(p.time_correlate_number), int(p.path_correlate_number))
        if spp_pp % (world * group) != 0:
            raise ValueError(f"spp_per_pass={spp_pp} cannot be split into {world} shards of whole correlate groups ({group})")
        p.shard_block = spp_pp // world
    elif mode == "tiles":
        p.shard_block = spp_pp * int(tile_pixels)
    elif mode == "seeds":
        p.seed = (int(p.seed) + rank) & 0xFFFFFFFF
        return p
    else:
        raise ValueError(f"unknown sharding mode '{mode}'")
    p.shard_count, p.shard_index = world, rank
    return p


def render_distributed(scene, seed: int = 0, spp: int = 0, mode: str = "slots", develop: bool = True,
                       render_fn: Optional[Callable] = None, tile_pixels: int = 64):
    """Collective render over the default torch.distributed process group. Every rank returns the full image.

    render_fn(flat_scene, params) -> torch.Tensor (H, W, 4) on the rank's device; default: the CUDA library,
    accumulating straight into a device tensor on the current CUDA stream (no host round trip before the reduce)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    integ = scene.integrator
    base = integ.params(scene.sensor.sampler, seed=seed, spp=spp)
    if render_fn is None:
        from .runtime import get_context
        ctx = get_context()
        from .integrator import _uploaded
        flat = _uploaded(ctx, scene)   # uploaded again when the scene changed (Scene.fingerprint)
        pi = ctx.pass_info(base)

        def render_fn(flat_, p_):   # noqa: ANN001
            film_ = torch.zeros((flat_.height, flat_.width, 4), dtype=torch.float32, device=f"cuda:{ctx.device}")
            ctx.render_device(p_, film_.data_ptr(), torch.cuda.current_stream().cuda_stream)
            return film_
    else:
        flat = scene.flatten()
        pi = _pass_info_host(flat, base)
    p = shard_params(base, pi, world, rank, mode, tile_pixels)
    film = render_fn(flat, p)

    def _develop(f):   # HDRFilm::develop
        w = f[..., 3:4]
        return f[..., :3] / torch.where(w == 0, torch.ones_like(w), w)

    if mode == "seeds" and develop:
        # the tutorials' estimator (program_runner.py:11-31) and dtof_render_multi_pass: the MEAN OF THE DEVELOPED
        # images, mean_r(rgb_r / w_r) -- not sum(rgb) / sum(w), which differs as soon as the filter weights of two seeds
        # differ (tent, gaussian and the lobed filters)
        img = _develop(film)
        if world > 1:
            dist.all_reduce(img)
            img /= world
        return img
    if world > 1:
        dist.all_reduce(film)            # sum of RGBW films: slots / tiles are shards of ONE film
        if mode == "seeds":
            film /= world                # develop=False: the mean RGBW film over the seeds (a different estimator than
                                         # the developed mean above; documented, the caller asked for raw films)
    if not develop:
        return film
    return _develop(film)


def _pass_info_host(flat, p):
    """src/render/integrator.cpp:121-134,227-245 (host restatement used when no CUDA context exists)."""
    spp = int(p.sample_count)
    spp_pp, n_passes = spp, 1
    wave = flat.width * flat.height * spp_pp
    if wave > 0xFFFFFFFF:
        spp_pp //= (wave + 0xFFFFFFFF - 1) // 0xFFFFFFFF
        n_passes = spp // spp_pp
        wave = flat.width * flat.height * spp_pp
    if spp % spp_pp:
        raise ValueError("sample_count should be a multiple of samples_per_wavefront!")
    return _abi.PassInfo(spp_pp, n_passes, wave)
