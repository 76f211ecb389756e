#!/usr/bin/env python
"""Benchmark of the hot path (dopplertofpath + correlated sampler) -- one JSON line on stdout.

Metric: Msamples/s = camera samples (pixels x spp, each a full path of <= max_depth bounces incl. NEE) per second
of render(), scene resident on the GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5] [--impl reference]

* step         one render() of the workload (film zeroed, all passes, film all-reduce when N > 1, develop)
* value        whole-job samples / device time (CUDA events per step on the render stream, max over ranks);
               an L2 flush (256 MiB write) runs between steps outside the timed events
* e2e          same metric through the host-buffer C ABI call (`dtof_update_instances` + `dtof_render`):
               H2D of the animated-instance keyframes + parameters, D2H of the RGBW film and the developed image
* roofline     traversal bytes: (64 B x nodes + 48 B x triangles + 112 B x instance entries) per sample, counted by a
               separate stats launch of the fused kernel (same walk), + 16 B film; / measured HBM copy bandwidth.
               Scenes whose BVH is walked from HBM (c5) render through the wavefront pipeline (csrc/dtof_wavefront.cuh)
* cpu_baseline the reference's own CPU build (oracle/_ref/mitsuba, scalar_rgb + Embree) on all host threads when it is
               present, else the CPU oracle (oracle/, a port of the reference algorithm); bounded sample
* N > 1        weak scaling: rank r renders the workload with seed r (the tutorials' multi-seed averaging,
               doppler_tutorials/src/program_runner.py:11-31), films are summed with one NCCL all-reduce per step
* --impl reference   the reference's own CPU build (oracle/_ref/mitsuba, scalar_rgb + Embree; llvm_rgb cannot load
               libLLVM in this image) on a bounded sample; falls back to the oracle port if the binary cannot run
"""
import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (scene file, xml params, description)  -- BASELINE.json configs[0..3]
    "c1": ("c1_example.xml", {}, "C1 configs_example/scene.xml verbatim: 256x256 @ 1024 spp, max_depth 4, heterodyne, antithetic"),
    "c2": ("c2_arealight.xml", {"resx": 512, "resy": 512},
           "C2 Cornell box + translating box, area light, heterodyne hf=1, antithetic, 512x512 @ 1024 spp"),
    "c3": ("c3_rotor.xml", {"resx": 512, "resy": 512, "wave": "rectangular", "tsm": "stratified", "shift": 0.0, "pcn": 4},
           "C3 rotating cube, rectangular waveform, stratified per-interval, pcn=4, 512x512 @ 1024 spp"),
    "c4": ("c4_domino.xml", {"resx": 1024, "resy": 1024, "spp": 4096, "wave": "trapezoidal", "tsm": "antithetic_mirror",
                              "shift": 0.0, "w_g": 150},
           "C4 domino (32 animated boxes), trapezoidal, antithetic_mirror, 1024x1024 @ 4096 spp (2 passes x 2048)"),
    # C5: one of the 16 renders (seed = rank) the reference needs for 2048^2 @ 16k spp: 1024 spp = 2 passes x 512
    "c5": ("c5_slabroom.xml", {"resx": 2048, "resy": 2048, "spp": 1024},
           "C5 slab room + 4.2 M-triangle displaced sphere as one animated instance, 2048x2048 @ 1024 spp per render "
           "(2 passes x 512; 16 renders with seed 0..15 make the 16k-spp image)"),
}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:   # noqa: BLE001
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(n)
            except Exception:   # noqa: BLE001
                pass
            time.sleep(0.2)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def load_workload(name, spp=0):
    import mitsuba3dopplertof_b200 as dt
    fn, params, desc = WORKLOADS[name]
    params = dict(params)
    if spp:
        params["spp"] = spp
        desc += f" [REDUCED to {spp} spp: profiling run, not a bench value]"
    scene = dt.load_file(os.path.join(ROOT, "tests", "scenes", fn), **params)
    if name == "c5":
        from mitsuba3dopplertof_b200 import procedural
        # 12 n^2 triangles: n = 592 -> 4.2 M (the workload); --mesh-n scales it for the mid-size experiments of the tuning log
        n = int(os.environ.get("DTOF_BENCH_MESH_N", "592"))
        if n != 592:
            desc += f" [mesh n = {n}: {12 * n * n} triangles, experiment, not a bench value]"
        scene = procedural.large_scene(scene, n=n, seed=1234)
    return scene, desc


def cpu_oracle_throughput(scene, seed, target_seconds=12.0):
    """Times the CPU oracle (port of the reference algorithm) on a bounded sample of the workload: same scene and
    resolution, reduced spp."""
    import oracle_lib
    flat = scene.flatten()
    osc = oracle_lib.OracleScene(flat)
    cores = os.cpu_count() or 1
    tcn = scene.sensor.sampler.time_correlate_number
    pcn = scene.sensor.sampler.path_correlate_number
    group = int(np.lcm(tcn, pcn))
    px = flat.width * flat.height
    spp = group
    p = scene.integrator.params(scene.sensor.sampler, seed=seed, spp=spp)
    t0 = time.perf_counter()
    osc.render(p, cores)
    dt0 = time.perf_counter() - t0
    rate = px * spp / dt0
    spp2 = int(max(group, min(scene.sensor.sampler.sample_count, (rate * target_seconds / px) // group * group)))
    p = scene.integrator.params(scene.sensor.sampler, seed=seed, spp=spp2)
    t0 = time.perf_counter()
    osc.render(p, cores)
    dt1 = time.perf_counter() - t0
    cpu_oracle_throughput.last_seconds = dt1
    return px * spp2 / dt1 / 1e6, cores, f"{flat.width}x{flat.height} @ {spp2} spp of the same scene ({dt1:.1f} s)"


def reference_binary_throughput(workload, spp, threads):
    """Runs the reference's own build (scalar_rgb + Embree) wrapped in `moment` and parses
    'Rendering finished. (took ...)'. Returns Msamples/s or raises."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    exe = os.path.join(ref, "mitsuba")
    if not os.path.exists(exe):
        raise RuntimeError("oracle/_ref/mitsuba not present")
    fn, params, _ = WORKLOADS[workload]
    xml = open(os.path.join(ROOT, "tests", "scenes", fn)).read()
    m = re.search(r"<integrator type=\"dopplertofpath\">.*?</integrator>", xml, flags=re.S)
    inner = m.group(0)
    wrapper = ('<integrator type="moment">\n<boolean name="is_doppler_integrator" value="true"/>\n'
               '<string name="time_sampling_method" value="$tsm"/>\n<float name="antithetic_shift" value="$shift"/>\n'
               '<integer name="path_correlation_depth" value="$pcd"/>\n' + inner + "\n</integrator>")
    xml = xml[:m.start()] + wrapper + xml[m.end():]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "scene_moment.xml")
        open(path, "w").write(xml)
        for f in os.listdir(os.path.join(ROOT, "tests", "scenes")):
            if f.endswith((".ply", ".serialized")):
                os.symlink(os.path.join(ROOT, "tests", "scenes", f), os.path.join(d, f))
        cmd = [exe, "-m", "scalar_rgb", "-t", str(threads), "-o", os.path.join(d, "out.exr"), f"-Dspp={spp}"]
        for k, v in params.items():
            if k != "spp":
                cmd.append(f"-D{k}={v}")
        cmd.append(path)
        env = dict(os.environ, LD_LIBRARY_PATH=ref + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
        r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
        out = r.stdout + r.stderr
        mm = re.search(r"Rendering finished\. \(took ([0-9.]+)(ms|s|m)", out)
        if r.returncode != 0 or not mm:
            raise RuntimeError(f"reference binary failed (rc={r.returncode}): {out[-300:]}")
        secs = float(mm.group(1)) * {"ms": 1e-3, "s": 1.0, "m": 60.0}[mm.group(2)]
    p = dict({"resx": 256, "resy": 256}, **params)
    return int(p["resx"]) * int(p["resy"]) * spp / secs / 1e6


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene, desc = load_workload(args.workload, args.spp)
    cores = os.cpu_count() or 1
    fn, params, _ = WORKLOADS[args.workload]
    p = dict({"resx": 256, "resy": 256, "spp": 1024}, **params)
    px = int(p["resx"]) * int(p["resy"])
    kind, vals, secs = "reference", [], []
    spp = 8
    try:
        probe = reference_binary_throughput(args.workload, 4, cores)
        # size each step for ~10 s of CPU work
        spp = int(max(4, min(int(p["spp"]), (probe * 1e6 * 10.0 / px) // 4 * 4)))
        for i in range(args.warmup + args.steps):
            v = reference_binary_throughput(args.workload, spp, cores)
            if i >= args.warmup:
                vals.append(v)
        sample = f"{p['resx']}x{p['resy']} @ {spp} spp per step, reference scalar_rgb+Embree binary wrapped in `moment`, -t {cores}"
    except Exception as e:   # noqa: BLE001
        sys.stderr.write(f"[bench] reference binary unusable here ({e}); timing the oracle port instead\n")
        kind = "port"
        for i in range(min(args.warmup, 1) + args.steps):
            v, cores, sample = cpu_oracle_throughput(scene, 0, target_seconds=10.0)
            if i >= min(args.warmup, 1):
                vals.append(v)
                secs.append(cpu_oracle_throughput.last_seconds)
    value = float(np.mean(vals))
    ms = px * spp / (value * 1e6) * 1e3 if kind == "reference" else float(np.mean(secs)) * 1e3
    note = ("reference binary (scalar_rgb + Embree; llvm_rgb cannot load libLLVM in this image)" if kind == "reference" else
            "oracle port: multithreaded C++ restatement of the reference algorithm (JIT stream semantics, own BVH); the "
            "reference itself only builds through its CMake tree, which is not available on the bench box (DESIGN.md 2)")
    print(json.dumps({
        "impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "note": note, "step": "one bounded sample of the workload (see cpu_baseline.sample)"},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--spp", type=int, default=0, help="override the workload's spp (profiling under ncu only)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong-slots", "strong-tiles"],
                    help="N > 1: weak = one seed per rank (default, the contract's line); strong-* = ONE render of the workload "
                         "sharded over the ranks by sample slots / pixel tiles (SURVEY.md 8e), films summed by one all-reduce")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    # stdout carries exactly ONE line, the JSON record: anything libraries print to fd 1 meanwhile (NCCL's version
    # banner under NCCL_DEBUG=VERSION, for one) is sent to stderr until the record is written
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from mitsuba3dopplertof_b200 import _abi, runtime

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    scene, desc = load_workload(args.workload, args.spp)
    ctx = runtime.Context(local)
    t0 = time.perf_counter()
    flat = ctx.upload(scene)
    upload_s = time.perf_counter() - t0
    sampler = scene.sensor.sampler
    H, W = flat.height, flat.width
    spp = sampler.sample_count
    samples_per_step = H * W * spp
    strong = args.scaling != "weak" and world > 1
    seed = 0 if strong else rank     # weak scaling: one seed per rank
    params = scene.integrator.params(sampler, seed=seed)
    pi = ctx.pass_info(params)
    full_params = params
    if strong:   # this rank's share of the one wavefront (interleaved shards, whole correlate groups / whole tiles)
        from mitsuba3dopplertof_b200.distributed import shard_params
        params = shard_params(params, pi, world, rank, "slots" if args.scaling == "strong-slots" else "tiles", tile_pixels=64)

    stream = torch.cuda.current_stream()
    film = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    img = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    launches_before = ctx.launch_count()

    def step():
        film.zero_()
        ctx.render_device(params, film.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_reduce(film)    # sum of the RGBW films over NVLink (NCCL)
        ctx.develop_device(film.data_ptr(), img.data_ptr(), stream.cuda_stream)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local)
    clocks.start()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    launches0 = ctx.launch_count()
    for s0, s1 in ev:
        flush.fill_(1)          # L2 flush, outside the timed events
        s0.record(stream)
        step()
        s1.record(stream)
        s1.synchronize()
        kernel_ms.append(ctx.last_kernel_ms())
    torch.cuda.synchronize()
    prod_mode = ctx.last_traversal_mode()
    prod_pipeline = ctx.last_pipeline()      # 0 = fused kernel, 1 = wavefront pipeline (HBM-resident scenes)
    if world > 1:
        dist.barrier()
    launches = ctx.launch_count() - launches0 + args.steps   # ours + one film memset per step
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks.stop_flag = True
    clocks.join(timeout=2)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    jobs = 1 if strong else world     # strong scaling: the ranks share ONE render
    value = jobs * samples_per_step * args.steps / (total_ms * 1e-3) / 1e6

    # ---- e2e through the host-buffer C ABI (what a plugin calls)
    anim = [(i, flat.instances[i]) for i in range(flat.desc.n_instances) if flat.instances[i].animated]
    h2d = len(anim) * C.sizeof(_abi.Instance) + C.sizeof(_abi.Params)
    d2h = H * W * (4 + 3) * 4
    def e2e_step():
        for i, inst in anim:     # per-frame keyframe upload, as an animation loop would do
            ctx.update_instances(i, [inst])
        return ctx.render(flat, params, both=True)
    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_img, e2e_rgbw = e2e_step()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = jobs * samples_per_step * args.steps / float(t.item()) / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        return

    # ---- roofline of the dominant kernel (render_kernel): algorithmic traversal bytes / kernel time
    ctx.set_stats(True)
    # the counters are per-sample averages: every K-th pixel (all its sample slots) is plenty, ~16 M lanes
    K = max(1, int(pi.wavefront_size // (1 << 24))) | 1
    ps = scene.integrator.params(sampler, seed=seed)
    params = full_params   # the roofline counters describe the whole workload
    if K > 1:
        ps.shard_block, ps.shard_count, ps.shard_index = pi.spp_per_pass, K, 0
    film.zero_()
    ctx.render_device(ps, film.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    st = ctx.stats()
    ctx.set_stats(False)
    n = max(st.samples, 1)
    bytes_per_sample = (64 * st.nodes_visited + 48 * st.tris_tested + 112 * st.inst_visits) / n + 16
    instr_per_sample = (40 * st.nodes_visited + 45 * st.tris_tested + 110 * st.inst_visits) / n + \
        420 * (st.rays_closest / n) + 60
    kms = float(np.mean(kernel_ms))
    mode_name = {0: "bvh_global", 1: "bvh_smem", 2: "flat_smem"}.get(prod_mode, str(prod_mode))
    peaks, peak_src = measured_peaks()
    samples_per_launch = samples_per_step / (world if strong else 1)   # rank 0's share of the render under strong scaling
    achieved = bytes_per_sample * samples_per_launch / (kms * 1e-3) / 1e9
    clk = clocks.summary()
    issue_peak = 148 * 4 * 32 * (clk["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)) * 1e6
    traffic, ncu_issue = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(
            args.workload + ("_wavefront" if prod_pipeline == 1 else ""))
        if tj:
            traffic = tj["bytes_per_sample"] * samples_per_step
            ncu_issue = tj.get("issue_active_pct")
    except Exception:   # noqa: BLE001
        pass
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
        "traffic": traffic, "traffic_source": "profiles/r01_traffic.json (ncu dram bytes per sample x samples per launch)" if traffic else None,
        "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})",
        "kernel": ("wavefront pipeline: wf_generate + per bounce wf_trace<closest> / wf_shade / wf_trace<any> + wf_splat, "
                   "several batches in flight (kernel_ms spans the whole pipeline of one render)") if prod_pipeline == 1
        else "render_kernel",
        "pipeline": "wavefront" if prod_pipeline == 1 else "fused",
        "kernel_ms": kms, "bytes_per_sample": bytes_per_sample,
        "per_sample": {"rays_closest": st.rays_closest / n, "rays_shadow": st.rays_shadow / n, "nodes": st.nodes_visited / n,
                       "tris": st.tris_tested / n, "inst": st.inst_visits / n},
        "traversal_mode": mode_name,
        "note": ("counts are those of the BVH walk (closest + shadow rays); " + (
            "the traversal data of this workload is staged in shared memory, so the byte rate is served by SMEM, not HBM"
            if mode_name != "bvh_global" else "nodes/triangles are read through L1/L2 from HBM") + (
            "; the wavefront pipeline additionally moves its ray / hit queues and per-lane path state through HBM, "
            "which is part of `traffic`, not of the algorithmic bytes" if prod_pipeline == 1 else "")),
        "issue": {"instr_per_sample_model": instr_per_sample,
                  "achieved_lane_instr_per_s": instr_per_sample * samples_per_launch / (kms * 1e-3),
                  "peak_lane_instr_per_s": issue_peak,
                  "frac": instr_per_sample * samples_per_launch / (kms * 1e-3) / issue_peak,
                  "ncu_issue_active_pct": ncu_issue},
    }

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        # the reference's own CPU build when oracle/_ref holds one (it is the faster of the two), else the oracle port
        try:
            cores = os.cpu_count() or 1
            fn, wl_params, _ = WORKLOADS[args.workload]
            wp = dict({"resx": 256, "resy": 256, "spp": 1024}, **wl_params)
            px = int(wp["resx"]) * int(wp["resy"])
            probe = reference_binary_throughput(args.workload, 4, cores)
            ref_spp = int(max(4, min(int(args.spp or wp["spp"]), (probe * 1e6 * 12.0 / px) // 4 * 4)))
            v = reference_binary_throughput(args.workload, ref_spp, cores)
            cpu = {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "reference",
                   "sample": f"{wp['resx']}x{wp['resy']} @ {ref_spp} spp of the same scene, reference scalar_rgb+Embree binary "
                             f"(oracle/_ref/mitsuba wrapped in `moment`, -t {cores}; llvm_rgb cannot load libLLVM in this image)"}
        except Exception as e:   # noqa: BLE001
            sys.stderr.write(f"[bench] reference binary unusable here ({e}); timing the oracle port instead\n")
            v, cores, sample = cpu_oracle_throughput(scene, seed)
            cpu = {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample}

    sys.stdout.flush()
    os.dup2(stdout_fd, 1)
    print(json.dumps({
        "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "width": W, "height": H, "spp": spp, "spp_per_pass": pi.spp_per_pass,
                   "n_passes": pi.n_passes, "triangles": flat.n_triangles, "instances": flat.desc.n_instances,
                   "seed": "0, one render sharded by " + args.scaling[7:] if strong else "rank index (multi-seed averaging)", "l2_flush": "256 MiB write between steps, outside the timed events",
                   "scene_upload_s": upload_s},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }), flush=True)
    os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
