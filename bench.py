#!/usr/bin/env python
"""Benchmark of the hot path (dopplertofpath + correlated sampler) -- one JSON line on stdout.

Metric: Msamples/s = camera samples (pixels x spp, each a full path of <= max_depth bounces incl. NEE) per second
of render(), scene resident on the GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload all|c1|c2|c3|c4|c5] [--impl reference]

* headline     C1 = configs_example/scene.xml verbatim (BASELINE.json configs[0], the "example scene" the north_star target
               is quoted on), exactly K timed steps after W >= 3 warm-up steps. With --workload all (the default) the same
               line carries a `workloads` array with one record per configuration of BASELINE.json:
                 N = 1: C1, C2, C3, C4 (1024^2 @ 4096 spp, 2 passes), C5 (4.2 M triangles, 2048^2 @ 2 x 512 spp), full size;
                 N > 1: C1 and C2 weak (one seed per rank), C4 STRONG by pixel tiles, C5 STRONG by sample slots
               each with value / e2e / roofline / clocks; secondary workloads run min(K, budget / step time) >= 3 timed steps
* step         one render() of the workload (film zeroed, all passes, film all-reduce when N > 1, develop)
* value        whole-job samples / device time (CUDA events per step on the render stream, max over ranks);
               an L2 flush (256 MiB write) runs between steps outside the timed events
* e2e          same metric through the host-buffer C ABI call (`dtof_update_instances` + `dtof_render`):
               H2D of the animated-instance keyframes + parameters, D2H of the RGBW film and the developed image
* roofline     the BINDING bound per workload: `issue` (lane-instructions of the DESIGN.md 4.1 model and, from the committed
               ncu capture, executed thread-instructions, against 148 x 4 x 32 x SM clock) where the traversal data is
               shared-memory resident (C1-C4); `hbm` (algorithmic traversal bytes 64 B x nodes + 48 B x triangles + 112 B x
               instance entries + 16 B film per sample, and the DRAM bytes ncu measured, against the measured HBM copy
               bandwidth) where the BVH is walked from HBM (C5, wavefront pipeline). Per-sample counts come from a stats
               launch of the same walk; both sub-records are always present
* cpu_baseline the reference's own CPU build (oracle/_ref/mitsuba, scalar_rgb + Embree) on all host threads when it is
               present, else the CPU oracle (oracle/, a port of the reference algorithm); bounded sample; headline only
* N > 1        weak scaling: rank r renders the workload with seed r (the tutorials' multi-seed averaging,
               doppler_tutorials/src/program_runner.py:11-31), films are summed with one NCCL all-reduce per step;
               strong scaling: ONE render sharded over the ranks inside the C ABI, same all-reduce
* --impl reference   the reference's own CPU build (oracle/_ref/mitsuba, scalar_rgb + Embree; llvm_rgb cannot load
               libLLVM in this image) on a bounded sample of the headline workload; falls back to the oracle port if the
               binary cannot run
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


HEADLINE = "c1"   # configs_example/scene.xml verbatim: the scene BASELINE.json's target is quoted on


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


def _counters(workload, pipeline):
    """ncu-measured per-sample counters of the shipped library (profiles/r02_counters.json, written from the committed
    ncu captures by profiles/tools/counters.py); None where no capture of this workload / pipeline exists."""
    for fn in ("r02_counters.json", "r01_traffic.json"):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", fn))).get(workload + ("_wavefront" if pipeline == 1 else ""))
            if tj:
                return dict(tj, file="profiles/" + fn)
        except Exception:   # noqa: BLE001
            pass
    return None


def bench_workload(name, args, steps, warmup, scaling, world, rank, local, dist, torch, budget_s=None, cpu_baseline=False):
    """One workload, one JSON-able record (rank 0; other ranks return None). `scaling`: weak | strong-slots | strong-tiles."""
    from mitsuba3dopplertof_b200 import _abi, runtime
    scene, desc = load_workload(name, args.spp)
    ctx = runtime.Context(local)
    t0 = time.perf_counter()
    flat = ctx.upload(scene)
    upload_s = time.perf_counter() - t0
    sampler = scene.sensor.sampler
    H, W = flat.height, flat.width
    spp = sampler.sample_count
    samples_per_step = H * W * spp
    strong = scaling != "weak" and world > 1
    seed = 0 if strong else rank     # weak scaling: one seed per rank
    params = scene.integrator.params(sampler, seed=seed)
    pi = ctx.pass_info(params)
    full_params = params
    if strong:   # this rank's share of the one wavefront (interleaved shards, whole correlate groups / whole tiles)
        from mitsuba3dopplertof_b200.distributed import shard_params
        params = shard_params(params, pi, world, rank, "slots" if scaling == "strong-slots" else "tiles", tile_pixels=64)

    stream = torch.cuda.current_stream()
    film = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    img = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def step():
        film.zero_()
        ctx.render_device(params, film.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_reduce(film)    # sum of the RGBW films over NVLink (NCCL)
        ctx.develop_device(film.data_ptr(), img.data_ptr(), stream.cuda_stream)

    warmup = max(warmup, 3)
    t0 = time.perf_counter()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    est_step_s = (time.perf_counter() - t0) / warmup
    if budget_s:    # secondary workloads: a bounded number of timed steps (>= 3), stated in the record
        steps = int(max(3, min(steps, budget_s / max(est_step_s, 1e-6))))
        if world > 1:
            tt = torch.tensor([steps], dtype=torch.int64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MIN)
            steps = int(tt.item())
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local)
    clocks.start()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
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
    launches = ctx.launch_count() - launches0 + steps   # ours + one film memset per step
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks.stop_flag = True
    clocks.join(timeout=2)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    jobs = 1 if strong else world     # strong scaling: the ranks share ONE render
    value = jobs * samples_per_step * steps / (total_ms * 1e-3) / 1e6

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
    for _ in range(steps):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = jobs * samples_per_step * steps / float(t.item()) / 1e6

    if rank != 0:
        ctx.close()
        del film, img, flush
        torch.cuda.empty_cache()
        return None

    # ---- roofline of the dominant kernel: per-sample counts of the BVH walk from a stats launch of the same walk
    ctx.set_stats(True)
    # the counters are per-sample averages: every K-th pixel (all its sample slots) is plenty, ~16 M lanes
    K = max(1, int(pi.wavefront_size // (1 << 24))) | 1
    ps = scene.integrator.params(sampler, seed=seed)
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
    rate = samples_per_launch / (kms * 1e-3)                           # samples / s of the dominant kernel(s)
    clk = clocks.summary()
    issue_peak = 148 * 4 * 32 * (clk["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)) * 1e6
    cnt = _counters(name, prod_pipeline)
    hbm = {"algorithmic_bytes_per_sample": bytes_per_sample, "achieved": bytes_per_sample * rate / 1e9, "peak": peaks["hbm_gbs"],
           "unit": "GB/s", "frac": bytes_per_sample * rate / 1e9 / peaks["hbm_gbs"],
           "measured_dram_bytes_per_sample": cnt.get("bytes_per_sample") if cnt else None,
           "measured_dram_frac": cnt["bytes_per_sample"] * rate / 1e9 / peaks["hbm_gbs"] if cnt and cnt.get("bytes_per_sample") else None,
           "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})"}
    issue = {"model_lane_instr_per_sample": instr_per_sample, "achieved": instr_per_sample * rate / 1e9, "peak": issue_peak / 1e9,
             "unit": "Glane-instr/s", "frac": instr_per_sample * rate / issue_peak,
             "executed_thread_instr_per_sample": cnt.get("thread_inst_per_sample") if cnt else None,
             "executed_frac": cnt["thread_inst_per_sample"] * rate / issue_peak if cnt and cnt.get("thread_inst_per_sample") else None,
             # warp-instructions ncu counted x the live sample rate / (148 x 4 issue slots x clock): how full the schedulers are
             "warp_issue_slot_frac": cnt["warp_inst_per_sample"] * rate / (issue_peak / 32) if cnt and cnt.get("warp_inst_per_sample") else None,
             "ncu_active_lanes_per_inst": cnt.get("lanes_per_inst") if cnt else None,
             "ncu_issue_active_pct": cnt.get("issue_active_pct") if cnt else None,
             "peak_source": "148 SMs x 4 SMSPs x 32 lanes x SM clock sampled under load"}
    smem_resident = mode_name != "bvh_global"
    # the BINDING bound: instruction issue where the traversal data is shared-memory resident (DRAM traffic ~ 0), HBM where
    # the BVH is walked from HBM / L2. Both sub-records are always present; the top-level fields repeat the binding one.
    top = issue if smem_resident else hbm
    roofline = {
        "bound": "issue" if smem_resident else "hbm", "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"],
        "frac": top["frac"],
        "traffic": cnt["bytes_per_sample"] * samples_per_launch if cnt and cnt.get("bytes_per_sample") else None,
        "traffic_source": f"{cnt['file']}: ncu dram__bytes_read.sum + dram__bytes_write.sum per sample x samples per launch ({cnt.get('source')})" if cnt else None,
        "kernel": ("wavefront pipeline: wf_generate + per bounce wf_trace<closest> / wf_shade / wf_trace<any> + wf_splat, "
                   "several batches in flight (kernel_ms spans the whole pipeline of one render)") if prod_pipeline == 1
        else "render_kernel",
        "pipeline": "wavefront" if prod_pipeline == 1 else "fused",
        "kernel_ms": kms, "samples_per_launch": samples_per_launch,
        "per_sample": {"rays_closest": st.rays_closest / n, "rays_shadow": st.rays_shadow / n, "nodes": st.nodes_visited / n,
                       "tris": st.tris_tested / n, "inst": st.inst_visits / n},
        "traversal_mode": mode_name,
        "issue": issue, "hbm": hbm,
        "note": ("per-sample counts are those of the BVH walk (closest + shadow rays). " + (
            "The traversal data of this workload is staged in shared memory: DRAM traffic is ~0 and the `hbm` sub-record is "
            "only the formal byte rate; the kernel is bound by instruction issue (`frac` = algorithmic lane-instructions of "
            "the DESIGN.md 4.1 model / peak issue rate, `issue.executed_frac` = thread-instructions ncu counted / peak)"
            if smem_resident else
            "Nodes / triangles are read through L1 / L2 from HBM: `frac` = algorithmic traversal bytes / measured HBM peak -- a "
            "byte RATE most of which the caches serve (L1 80 %, L2 65 % hit rate); `hbm.measured_dram_frac` = DRAM bytes ncu "
            "counted (incl. the pipeline's own queues and per-lane state) / peak, `issue.warp_issue_slot_frac` = how full the "
            "warp schedulers are. Neither is near 1: the pipeline is bound by warp-instructions issued for few lanes in the leaf "
            "phases of the walk (profiles/r02_tuning.md)")),
    }

    cpu = None
    if cpu_baseline:
        # the reference's own CPU build when oracle/_ref holds one (it is the faster of the two), else the oracle port
        try:
            cores = os.cpu_count() or 1
            fn, wl_params, _ = WORKLOADS[name]
            wp = dict({"resx": 256, "resy": 256, "spp": 1024}, **wl_params)
            px = int(wp["resx"]) * int(wp["resy"])
            probe = reference_binary_throughput(name, 4, cores)
            ref_spp = int(max(4, min(int(args.spp or wp["spp"]), (probe * 1e6 * 12.0 / px) // 4 * 4)))
            v = reference_binary_throughput(name, ref_spp, cores)
            cpu = {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "reference",
                   "sample": f"{wp['resx']}x{wp['resy']} @ {ref_spp} spp of the same scene, reference scalar_rgb+Embree binary "
                             f"(oracle/_ref/mitsuba wrapped in `moment`, -t {cores}; llvm_rgb cannot load libLLVM in this image)"}
        except Exception as e:   # noqa: BLE001
            sys.stderr.write(f"[bench] reference binary unusable here ({e}); timing the oracle port instead\n")
            v, cores, sample = cpu_oracle_throughput(scene, seed)
            cpu = {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample}

    rec = {
        "workload": name, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": total_ms / steps, "scaling": "strong" if strong else "weak",
        "sharding": scaling[7:] if strong else ("seeds" if world > 1 else None),
        "config": {"workload": desc, "width": W, "height": H, "spp": spp, "spp_per_pass": pi.spp_per_pass,
                   "n_passes": pi.n_passes, "triangles": flat.n_triangles, "instances": flat.desc.n_instances,
                   "seed": "0, one render sharded by " + scaling[7:] if strong else "rank index (multi-seed averaging)",
                   "l2_flush": "256 MiB write between steps, outside the timed events", "scene_upload_s": upload_s},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    ctx.close()
    del film, img, flush
    torch.cuda.empty_cache()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="all", choices=["all"] + sorted(WORKLOADS),
                    help="all (default): the headline line is C1, the example scene the north_star target is quoted on, and a "
                         "`workloads` array carries every configuration of BASELINE.json (N = 1: C1..C5; N > 1: C1 and C2 weak, "
                         "C4 strong by tiles, C5 strong by sample slots); cN: that workload alone")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--spp", type=int, default=0, help="override the workload's spp (profiling under ncu only)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong-slots", "strong-tiles"],
                    help="with --workload cN and N > 1: weak = one seed per rank; strong-* = ONE render of the workload "
                         "sharded over the ranks by sample slots / pixel tiles (SURVEY.md 8e), films summed by one all-reduce")
    ap.add_argument("--secondary-budget", type=float, default=10.0,
                    help="seconds of timed steps per secondary workload of --workload all (>= 3 steps each)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.workload == "all":
            args.workload = HEADLINE
        return run_reference_arm(args)

    # stdout carries exactly ONE line, the JSON record: anything libraries print to fd 1 meanwhile (NCCL's version
    # banner under NCCL_DEBUG=VERSION, for one) is sent to stderr until the record is written
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    common = dict(world=world, rank=rank, local=local, dist=dist, torch=torch)
    if args.workload != "all":
        plan = [(args.workload, args.scaling, None)]
    elif world == 1:
        plan = [(HEADLINE, "weak", None)] + [(w, "weak", args.secondary_budget) for w in sorted(WORKLOADS) if w != HEADLINE]
    else:
        plan = [(HEADLINE, "weak", None), ("c2", "weak", args.secondary_budget),
                ("c4", "strong-tiles", args.secondary_budget), ("c5", "strong-slots", args.secondary_budget)]
    records = []
    for i, (name, scaling, budget) in enumerate(plan):
        rec = bench_workload(name, args, args.steps, args.warmup, scaling, budget_s=budget,
                             cpu_baseline=(i == 0 and world == 1 and not args.no_cpu_baseline), **common)
        records.append(rec)
        if rank == 0:
            sys.stderr.write(f"[bench] {name} {scaling} N={world}: {rec['value']:.1f} Msamples/s, e2e {rec['e2e']['value']:.1f} "
                             f"({rec['steps']} steps, {rec['ms_per_step']:.2f} ms)\n")

    if rank == 0:
        h = records[0]
        line = {
            "metric": "Msamples/s", "value": h["value"], "unit": "Msamples/s", "n_gpus": world, "steps": h["steps"],
            "warmup": h["warmup"], "ms_per_step": h["ms_per_step"], "higher_is_better": True, "scaling": h["scaling"],
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": h["config"], "clocks": h["clocks"], "e2e": h["e2e"], "gpu_launches": h["gpu_launches"],
            "roofline": h["roofline"], "cpu_baseline": h["cpu_baseline"],
        }
        if len(records) > 1:
            line["gpu_launches"] = int(sum(r["gpu_launches"] for r in records))
            line["workloads"] = [{k: v for k, v in r.items() if k != "cpu_baseline"} for r in records]
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
