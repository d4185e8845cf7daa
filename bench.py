#!/usr/bin/env python
"""bench.py -- headline measurement of the per-pixel tracing loop on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c1..c5] [--all-configs]

A "step" is one frame of the workload.  Default workload = BASELINE.json configs[1]:
scenes/spheres2.scn 1920x1080 --jsample 5 --shadow (the scene snapshot in tests/golden/scenes, produced by the
reference's own parser).  Metric: Mrays/s, rays = closest-hit rays (shade() invocations with depth > 0) + distinct
shadow rays, counted on the device in an untimed pass with the same seed.

  value   : whole-job Mrays/s, scene resident in HBM, frame left in HBM (device time, CUDA events on the library's
            stream, per step, L2 flushed between steps outside the events; max over ranks).
  e2e     : the same metric through the reference-facing call with HOST buffers: skr_scene_upload (H2D) + skr_render
            into pinned host memory (D2H) every step, wall clock around the calls.
  N > 1   : the frame is split into interleaved 32x32 tiles over the ranks (one process per GPU, torchrun); each rank's
            render kernel stores its finished pixels straight into every rank's frame over NVLink (torch symmetric
            memory, skr_render_peers_device) and one symmetric-memory barrier ends the frame; where peer mapping is
            unavailable (or SKR_BENCH_NO_P2P=1): ONE all-gather (NCCL) of the RGB8 tiles + de-interleave kernel.
            Total work is fixed -> "strong".
  roofline: FP32 CUDA-core pipe (this path has no dense contraction; tensor cores unused; HBM traffic is the
            framebuffer only).  peak = FMA microbenchmark measured live in this run (MEASURED_PEAKS.json has no FP32).
            `achieved` = the reference algorithm's arithmetic / kernel time; `executed` = what the kernels really ran
            (bundle culling proves most sphere tests of a jittered pixel unnecessary and skips them).
  --impl reference : the reference's own CPU code (oracle/_ref/libskr_ref.so, compiled from the reference's sources;
            else the C port) with all host threads on a bounded row window of the same frame.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden", "scenes")

WORKLOADS = {
    # BASELINE.json configs[0..4]
    "c1": ("spheres1", dict(width=1920, height=1080, max_depth=1), "scenes/spheres1.scn 1920x1080 --depth 1 (no shadows)"),
    "c2": ("spheres2", dict(width=1920, height=1080, grid_size=5, use_shadows=True), "scenes/spheres2.scn 1920x1080 --jsample 5 --shadow"),
    "c3": ("spheres2", dict(width=1920, height=1080, max_depth=4, monte_carlo=True, num_path_traces=16),
           "scenes/spheres2.scn 1920x1080 --gillum 16 --depth 4"),
    "c4": ("dragon", dict(width=1920, height=1080, use_shadows=True), "scenes/dragon.scn 1920x1080 --shadow"),
    "c5": ("bear", dict(width=3840, height=2160, monte_carlo=True, num_path_traces=64, grid_size=4, use_shadows=True),
           "scenes/bear.scn 3840x2160 --gillum 64 --jsample 4 --shadow"),
}
SEED = 1
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed ncu capture
# (bench.py cannot run ncu on itself): profiles/r01_c2_primary_ncu_raw.txt
NCU_TRAFFIC_BYTES = {"c2": (583680 + 129536, "profiles/r01_c2_primary_ncu_raw.txt"), "c4": (613376 + 1024, "profiles/r01_c4_primary_ncu_raw.txt")}


def algorithmic_flops(st, primary_samples):
    """SURVEY 8(d) per-unit figures x the device counters of one frame (FMA = 2, everything else = 1)."""
    return (25.0 * st["sphere_tests"] + 8.0 * st["sphere_tests_pos"] + 18.0 * st["sphere_hits"]  # F_ch(S) = 25 S + 8 h + 18 [hit]
            + 15.0 * st["shadow_rays"]                                                           # shadow ray set-up (tests counted above at 25/33)
            + 33.0 * st["sphere_hits"] + 89.0 * st["light_evals"]                                # direct shading 33 + 89 L
            + 24.0 * st["bvh_node_visits"] + 35.0 * st["tri_tests"]                              # line-slab test / triangle test
            + 20.0 * primary_samples)                                                            # ray generation


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (profiling recipe's clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                                          str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return None
        sm = [float(r[1]) for r in rows]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": sorted(reasons)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores
# ------------------------------------------------------------------------------------------------

def cpu_reference_run(workload, budget_s, steps, warmup, force_port=False):
    """Times `steps` (+ `warmup`) bounded samples of the workload on the host cores.  Each sample is a window of rows
    of the same frame, sized so that one sample takes about budget_s.  -> dict(value Mrays/s, kind, cores, sample,
    ms_per_step)."""
    from oracle import oracle_lib as O

    scene_name, kw, desc = WORKLOADS[workload]
    sc = O.Scene.load(os.path.join(GOLD, scene_name + ".npz"))
    opt = O.Options(**kw)
    port = O.Port()
    use_ref = O.ref_available() and not force_port
    ref = O.Ref() if use_ref else None
    cores = ref.max_threads() if use_ref else port.max_threads()
    h = opt.height

    def run(y0, y1):
        if use_ref:
            return ref.render(sc, opt, seed=SEED, threads=cores, y0=y0, y1=y1)[2]
        return port.render(sc, opt, rng_mode=O.RNG_PHILOX, seed=SEED, threads=cores, y0=y0, y1=y1, want_rgb8=False)[3]

    mid = h // 2
    probe_rows = max(1, min(4, h))
    t = run(mid, mid + probe_rows)
    rows = int(max(1, min(h, probe_rows * budget_s / max(t, 1e-4))))
    y0 = max(0, mid - rows // 2)
    y1 = min(h, y0 + rows)
    # rays in the window: counted by the port on the same window (same algorithm; keyed RNG)
    st = port.render(sc, opt, rng_mode=O.RNG_PHILOX, seed=SEED, y0=y0, y1=y1, want_rgb8=False)[2]
    rays = st["closest_hit_rays"] + st["shadow_rays"]
    for _ in range(warmup):
        run(y0, y1)
    times = [run(y0, y1) for _ in range(steps)]
    tot = sum(times)
    return {"value": rays * steps / tot / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "reference" if use_ref else "port",
            "sample": f"rows [{y0},{y1}) of the {opt.width}x{opt.height} frame of {desc}: {rays} rays per sample, {steps} timed samples, "
                      f"render loop only (no parse, no PPM write); " + ("reference's own src/ compiled -O2 -fopenmp (oracle/_ref), OMP threads share rand()"
                                                                       if use_ref else "C port oracle/skr_oracle.c -O2 -fopenmp"),
            "ms_per_step": tot / steps * 1e3, "rays_per_sample": rays}


def main_reference(args, rank):
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    budget = max(0.5, min(6.0, 120.0 / (steps + warmup)))
    r = cpu_reference_run(args.workload, budget, steps, warmup)
    scene_name, kw, desc = WORKLOADS[args.workload]
    line = {"impl": "reference", "metric": "Mrays/s", "value": r["value"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "scene": scene_name, **kw, "seed": SEED, "note": "each step is a bounded row window of the frame"},
            "cpu_baseline": {"value": r["value"], "unit": "Mrays/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------

def measure_gpu(S, torch, dist, r, workload, steps, warmup, rank, world, flush_buf, want_e2e=True):
    import numpy as np

    scene_name, kw, desc = WORKLOADS[workload]
    scene = S.Scene.load(os.path.join(GOLD, scene_name + ".npz"))
    r.upload(scene)
    base = S.Options(seed=SEED, rank=rank, world=world, **kw)
    dev = torch.device("cuda", torch.cuda.current_device())
    ext = torch.cuda.ExternalStream(r.stream(), device=dev)

    # untimed counting pass (same seed -> same rays)
    import dataclasses
    cst = r.render_device(dataclasses.replace(base, collect_stats=True), 0, 0).as_dict() if world == 1 else None
    if world > 1:
        tiles = torch.empty(r.tiles_bytes(base), dtype=torch.uint8, device=dev)
        cst = r.render_tiles_device(dataclasses.replace(base, collect_stats=True), tiles.data_ptr()).as_dict()
        keys = ["closest_hit_rays", "shadow_rays", "sphere_tests", "sphere_tests_pos", "tri_tests", "bvh_node_visits", "sphere_hits", "light_evals", "sphere_tests_executed"]
        cst_local = dict(cst)
        t = torch.tensor([float(cst[k]) for k in keys], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        for k, v in zip(keys, t.tolist()):
            cst[k] = int(v)
        gathered = torch.empty(tiles.numel() * world, dtype=torch.uint8, device=dev)
        from skele_raytracer_b200.distributed import PeerFrames
        peer_frames = None if os.environ.get("SKR_BENCH_NO_P2P") == "1" else PeerFrames.create(base.height, base.width, dev)
        ok = torch.tensor([1 if peer_frames is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # every rank takes the same path
        if int(ok.item()) == 0:
            peer_frames = None
    else:
        cst_local = dict(cst)
        peer_frames = None
    frame = torch.empty((base.height, base.width, 3), dtype=torch.uint8, device=dev)
    rays = cst["closest_hit_rays"] + cst["shadow_rays"]

    tree = bool(kw.get("monte_carlo") or kw.get("fresnel"))  # wavefront frames are scheduled with host read-backs

    def step(want_stats=False):
        """One frame.  Single-kernel frames are enqueued asynchronously (stats=NULL): consecutive steps, the all-gather
        and the de-interleave queue up on the library's stream without a host round trip in between."""
        ws = want_stats or tree
        if world == 1:
            return r.render_device(base, frame.data_ptr(), 0, want_stats=ws)
        if peer_frames is not None:
            # collective-free: P2P stores of the finished pixels into every rank's frame + symmetric-memory barrier
            step.frame, st = peer_frames.render(r, base, rank, world, want_stats=ws)
            return st
        st = r.render_tiles_device(base, tiles.data_ptr(), want_stats=ws)
        dist.all_gather_into_tensor(gathered, tiles)
        r.deinterleave_device(base, gathered.data_ptr(), frame.data_ptr())
        return st

    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    launches = 0
    kernel_ms = {"primary": 0.0, "bounce": 0.0, "resolve": 0.0}
    with torch.cuda.stream(ext):
        for _ in range(warmup):
            flush_buf.zero_()  # (also pays torch's lazy load of its fill kernel before the timed region)
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_begin = time.time()
        for i in range(steps):
            flush_buf.zero_()  # L2 flush (256 MiB > 126 MB L2), outside the step's events
            ev[i][0].record(ext)
            st = step()
            ev[i][1].record(ext)
            launches += (st.kernel_launches if st is not None else 1) + (1 if (world > 1 and peer_frames is None) else 0)
            if st is not None:
                kernel_ms["primary"] += st.ms_primary
                kernel_ms["bounce"] += st.ms_bounce
                kernel_ms["resolve"] += st.ms_resolve
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_end = time.time()
        if not tree:
            # the dominant kernel's own duration (CUDA events around the launch, on its stream), from extra untimed steps
            for _ in range(steps):
                st = step(want_stats=True)
                kernel_ms["primary"] += st.ms_primary
            torch.cuda.synchronize()
    per_step = [a.elapsed_time(b) for a, b in ev]
    dev_ms = sum(per_step)
    if world > 1:
        t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    out = {"desc": desc, "scene": scene_name, "kw": kw, "rays": rays, "p2p": world > 1 and peer_frames is not None, "ms_per_step": dev_ms / steps, "value": rays / (dev_ms / steps) / 1e3,
           "ms_per_step_min": min(per_step), "ms_per_step_median": sorted(per_step)[len(per_step) // 2],
           "wall_ms_per_step": (t_end - t_begin) * 1e3 / steps, "launches": launches, "stats": cst, "stats_rank0": cst_local, "t_begin": t_begin, "t_end": t_end,
           "kernel_ms_per_step": {k: v / steps for k, v in kernel_ms.items()},
           "primary_samples": base.width * base.height * (base.grid_size ** 2 if base.grid_size else 1)}

    if want_e2e:
        # end to end through the reference-facing call: host scene arrays -> skr_scene_upload, skr_render -> pinned host RGB8
        host = torch.empty((base.height, base.width, 3), dtype=torch.uint8).pin_memory().numpy()
        sc_bytes = sum(getattr(scene, f).nbytes for f in ("spheres", "tris", "plights", "dlights", "fogs", "camera", "ambient", "background"))
        opt1 = dataclasses.replace(base, rank=0, world=1)
        n_e2e = max(3, min(steps, 20))
        for _ in range(2):
            r.upload(scene)
            r.render(opt1, rgb8=host, want_rgb32=False)
        t0 = time.time()
        for _ in range(n_e2e):
            r.upload(scene)
            r.render(opt1, rgb8=host, want_rgb32=False)
        t1 = time.time()
        full_rays = rays if world == 1 else None
        if world > 1:
            # e2e at N > 1: every rank uploads, renders its tiles, gathers; rank 0 copies the frame to pinned host memory
            def e2e_step():
                r.upload(scene)
                step()
                if rank == 0:
                    torch.from_numpy(host).copy_(getattr(step, "frame", frame), non_blocking=False)
            with torch.cuda.stream(ext):
                e2e_step()
                torch.cuda.synchronize()
                dist.barrier()
                t0 = time.time()
                for _ in range(n_e2e):
                    e2e_step()
                torch.cuda.synchronize()
                dist.barrier()
                t1 = time.time()
            t = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t1 = t0 + float(t.item())
            full_rays = rays
        out["e2e"] = {"value": full_rays * n_e2e / (t1 - t0) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(sc_bytes) * (world if world > 1 else 1),
                      "d2h_bytes_per_step": int(host.nbytes), "ms_per_step": (t1 - t0) * 1e3 / n_e2e, "steps": n_e2e,
                      "path": "skr_scene_upload + skr_render (host arrays in, pinned host RGB8 out), wall clock"
                      if world == 1 else "per rank skr_scene_upload + frame split (same exchange as `value`), D2H of the frame on rank 0; wall clock, max over ranks"}
    return out


def main_gpu(args, rank, world, local_rank):
    import torch

    import skele_raytracer_b200 as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # NCCL prints its version banner to STDOUT; keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    r = S.Renderer(local_rank)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    steps, warmup = max(1, args.steps), max(3, args.warmup)

    fp32_peak = r.measure_fp32_peak(4096)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    m = measure_gpu(S, torch, dist, r, args.workload, steps, warmup, rank, world, flush_buf)
    clocks = sampler.stop(m["t_begin"], m["t_end"])

    others = {}
    if args.all_configs and world == 1:
        for w in WORKLOADS:
            if w == args.workload:
                continue
            k = 3 if w == "c5" else 10
            o = measure_gpu(S, torch, dist, r, w, k, 3, rank, world, flush_buf, want_e2e=False)
            fl = algorithmic_flops(o["stats"], o["primary_samples"])
            others[w] = {"workload": o["desc"], "ms_per_frame": o["ms_per_step"], "ms_per_frame_median": o["ms_per_step_median"],
                         "ms_per_frame_min": o["ms_per_step_min"], "mrays_per_s": o["value"], "rays_per_frame": o["rays"],
                         "fp32_tflops_algorithmic": fl / (o["ms_per_step"] * 1e-3) / 1e12, "frac_of_fp32_peak": fl / (o["ms_per_step"] * 1e-3) / 1e12 / fp32_peak,
                         "kernel_launches_per_frame": o["launches"] / k}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_run(args.workload, 12.0, 1, 0)

    if rank == 0:
        peaks, peaks_src = load_peaks()
        st = m["stats"]
        # roofline of the dominant kernel AS LAUNCHED ON RANK 0: its own share of the frame's work / its own duration
        flops = algorithmic_flops(m["stats_rank0"], m["primary_samples"] / world)
        # the dominant kernel: primary_kernel without --gillum, shade_expand_kernel with it
        dom = "bounce" if m["kw"].get("monte_carlo") else "primary"
        dom_ms = m["kernel_ms_per_step"][dom] or m["ms_per_step"]
        dom_launches = 1 if dom == "primary" else max(1, round(m["launches"] / steps))
        achieved = flops / (dom_ms * 1e-3) / 1e12
        # what the kernels executed: bundle culling proves most sphere tests of a jittered pixel unnecessary
        st0 = m["stats_rank0"]
        flops_exec = flops - 25.0 * (st0["sphere_tests"] - st0["sphere_tests_executed"])
        executed = flops_exec / (dom_ms * 1e-3) / 1e12
        frame_bytes = m["kw"]["width"] * m["kw"]["height"] * 3
        line = {
            "metric": "Mrays/s", "value": m["value"], "unit": "Mrays/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": m["desc"], "scene": m["scene"] + " (snapshot of the reference parser's Scene, tests/golden/scenes)", **m["kw"], "seed": SEED,
                       "rays_per_frame": m["rays"], "closest_hit_rays": st["closest_hit_rays"], "shadow_rays": st["shadow_rays"],
                       "l2": "flushed between steps with a 256 MiB memset, outside the per-step CUDA events", "timing": "CUDA events per step on the library stream, summed; max over ranks",
                       "frame_split": ("single GPU, whole frame" if world == 1 else
                                       f"{world} ranks, interleaved 32x32 tiles; finished pixels stored straight into every rank's frame over NVLink "
                                       "(skr_render_peers_device, torch symmetric memory), one symmetric-memory barrier per frame" if m["p2p"] else
                                       f"{world} ranks, interleaved 32x32 tiles, one NCCL all-gather of RGB8 tiles per frame + de-interleave kernel"),
                       "ms_per_step_median_this_rank": m["ms_per_step_median"], "ms_per_step_min_this_rank": m["ms_per_step_min"],
                       "wall_ms_per_step_incl_flush": m["wall_ms_per_step"]},
            "clocks": clocks,
            "e2e": m.get("e2e"),
            "gpu_launches": m["launches"],
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                         "traffic": NCU_TRAFFIC_BYTES.get(args.workload, (None, None))[0] if world == 1 else None,
                         "traffic_source": NCU_TRAFFIC_BYTES.get(args.workload, (None, None))[1], "kernel": "primary_kernel" if dom == "primary" else "shade_expand_kernel",
                         "kernel_ms_per_frame": dom_ms, "kernel_launches_per_frame": dom_launches,
                         "flops_per_frame_algorithmic_this_rank": flops,
                         "executed": {"tflops": executed, "frac": executed / fp32_peak, "flops_per_frame_this_rank": flops_exec,
                                      "sphere_tests_algorithmic": st0["sphere_tests"], "sphere_tests_executed": st0["sphere_tests_executed"],
                                      "note": "`achieved` counts the reference algorithm's arithmetic (every sphere per query); conservative bundle culling "
                                              "skips tests that provably fail, so the FP32 pipe executed only this much"},
                         "peak_source": "FFMA microbenchmark measured live in this run (skr_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 figure",
                         "note": "compute-bound FP32 CUDA-core path (no dense contraction -> tensor cores unused); algorithmic HBM traffic is the RGB8 frame only",
                         "hbm": {"algorithmic_bytes_per_frame": frame_bytes, "achieved_gbs": frame_bytes / (m["ms_per_step"] * 1e-3) / 1e9,
                                 "peak_gbs": peaks.get("hbm_gbs"), "peak_source": peaks_src}},
            "cpu_baseline": None if cpu is None else {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
        }
        if others:
            line["other_configs"] = others
        emit(line)
    r.close()
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _quiet_stdout():
    """Everything libraries print to file descriptor 1 (NCCL's version banner, for one) goes to stderr from here on; the
    JSON line is written to the saved descriptor at the end, so stdout carries exactly one line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--all-configs", action="store_true", help="also measure the other BASELINE.json configs (N=1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        main_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd, stdout=_REAL_STDOUT))
    main_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
