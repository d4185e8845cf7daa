#!/usr/bin/env python
"""bench.py -- measurement of the per-pixel tracing loop on B200: headline line + every BASELINE config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c1..c5] [--only-headline]

A "step" is one frame of a workload.  Headline workload = BASELINE.json configs[1]: scenes/spheres2.scn 1920x1080
--jsample 5 --shadow (the scene snapshot in tests/golden/scenes, produced by the reference's own parser).  Metric:
Mrays/s, rays = closest-hit rays (shade() invocations with depth > 0) + distinct shadow rays, counted on
the device in an untimed pass with the same seed.

  value   : whole-job Mrays/s, scene resident in HBM, frame left in HBM (device time, CUDA events on the library's
            stream, per step, L2 flushed between steps outside the events; max over ranks).
  e2e     : the same metric through the reference-facing call with HOST buffers, wall clock: N = 1: skr_scene_upload
            (H2D) + skr_render into pinned host memory, every step.  N > 1: every rank uploads the scene and renders its
            tiles with the kernel storing each finished pixel straight into ONE page-locked host frame shared by the
            ranks (a /dev/shm segment, skr_pin_host), each GPU over its own PCIe link; a barrier ends the step.
  N > 1   : the frame is split into interleaved 32x32 tiles over the ranks (one process per GPU, torchrun); each rank's
            render kernel stores its finished pixels straight into rank 0's frame over NVLink (torch symmetric memory,
            skr_render_peers_device) and one symmetric-memory barrier ends the frame; where peer mapping is unavailable
            (or SKR_BENCH_NO_P2P=1): ONE all-gather (NCCL) of the RGB8 tiles + de-interleave kernel.
            Total work is fixed -> "strong".
  configs : the same measurement (device time, e2e, roofline, CPU baseline) for EVERY BASELINE.json config c1..c5, at
            every N; the headline workload's entry repeats the top-level figures.
  roofline: sphere scenes: FP32 CUDA-core pipe (no dense contraction on this path; tensor cores unused; HBM traffic is
            the framebuffer only), peak = FMA microbenchmark measured live in this run.  `achieved` / `frac` = the
            arithmetic the kernels EXECUTED; `algorithmic` = the reference algorithm's count (every sphere per query),
            which conservative bundle culling proves partly unnecessary.  Triangle scene (c4): the BVH / primitive
            fetch goes through L1 -> bound "l1", bytes = 64 B per node visit + 48 B per leaf test, against L1 / L2 /
            shared-memory read bandwidths measured live (skr_measure_bandwidth) and the HBM figure of MEASURED_PEAKS.json.
  cpu_baseline: the reference's own code (oracle/_ref, compiled from /root/reference/src) on the host cores, on a bounded
            row window of the same frame: the -O2 build and the reference's own flags (-g, no -O: src/Makefile:2).
  --impl reference : that CPU code as the timed arm (all host threads), same config / metric / unit.
"""
import argparse
import dataclasses
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
MAX_STEPS = {"c1": 20, "c2": 1000, "c3": 10, "c4": 20, "c5": 3}   # timed steps per config in the `configs` block
CPU_BUDGET_S = {"c1": 1.0, "c2": 2.0, "c3": 2.0, "c4": 2.0, "c5": 2.0}  # seconds per sample and build; the headline gets 8 s
SEED = 1
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed ncu captures of this
# round (bench.py cannot run ncu on itself).  Tree configs launch the kernel several times per frame with different loads:
# the figure is the mean over the captured launches.
NCU_TRAFFIC_BYTES = {
    "c2": (595968 + 94720, "profiles/r02_c2_ncu_raw.txt (primary_kernel<GI=0,STATS=0,SMEM=1,TRIS=0,FOG=1,HALVES=1>)"),
    "c3": (int((838.494 + 3154.648 + 2658.218 + 966.383 + 759.633) * 1e6 / 5),
           "profiles/r02_c3_ncu_raw.txt (mean of 5 shade_expand_kernel launches of the final build: queue entries in, queue entries out)"),
    "c4": (477184 + 768, "profiles/r02_c4_ncu_raw.txt (primary_kernel<GI=0,STATS=0,SMEM=1,TRIS=1,FOG=0,HALVES=0>)"),
    "c5": (int((114.027776 + 274.185472 + 328.451072 + 8.628224 + 115.257088 + 149.899776) * 1e6 / 3),
           "profiles/r02_c5_ncu_raw.txt (mean of 3 shade_expand_kernel launches: expand, leaves in place, expand)"),
}
STAT_KEYS = ["closest_hit_rays", "shadow_rays", "sphere_tests", "sphere_tests_pos", "tri_tests", "bvh_node_visits", "sphere_hits", "light_evals",
             "sphere_tests_executed"]


def algorithmic_flops(st, primary_samples, executed=False):
    """SURVEY 8(d) per-unit figures x the device counters of one frame (FMA = 2, everything else = 1).
    executed=True: sphere tests the kernels really ran (bundle culling) instead of the reference algorithm's count."""
    tests = st["sphere_tests_executed"] if executed else st["sphere_tests"]
    return (25.0 * tests + 8.0 * st["sphere_tests_pos"] + 18.0 * st["sphere_hits"]   # F_ch(S) = 25 S + 8 h + 18 [hit]
            + 15.0 * st["shadow_rays"]                                               # shadow ray set-up (tests counted above at 25/33)
            + 33.0 * st["sphere_hits"] + 89.0 * st["light_evals"]                    # direct shading 33 + 89 L
            + 24.0 * st["bvh_node_visits"] + 35.0 * st["tri_tests"]                  # line-slab test / triangle test
            + 20.0 * primary_samples)                                                # ray generation


def fetch_bytes(st):
    """Bytes the triangle path pulls through L1 per frame: one 64 B node (both child boxes) per node visit, 48 B of
    vertices per leaf test (SURVEY 8d states 32 B per child box)."""
    return 64.0 * st["bvh_node_visits"] + 48.0 * st["tri_tests"]


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
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores
# ------------------------------------------------------------------------------------------------

def cpu_reference_run(workload, budget_s, steps, warmup, force_port=False, unoptimised=False):
    """Times `steps` (+ `warmup`) bounded samples of the workload on the host cores.  A sample is the SAME frame -- same
    scene, camera, aspect and flags -- at 1/k^2 of the pixels (k = 1: the full frame), k chosen from a probe so that one
    sample takes about budget_s; rays per pixel do not depend on the resolution, so the rate (Mrays/s) is the full
    frame's.  -> dict(value Mrays/s, kind, cores, sample, ms_per_step, frame_s_extrapolated)."""
    from oracle import oracle_lib as O

    scene_name, kw, desc = WORKLOADS[workload]
    sc = O.Scene.load(os.path.join(GOLD, scene_name + ".npz"))
    full = O.Options(**kw)
    port = O.Port()
    use_ref = O.ref_available() and not force_port
    if unoptimised and not (use_ref and os.path.exists(O.REF_O0_SO)):
        return None
    ref = O.Ref(unoptimised=unoptimised) if use_ref else None
    cores = ref.max_threads() if use_ref else port.max_threads()

    def at_scale(k):
        return dataclasses.replace(full, width=max(1, full.width // k), height=max(1, full.height // k))

    def run(opt):
        if use_ref:
            return ref.render(sc, opt, seed=SEED, threads=cores)[2]
        return port.render(sc, opt, rng_mode=O.RNG_PHILOX, seed=SEED, threads=cores, want_rgb8=False)[3]

    # two probes: a tiny frame (config 5 costs ~10 ms of CPU per PIXEL), then one sized for ~1/8 of the budget
    probe = at_scale(160)                                    # 12x6 for 1080p, 24x13 for 4K
    t_px = max(run(probe), 1e-5) / (probe.width * probe.height)
    k = 1
    while k < 160 and (full.width // k) * (full.height // k) * t_px > budget_s / 8:
        k += 1
    if k < 160:
        probe = at_scale(k)
        t_px = max(run(probe), 1e-5) / (probe.width * probe.height)
    k = 1
    while k < 160 and (full.width // k) * (full.height // k) * t_px > budget_s:
        k += 1
    opt = at_scale(k)
    for _ in range(3):                                       # a sample far below the budget (probe noise): grow it
        t = run(opt)
        if k == 1 or t > budget_s / 3:
            break
        k = max(1, min(k - 1, int(k * (max(t, 1e-4) / budget_s) ** 0.5)))
        opt = at_scale(k)
    # rays of the sample: counted by the port on the same frame (same algorithm; keyed RNG)
    st = port.render(sc, opt, rng_mode=O.RNG_PHILOX, seed=SEED, want_rgb8=False)[2]
    rays = st["closest_hit_rays"] + st["shadow_rays"]
    for _ in range(warmup):
        run(opt)
    times = [run(opt) for _ in range(steps)]
    tot = sum(times)
    build = ("reference's own src/ compiled -O2 -fopenmp (oracle/_ref), OMP threads share rand()" if use_ref and not unoptimised else
             "reference's own src/ compiled with ITS flags (-g, no -O, src/Makefile:2) -fopenmp (oracle/_ref)" if use_ref else
             "C port oracle/skr_oracle.c -O2 -fopenmp")
    frac = (opt.width * opt.height) / float(full.width * full.height)
    return {"value": rays * steps / tot / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "reference" if use_ref else "port",
            "sample": f"the frame of {desc} at {opt.width}x{opt.height} ({frac:.4f} of the {full.width}x{full.height} pixels; same scene, camera, aspect and "
                      f"flags): {rays} rays per sample, {steps} timed samples, render loop only (no parse, no PPM write); " + build,
            "ms_per_step": tot / steps * 1e3, "rays_per_sample": rays, "sample_size": [opt.width, opt.height],
            "frame_s_extrapolated": tot / steps / frac}


def main_reference(args, rank):
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    budget = max(0.5, min(6.0, 120.0 / (steps + warmup)))
    r = cpu_reference_run(args.workload, budget, steps, warmup)
    scene_name, kw, desc = WORKLOADS[args.workload]
    line = {"impl": "reference", "metric": "Mrays/s", "value": r["value"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "scene": scene_name, **kw, "seed": SEED,
                       "note": f"each step is the same frame at {r['sample_size'][0]}x{r['sample_size'][1]} (bounded sample, see cpu_baseline.sample)"},
            "cpu_baseline": {"value": r["value"], "unit": "Mrays/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------

class SharedHostFrame:
    """ONE host frame for all ranks of the box: a /dev/shm segment every rank maps and page-locks (skr_pin_host), so that
    each rank's render kernel stores its own tiles into it over its own PCIe link."""

    def __init__(self, torch, dist, r, nbytes, rank, world):
        import mmap

        self.r, self.rank = r, rank
        name = [f"/dev/shm/skr_bench_{os.getpid()}_{nbytes}" if rank == 0 else None]
        if world > 1:
            dist.broadcast_object_list(name, src=0)
        self.path = name[0]
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(nbytes)
        if world > 1:
            dist.barrier()
        self.f = open(self.path, "r+b")
        self.mm = mmap.mmap(self.f.fileno(), nbytes)
        import ctypes
        import numpy as np
        self.np = np.frombuffer(self.mm, dtype=np.uint8)
        self.host_ptr = ctypes.addressof(ctypes.c_char.from_buffer(self.mm))
        self.dev_ptr, self.registered = r.pin_host(self.host_ptr, nbytes)

    def close(self, dist, world):
        try:
            if self.registered:
                self.r.unpin_host(self.host_ptr)
        except Exception:
            pass
        if world > 1:
            dist.barrier()
        self.np = None
        try:
            self.mm.close()
        except BufferError:
            pass
        self.f.close()
        if self.rank == 0:
            try:
                os.remove(self.path)
            except OSError:
                pass


def measure_gpu(S, torch, dist, r, workload, steps, warmup, rank, world, flush_buf, want_e2e=True):
    scene_name, kw, desc = WORKLOADS[workload]
    scene = S.Scene.load(os.path.join(GOLD, scene_name + ".npz"))
    r.upload(scene)
    base = S.Options(seed=SEED, rank=rank, world=world, **kw)
    dev = torch.device("cuda", torch.cuda.current_device())
    ext = torch.cuda.ExternalStream(r.stream(), device=dev)
    t_first0 = time.time()
    r.reserve(base)   # queue arena / accumulators / kernels loaded ahead of the first frame
    reserve_ms = (time.time() - t_first0) * 1e3
    # the first frame of this context with these options (device time of the whole render, as every later frame is timed)
    if world == 1:
        first = r.render_device(base, 0, 0)
        tiles = gathered = None
    else:
        tiles = torch.empty(r.tiles_bytes(base), dtype=torch.uint8, device=dev)
        first = r.render_tiles_device(base, tiles.data_ptr())

    # untimed counting pass (same seed -> same rays)
    if world == 1:
        cst = r.render_device(dataclasses.replace(base, collect_stats=True), 0, 0).as_dict()
        cst_local = dict(cst)
        peer_frames = None
    else:
        cst = r.render_tiles_device(dataclasses.replace(base, collect_stats=True), tiles.data_ptr()).as_dict()
        cst_local = dict(cst)
        t = torch.tensor([float(cst[k]) for k in STAT_KEYS], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        for k, v in zip(STAT_KEYS, t.tolist()):
            cst[k] = int(v)
        gathered = torch.empty(tiles.numel() * world, dtype=torch.uint8, device=dev)
        from skele_raytracer_b200.distributed import PeerFrames
        peer_frames = None if os.environ.get("SKR_BENCH_NO_P2P") == "1" else PeerFrames.create(base.height, base.width, dev)
        ok = torch.tensor([1 if peer_frames is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # every rank takes the same path
        if int(ok.item()) == 0:
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
            # collective-free: P2P stores of the finished pixels into rank 0's frame + symmetric-memory barrier
            step.frame, st = peer_frames.render(r, base, rank, world, want_stats=ws, targets=[0])
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
    out = {"workload": workload, "desc": desc, "scene": scene_name, "kw": kw, "rays": rays, "p2p": world > 1 and peer_frames is not None,
           "ms_per_step": dev_ms / steps, "value": rays / (dev_ms / steps) / 1e3, "steps": steps, "warmup": warmup,
           "ms_per_step_min": min(per_step), "ms_per_step_median": sorted(per_step)[len(per_step) // 2],
           "wall_ms_per_step": (t_end - t_begin) * 1e3 / steps, "launches": launches, "stats": cst, "stats_rank0": cst_local, "t_begin": t_begin, "t_end": t_end,
           "kernel_ms_per_step": {k: v / steps for k, v in kernel_ms.items()},
           "first_frame_ms": first.ms_total, "reserve_ms": reserve_ms,
           "primary_samples": base.width * base.height * (base.grid_size ** 2 if base.grid_size else 1)}

    if want_e2e:
        sc_bytes = sum(getattr(scene, f).nbytes for f in ("spheres", "tris", "plights", "dlights", "fogs", "camera", "ambient", "background"))
        n_e2e = max(3, min(steps, 20))
        nbytes = base.height * base.width * 3
        if world == 1:
            # end to end through the reference-facing call: host scene arrays -> skr_scene_upload, skr_render -> pinned host RGB8
            host = torch.empty((base.height, base.width, 3), dtype=torch.uint8).pin_memory().numpy()
            opt1 = dataclasses.replace(base, rank=0, world=1)
            for _ in range(2):
                r.upload(scene)
                r.render(opt1, rgb8=host, want_rgb32=False)
            t0 = time.time()
            for _ in range(n_e2e):
                r.upload(scene)
                r.render(opt1, rgb8=host, want_rgb32=False)
            t1 = time.time()
            path = "skr_scene_upload + skr_render (host arrays in, pinned host RGB8 out), wall clock"
        else:
            # Every rank: upload + render its tiles; a finished pixel is stored (over NVLink, while the kernels trace) into the
            # device buffer of the rank that OWNS its band of rows (skr_render_bands_device); after the symmetric-memory barrier
            # each rank holds its band complete and copies it into ONE page-locked host frame shared by the ranks (a /dev/shm
            # segment, skr_pin_host) over its own PCIe link; a second barrier ends the step.  Without peer mapping: the kernels
            # store straight into the host frame.
            shf = SharedHostFrame(torch, dist, r, nbytes, rank, world)
            rows = 0 if peer_frames is None else peer_frames.band_rows(base.height, world)
            row_bytes = base.width * 3

            if peer_frames is not None:
                hdl, buf = peer_frames.hdls[0], peer_frames.bufs[0]
                ptrs = list(hdl.buffer_ptrs)
                y0, y1 = min(base.height, rank * rows), min(base.height, (rank + 1) * rows)
                src, dst, nb = buf.data_ptr() + y0 * row_bytes, shf.host_ptr + y0 * row_bytes, (y1 - y0) * row_bytes

            def e2e_step():
                # (called with the library's stream current: the barriers are enqueued behind the render and the copy)
                r.upload(scene)
                if peer_frames is None:
                    r.render_peers_device(base, [shf.dev_ptr], want_stats=tree)
                    r.sync()
                    dist.barrier()
                    return
                r.render_bands_device(base, ptrs, rows, want_stats=tree)
                hdl.barrier()                      # every rank's kernel is done: this rank's band is whole
                r.copy_to_host(dst, src, nb)       # ... and leaves over this rank's own PCIe link
                hdl.barrier(channel=1)             # every rank's copy is done
                r.sync()

            with torch.cuda.stream(ext):
                e2e_step()
                e2e_step()
                torch.cuda.synchronize()
                dist.barrier()
                t0 = time.time()
                for _ in range(n_e2e):
                    e2e_step()
                t1 = time.time()
            t = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t1 = t0 + float(t.item())
            dist.barrier()
            if rank == 0:
                # the frame the ranks assembled in host memory must be the frame one GPU renders
                whole = torch.empty((base.height, base.width, 3), dtype=torch.uint8, device=dev)
                r.render_device(dataclasses.replace(base, rank=0, world=1), whole.data_ptr(), 0)
                out["e2e_frame_identical_to_one_gpu"] = bool((whole.cpu().numpy().reshape(-1) == shf.np).all())
            shf.close(dist, world)
            path = ("per rank skr_scene_upload + skr_render_bands_device (pixels stored over NVLink into the band owner's memory), symmetric-memory "
                    "barrier, each rank copies its band of rows into ONE page-locked host frame shared by the ranks over its own PCIe link, second "
                    "barrier + stream sync; wall clock, max over ranks" if peer_frames is not None else
                    "per rank skr_scene_upload + skr_render_peers_device storing straight into ONE page-locked host frame shared by the ranks, stream "
                    "sync + barrier per step; wall clock, max over ranks")
        out["e2e"] = {"value": rays * n_e2e / (t1 - t0) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(sc_bytes) * world,
                      "d2h_bytes_per_step": int(nbytes), "ms_per_step": (t1 - t0) * 1e3 / n_e2e, "steps": n_e2e, "path": path}
        if "e2e_frame_identical_to_one_gpu" in out:
            out["e2e"]["host_frame_identical_to_one_gpu"] = out.pop("e2e_frame_identical_to_one_gpu")
    return out


def measure_native(S, torch, workload, n_gpus, steps):
    """e2e through the shipped single-process multi-GPU C API (include/skr_mgpu.h): skr_mgpu_scene_upload + skr_mgpu_render
    into pinned host memory every step, wall clock.  Run by rank 0 alone (it drives all n_gpus GPUs itself) while the other
    ranks wait at a HOST barrier."""
    import numpy as np  # noqa: F401

    scene_name, kw, desc = WORKLOADS[workload]
    scene = S.Scene.load(os.path.join(GOLD, scene_name + ".npz"))
    opt = S.Options(seed=SEED, **kw)
    m = S.MgpuRenderer(n_gpus)
    try:
        host = torch.empty((opt.height, opt.width, 3), dtype=torch.uint8).pin_memory().numpy()
        for _ in range(2):
            m.upload(scene)
            m.render(opt, host)
        t0 = time.time()
        for _ in range(steps):
            m.upload(scene)
            m.render(opt, host)
        t1 = time.time()
        return {"ms_per_step": (t1 - t0) * 1e3 / steps, "steps": steps, "frame_nonzero": bool(host.any()),
                "path": f"skr_mgpu_scene_upload + skr_mgpu_render (one process, {n_gpus} GPUs, every GPU stores its tiles into the pinned host frame "
                        "over its own PCIe link), wall clock"}
    finally:
        m.close()


def roofline_of(m, world, peaks, peaks_src, fp32_peak, bw):
    """Roofline of the dominant kernel AS LAUNCHED ON RANK 0: its own share of the frame's work / its own duration."""
    st0 = m["stats_rank0"]
    kw = m["kw"]
    dom = "bounce" if kw.get("monte_carlo") else "primary"
    dom_ms = m["kernel_ms_per_step"][dom] or m["ms_per_step"]
    dom_launches = 1 if dom == "primary" else max(1, round(m["launches"] / m["steps"]) - 2)
    frame_bytes = kw["width"] * kw["height"] * 3
    hbm = {"algorithmic_bytes_per_frame": frame_bytes, "achieved_gbs": frame_bytes / (m["ms_per_step"] * 1e-3) / 1e9,
           "peak_gbs": peaks.get("hbm_gbs"), "peak_source": peaks_src}
    traffic = NCU_TRAFFIC_BYTES.get(m["workload"], (None, None)) if world == 1 else (None, None)
    if st0["sphere_tests"] == 0 and st0["bvh_node_visits"] > 0:
        # triangle scene: memory hierarchy (the whole BVH is cache resident; the fetch goes through L1)
        b = fetch_bytes(st0)
        gbs = b / (dom_ms * 1e-3) / 1e9
        fl = algorithmic_flops(st0, m["primary_samples"] / world)
        return {"bound": "l1", "achieved": gbs, "peak": bw["l1_gbs"], "unit": "GB/s", "frac": gbs / bw["l1_gbs"] if bw["l1_gbs"] > 0 else None,
                "traffic": traffic[0], "traffic_source": traffic[1], "kernel": "primary_kernel<TRIS=1> (LBVH line any-hit query walked in place)", "kernel_ms_per_frame": dom_ms,
                "kernel_launches_per_frame": 1, "bytes_per_frame_algorithmic_this_rank": b,
                "bytes_per_unit": "64 B per BVH node visit (both child boxes) + 48 B per triangle leaf test",
                "levels": {"l1": {"achieved_gbs": gbs, "peak_gbs": bw["l1_gbs"]}, "l2": {"peak_gbs": bw["l2_gbs"]}, "shared": {"peak_gbs": bw["lds_gbs"]}, "hbm": hbm},
                "fp32": {"tflops": fl / (dom_ms * 1e-3) / 1e12, "frac": fl / (dom_ms * 1e-3) / 1e12 / fp32_peak},
                "peak_source": "ld.global.ca microbenchmark measured live in this run (skr_measure_bandwidth level 1)",
                "note": "latency / divergence bound traversal: the 1.1 MB hierarchy is L1/L2 resident, HBM traffic is the frame"}
    flops_alg = algorithmic_flops(st0, m["primary_samples"] / world)
    flops_exec = algorithmic_flops(st0, m["primary_samples"] / world, executed=True)
    achieved = flops_exec / (dom_ms * 1e-3) / 1e12
    alg = flops_alg / (dom_ms * 1e-3) / 1e12
    return {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
            "traffic": traffic[0], "traffic_source": traffic[1], "kernel": "primary_kernel" if dom == "primary" else "shade_expand_kernel",
            "kernel_ms_per_frame": dom_ms, "kernel_launches_per_frame": dom_launches, "flops_per_frame_executed_this_rank": flops_exec,
            "algorithmic": {"tflops": alg, "frac": alg / fp32_peak, "flops_per_frame_this_rank": flops_alg,
                            "sphere_tests_algorithmic": st0["sphere_tests"], "sphere_tests_executed": st0["sphere_tests_executed"],
                            "note": "the reference algorithm's arithmetic (every sphere per query); conservative bundle culling skips tests that "
                                    "provably fail, so `achieved` counts only what the FP32 pipe executed"},
            "peak_source": "FFMA microbenchmark measured live in this run (skr_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 figure",
            "note": "compute-bound FP32 CUDA-core path (no dense contraction -> tensor cores unused); algorithmic HBM traffic is the RGB8 frame only",
            "hbm": hbm}


def frame_split_text(m, world):
    if world == 1:
        return "single GPU, whole frame"
    if m["p2p"]:
        return (f"{world} ranks, interleaved 32x32 tiles; finished pixels stored straight into rank 0's frame over NVLink "
                "(skr_render_peers_device, torch symmetric memory), one symmetric-memory barrier per frame")
    return f"{world} ranks, interleaved 32x32 tiles, one NCCL all-gather of RGB8 tiles per frame + de-interleave kernel"


def main_gpu(args, rank, world, local_rank):
    import torch

    import skele_raytracer_b200 as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dist = None
    host_group = None
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # NCCL prints its version banner to STDOUT; keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        host_group = dist.new_group(backend="gloo")  # host-side barriers (no kernel spinning on a GPU while rank 0 drives all of them)
    r = S.Renderer(local_rank)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    steps, warmup = max(1, args.steps), max(3, args.warmup)

    fp32_peak = r.measure_fp32_peak(4096)
    bw = {"lds_gbs": r.measure_bandwidth(0), "l1_gbs": r.measure_bandwidth(1), "l2_gbs": r.measure_bandwidth(2)}
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    m = measure_gpu(S, torch, dist, r, args.workload, steps, warmup, rank, world, flush_buf)
    clocks = sampler.stop(m["t_begin"], m["t_end"])

    results = {args.workload: m}
    if not args.only_headline:
        for w in WORKLOADS:
            if w not in results:
                results[w] = measure_gpu(S, torch, dist, r, w, max(1, min(steps, MAX_STEPS[w])), 3, rank, world, flush_buf)

    # the shipped single-process multi-GPU API over the same GPUs: rank 0 drives all of them, the others wait on the host
    native = {}
    if world > 1 and not args.no_native:
        torch.cuda.synchronize()
        dist.barrier(group=host_group)
        if rank == 0:
            for w in results:
                try:
                    nv = measure_native(S, torch, w, world, max(3, min(results[w]["steps"], 10)))
                    nv["value"] = results[w]["rays"] / (nv["ms_per_step"] * 1e-3) / 1e6
                    nv["unit"] = "Mrays/s"
                    native[w] = nv
                except Exception as e:  # reported, never fatal for the distributed numbers
                    native[w] = {"error": str(e)[:300]}
        dist.barrier(group=host_group)

    cpu = {}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        for w in results:
            head = w == args.workload
            o2 = cpu_reference_run(w, 8.0 if head else CPU_BUDGET_S[w], 1, 0)
            o0 = cpu_reference_run(w, CPU_BUDGET_S[w], 1, 0, unoptimised=True)
            cpu[w] = (o2, o0)

    if rank == 0:
        peaks, peaks_src = load_peaks()
        st = m["stats"]

        def cpu_block(w):
            if w not in cpu:
                return None
            o2, o0 = cpu[w]
            blk = {k: o2[k] for k in ("value", "unit", "cores", "kind", "sample")}
            blk["frame_s_extrapolated"] = o2["frame_s_extrapolated"]
            if o0 is not None:
                blk["reference_flags_build"] = {"value": o0["value"], "unit": "Mrays/s", "flags": "-g, no -O (src/Makefile:2)", "cores": o0["cores"],
                                                "sample": o0["sample"], "frame_s_extrapolated": o0["frame_s_extrapolated"]}
            return blk

        configs = {}
        for w in sorted(results):
            o = results[w]
            configs[w] = {"workload": o["desc"], "ms_per_frame": o["ms_per_step"], "ms_per_frame_median_this_rank": o["ms_per_step_median"],
                          "ms_per_frame_min_this_rank": o["ms_per_step_min"], "mrays_per_s": o["value"], "rays_per_frame": o["rays"], "steps": o["steps"],
                          "warmup": o["warmup"], "kernel_launches_per_frame": o["launches"] / o["steps"], "first_frame_ms": o["first_frame_ms"],
                          "reserve_ms": o["reserve_ms"], "e2e": o.get("e2e"), "e2e_native": native.get(w), "frame_split": frame_split_text(o, world),
                          "roofline": roofline_of(o, world, peaks, peaks_src, fp32_peak, bw), "cpu_baseline": cpu_block(w)}
        line = {
            "metric": "Mrays/s", "value": m["value"], "unit": "Mrays/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": m["desc"], "scene": m["scene"] + " (snapshot of the reference parser's Scene, tests/golden/scenes)", **m["kw"], "seed": SEED,
                       "rays_per_frame": m["rays"], "closest_hit_rays": st["closest_hit_rays"], "shadow_rays": st["shadow_rays"],
                       "l2": "flushed between steps with a 256 MiB memset, outside the per-step CUDA events", "timing": "CUDA events per step on the library stream, summed; max over ranks",
                       "frame_split": frame_split_text(m, world),
                       "ms_per_step_median_this_rank": m["ms_per_step_median"], "ms_per_step_min_this_rank": m["ms_per_step_min"],
                       "wall_ms_per_step_incl_flush": m["wall_ms_per_step"]},
            "clocks": clocks,
            "e2e": m.get("e2e"),
            "e2e_native": native.get(args.workload),
            "gpu_launches": m["launches"],
            "roofline": roofline_of(m, world, peaks, peaks_src, fp32_peak, bw),
            "cpu_baseline": None if args.workload not in cpu else {k: cpu[args.workload][0][k] for k in ("value", "unit", "cores", "kind", "sample")},
            "build": S.build_info(),
            "peaks_measured_live": {"fp32_tflops": fp32_peak, **bw, "hbm_gbs": peaks.get("hbm_gbs"), "hbm_source": peaks_src},
            "configs": configs,
        }
        emit(line)
    r.close()
    if world > 1:
        dist.barrier()
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
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="headline workload (default: BASELINE.json configs[1])")
    ap.add_argument("--only-headline", action="store_true", help="skip the `configs` block (the other BASELINE.json configs)")
    ap.add_argument("--all-configs", action="store_true", help="(default now; kept for compatibility)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-native", action="store_true", help="skip e2e_native (skr_mgpu_render driven by rank 0) at N > 1")
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
