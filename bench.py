#!/usr/bin/env python
"""Benchmark of the smoothing step (BASELINE.json metric: smoothing steps/s and
vertex-updates/s at ~10M vertices, achieved HBM GB/s against peak).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one iteration of the optimize() loop on the device: fused point update
(K1) + flip-until-Delaunay.  At N=1 the workload is BASELINE.json configs[1]: CVT
block-diagonal on a ~10M-vertex RANDOM disk mesh, fp64 (generators.disk_gpu: mapped grid +
120 random-walk rounds with flips on the device; Qhull needs 270 s for this size on the host).
Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "vertex-updates/s"
DEFAULT_GRID = 3154  # disk_gpu(3154): 9,947,716 vertices, 19,882,818 cells
WALK_ROUNDS = 120    # random-walk rounds of disk_gpu (vertex degrees like a Qhull random mesh)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--method", default=None)
    ap.add_argument("--omega", type=float, default=None)
    ap.add_argument("--grid", type=int, default=None, help="n of disk_gpu(n) per GPU")
    ap.add_argument("--rounds", type=int, default=WALK_ROUNDS,
                    help="random-walk rounds of the mesh generator (0: jittered mapped grid)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config5", action="store_true",
                    help="skip the config-5 block (Lloyd omega=2, 100M vertices, strong scaling)")
    ap.add_argument("--config5-grid", type=int, default=10000)
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload(args):
    """BASELINE.json configs[1] (CVT block-diagonal, ~10M vertices, fp64) per GPU: the mesh
    grows with the number of GPUs (weak scaling), each rank updates its share of it."""
    method = args.method or "cvt-block-diagonal"
    omega = 1.0 if args.omega is None else args.omega
    return method, omega, args.grid or DEFAULT_GRID


def alg_bytes(n, c, d):
    # coords read + written once, int32 cells read once, 1-byte pin flag
    return (16 * d + 1) * n + 12 * c


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML during the timed region."""

    def __init__(self, gpu_index, period_s=0.01):
        import threading

        self.idx = gpu_index
        self.period = period_s
        self.sm, self.reasons = [], set()
        self.smax = None
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if visible:
                ids = [v for v in visible.split(",") if v.strip() != ""]
                if gpu_index < len(ids) and ids[gpu_index].strip().isdigit():
                    phys = int(ids[gpu_index])
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._ok = True
        except Exception:
            self._ok = False

    def _sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4),
                          ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
                          ("hw_power_brake_slowdown", 0x80)):
            if r & bit:
                self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self._ok:
            self._thread.start()

    def stop(self):
        if self._ok:
            self._stop.set()
            self._thread.join(timeout=2)
            try:
                self._sample()
            except Exception:
                pass
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.smax, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.smax,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def random_disk_host(n_target):
    """Random disk mesh of about n_target vertices built on the host alone (Qhull): the
    reference's own workload class, used by the CPU arms."""
    from optimesh_b200 import generators as G

    nb = max(8, int(round(2.0 * np.pi / np.sqrt(4.0 * np.pi / (np.sqrt(3.0) * n_target * 2.0)))))
    return G.disk(nb, 0), nb


# --------------------------------------------------------------------------- CPU arms
def _cpu_worker(method, omega, n_target, steps, warmup, start, out):
    """One replica of the CPU arm (its own process): builds its sample mesh, waits for the
    others, steps it.  The oracle is numpy (gather/bincount: one core per process)."""
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    import oracle

    (pts, cells), nb = random_disk_host(n_target)
    mesh = oracle.MeshTri(pts, cells)
    mesh.flip_until_delaunay()
    for _ in range(warmup):
        oracle.driver.step(mesh, method, omega=omega)
        mesh.flip_until_delaunay()
    start.wait()
    t0 = time.time()
    for _ in range(steps):
        oracle.driver.step(mesh, method, omega=omega)
        mesh.flip_until_delaunay()
    out.put((t0, time.time(), pts.shape[0], cells.shape[0], nb))


def cpu_workers(n_target):
    """How many replicas the host takes: one per core, bounded by memory (about 4 kB per vertex
    of oracle state and temporaries) and by 64."""
    cores = os.cpu_count() or 1
    try:
        import psutil

        mem = psutil.virtual_memory().available
        cores = min(cores, len(os.sched_getaffinity(0)))
    except Exception:
        mem = 16 << 30
    return int(max(1, min(cores, 64, 0.5 * mem / (4096.0 * n_target))))


def cpu_step_rate(method, omega, n_target, steps, warmup, workers=None):
    """The oracle (numpy port of the reference's algorithm) timed on ALL host cores: the
    reference's loop is a single process, so the host is filled with independent replicas of the
    same bounded sample -- update + limiter + flip-until-Delaunay on a random disk mesh -- that
    start together; the rate is all their vertex updates over the span from the first start to
    the last finish.  Returns (rate, seconds, vertices, cells, nb, workers)."""
    import multiprocessing as mp

    workers = workers or cpu_workers(n_target)
    ctx = mp.get_context("spawn")  # the GPU arm has CUDA initialised: no fork
    start = ctx.Barrier(workers)
    out = ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker,
                         args=(method, omega, n_target, steps, warmup, start, out))
             for _ in range(workers)]
    for pr in procs:
        pr.start()
    res = []
    try:
        for _ in range(workers):
            res.append(out.get(timeout=1200))
    finally:
        for pr in procs:
            pr.join(timeout=30)
            if pr.is_alive():
                pr.kill()
    t0 = min(r[0] for r in res)
    t1 = max(r[1] for r in res)
    n, c, nb = res[0][2], res[0][3], res[0][4]
    return n * steps * workers / (t1 - t0), t1 - t0, n, c, nb, workers


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    method, omega, grid = workload(args)
    total = args.steps + args.warmup
    # size the sample so the whole run stays within a few minutes: the numpy port costs 6-12 us
    # per vertex and step and core at these sizes (more on larger meshes: cache misses)
    budget_s = 120.0
    n_target = int(min(4.0e5, max(2.0e4, budget_s / (total * 12e-6))))
    n_target = int(os.environ.get("OM_BENCH_REF_VERTICES", n_target))  # (tests: a small sample)
    v, dt, n, c, nb, workers = cpu_step_rate(method, omega, n_target, args.steps, args.warmup)
    sample = (f"{workers} replicas (one per host core) of a random disk mesh disk({nb}) (Qhull): "
              f"{n} vertices / {c} cells each, {args.steps} steps of {method} (omega={omega}) "
              f"incl. limiter and flip-until-Delaunay after {args.warmup} warm-up steps")
    line = {
        "impl": "reference",
        "metric": METRIC, "value": v, "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{method} omega={omega} on a random disk mesh "
                               f"[reference arm timed on a bounded sample of {n} vertices]",
                   "method": method, "omega": omega},
        "cpu_baseline": {"value": v, "unit": METRIC, "cores": workers, "kind": "port",
                         "sample": sample, "host_cores_available": os.cpu_count(),
                         "note": "reference source absent (README only, licence-gated): the "
                                 "committed numpy oracle is the stated CPU baseline; the "
                                 "reference's loop is one process (numpy gather/bincount), so "
                                 "the host cores are filled with independent replicas"},
        "e2e": {"value": v, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm
def k1_profile(n, method):
    """What the committed ncu capture of K1 says for this workload (profiles/k1_traffic.json):
    DRAM bytes and fp64 warp instructions per launch."""
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("n_vertices") == n and tj.get("method") == method:
            return tj
    return {}


def roofline_of(k1_ms, b_alg, n, method, extra):
    peak, peak_src = measured_peak()
    achieved = b_alg / (k1_ms * 1e-3) / 1e9
    prof = k1_profile(n, method)
    out = {
        "kernel": "k_step_ring (fused smoothing step: one thread per vertex, star evaluated as "
                  "a chain of spokes, fused Delaunay pre-check)",
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": prof.get("dram_bytes_per_launch"),
        "peak_source": peak_src, "algorithmic_bytes_per_launch": b_alg, "kernel_ms": k1_ms,
    }
    fp64 = prof.get("fp64_warp_instructions_per_launch")
    if fp64:
        # second roofline: the fp64 pipe issues one warp instruction every 2 cycles per SM
        # sub-partition (64 DFMA/clk/SM): 148 SMs x 4 x 0.5 x 1.965 GHz
        rate = 148 * 4 * 0.5 * 1.965e9
        floor_ms = fp64 / rate * 1e3
        out["fp64_co_roofline"] = {
            "fp64_warp_instructions_per_launch": fp64,
            "peak_warp_instructions_per_s": rate,
            "floor_ms": floor_ms,
            "frac_of_fp64_peak": floor_ms / k1_ms,
            "hbm_frac_at_fp64_peak": b_alg / (floor_ms * 1e-3) / 1e9 / peak,
            "note": "the step is fp64-pipe / HBM co-bound (SURVEY.md hard part 1): with this "
                    "many fp64 instructions per launch the kernel cannot exceed "
                    "hbm_frac_at_fp64_peak of the HBM roofline, whatever the memory system does",
        }
    out.update(extra)
    return out


def config5_block(args, torch, dist, ob, world, rank, local, stream):
    """BASELINE.json configs[4]: Lloyd omega=2.0 on a FIXED 100M-vertex disk mesh at N GPUs
    (strong scaling: the reader divides the N=1 time by N x this time).  Same measurement
    rules as the main line; every rank builds the same mesh on its device."""
    from optimesh_b200 import generators as G

    grid, rounds, warm, steps = args.config5_grid, 40, 5, 10
    t0 = time.perf_counter()
    dm = G.disk_gpu(grid, rounds, 0, device=local, stream=stream)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    n, c = dm.n, dm.c
    dm.set_method("lloyd", 2.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world == 1:
        dm.run_prepare()
        dm.run(0.0, warm)
        torch.cuda.synchronize()
        e0.record()
        dm.run(0.0, steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        tot = dm.run_totals()
        flips = tot["n_flips"]
    else:
        from optimesh_b200.dist import SharedMesh

        sm = SharedMesh.from_complete(dm)
        dm.close()
        ob._lib.load().om_release_cached_memory(local)
        dm = sm
        dm.run_prepare()
        dm.run(0.0, warm)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        dm.run(0.0, steps)
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) * 1e3
        flips = dm.run_totals()["n_flips"]
        torch.cuda.synchronize()
        dist.barrier()
    free, total = torch.cuda.mem_get_info()
    dm.close()
    return {
        "workload": f"lloyd omega=2.0 on disk_gpu({grid}, rounds={rounds}): {n} vertices / {c} "
                    f"cells, FIXED size at every N (strong scaling)",
        "n_gpus": world, "n_vertices": n, "steps": steps, "warmup": warm,
        "ms_per_step": ms / steps, "value": n * steps / (ms * 1e-3), "unit": METRIC,
        "flips_per_step": flips / steps, "mesh_generation_s": t_gen,
        "device_memory_in_use_gb_this_rank": (total - free) / 1e9,
        "strong_scaling": "efficiency at N GPUs = ms_per_step(N=1) / (N x ms_per_step(N))",
    }


def run_single(args, torch, ob, local, stream):
    """N = 1: BASELINE.json configs[1] through the public loop (DeviceMesh.run = om_run: one
    CUDA graph holds update + fused Delaunay check + flip rounds + fix-up for all K steps)."""
    from optimesh_b200 import generators as G

    method, omega, grid = workload(args)
    t_gen = time.perf_counter()
    if args.rounds > 0:
        dm = G.disk_gpu(grid, args.rounds, 0, device=local, stream=stream)
        mesh_name = (f"random disk mesh disk_gpu({grid}, rounds={args.rounds}): mapped grid + "
                     f"{args.rounds} random-walk rounds with flips on the device")
    else:
        tp, tc = G.disk_mapped_grid_torch(grid, 0.25, 0, device=f"cuda:{local}")
        dm = ob.DeviceMesh.from_torch(tp, tc, stream=stream)
        del tp, tc
        mesh_name = f"jittered mapped grid disk_mapped_grid({grid})"
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    n, d, c = dm.n, dm.dim, dm.c
    # host copy of the workload (input of the end-to-end calls), taken before any smoothing
    pts = dm.points
    cells64 = dm.cells(np.int64)
    val = np.bincount(cells64.reshape(-1), minlength=n)
    bnd = dm.is_boundary_point
    high_valence = float((val[~bnd] > 8).mean())
    dm.set_method(method, omega)
    dm.run_prepare()  # graph build stays out of every timed region

    def timed_run(k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        steps, last = dm.run(0.0, k)
        e1.record()
        torch.cuda.synchronize()
        assert steps == k
        return e0.elapsed_time(e1), dm.run_totals()

    # early phase: the first steps on the fresh random mesh (most vertices limited: exact
    # limiter variant; many flips).  Also serves as warm-up.
    early_steps = 5
    ms_early, tot_early = timed_run(early_steps)
    if args.warmup > 0:
        timed_run(args.warmup)
    sampler = ClockSampler(local)
    torch.cuda.synchronize()
    sampler.start()
    l0 = dm.launch_count
    torch.cuda.profiler.start()  # ncu --profile-from-start off: only the timed region
    ms, tot = timed_run(args.steps)
    torch.cuda.profiler.stop()
    clocks = sampler.stop()
    launches = dm.launch_count - l0
    value = n * args.steps / (ms * 1e-3)

    # K1 with CUDA events around every launch: the same loop driven from the stream
    # (om_set_timing), over the steps that follow the timed region
    dm.set_timing(True)
    dm.run(0.0, args.steps)
    tim = dm.timing()
    phases = dm.phase_timing()
    dm.set_timing(False)
    k1_ms = tim["step_kernel_ms"] / max(tim["step_kernel_launches"], 1)
    b_alg = alg_bytes(n, c, d)
    roofline = roofline_of(k1_ms, b_alg, n, method, {
        "kernel_share_of_step": k1_ms / (ms / args.steps),
        "kernel_timing": f"CUDA events around each of {tim['step_kernel_launches']} launches of "
                         f"the ring kernel (the limiter variant the loop selects) in a second "
                         f"run of {args.steps} steps that follows the timed one: the same loop, "
                         f"its kernels launched from the stream instead of as one CUDA graph",
        "rest_of_step_ms": ms / args.steps - k1_ms,
        "phases_ms_stream_loop": phases,
        "note": "rest of the step = k_post + flag-driven Delaunay check + flip rounds + ring "
                "rows + recomputation of touched vertices + statistics",
    })

    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": f"{method} omega={omega} on a {mesh_name}: {n} vertices / {c} cells, "
                        f"fp64, step = point update + limiter + flip-until-Delaunay",
            "method": method, "omega": omega, "n_vertices": n, "n_cells": c,
            "interior_vertices_with_more_than_8_cells": high_valence,
            "mesh_generation_s": t_gen,
            "parallelism": "1 GPU",
            "l2": "inputs (points + ring rows = %.0f MB) larger than the 126 MB L2"
                  % ((16 * n + 32 * n) / 1e6),
        },
        "steps_per_s": args.steps / (ms * 1e-3),
        "flips_per_step": tot["n_flips"] / args.steps,
        "flip_rounds_per_step": tot["n_flip_rounds"] / args.steps,
        "limited_vertices_per_step": tot["n_limited"] / args.steps,
        "deferred_vertices_per_step": tot["n_deferred"] / args.steps,
        "early_phase": {
            "steps": early_steps, "ms_per_step": ms_early / early_steps,
            "value": n * early_steps / (ms_early * 1e-3),
            "flips_per_step": tot_early["n_flips"] / early_steps,
            "flip_rounds_per_step": tot_early["n_flip_rounds"] / early_steps,
            "limited_vertices_per_step": tot_early["n_limited"] / early_steps,
            "deferred_vertices_per_step": tot_early["n_deferred"] / early_steps,
            "note": "steps 1-5 on the fresh random mesh (exact-limiter kernel variant)",
        },
        "timed_steps": f"steps {early_steps + args.warmup + 1}-"
                       f"{early_steps + args.warmup + args.steps} of the run",
        "roofline": roofline,
        "clocks": clocks,
        "gpu_launches": launches,
    }
    dm.close()  # its device memory returns to the pool before the end-to-end calls
    if not args.no_e2e:
        # end to end through the public API with HOST buffers: upload, setup, K steps,
        # download -- all inside the timed region.  Two warm-up calls, then five timed; the
        # median is reported.
        e2e_steps = args.steps
        times = []
        p_out = c_out = None
        for call in range(7):
            del p_out, c_out
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            p_out, c_out = ob.optimize_points_cells(pts, cells64, method, 0.0, e2e_steps,
                                                    omega=omega, device=local)
            torch.cuda.synchronize()
            if call >= 2:  # two untimed warm-up calls (staging buffers, result cache, graph)
                times.append(time.perf_counter() - t0)
        dt = float(np.median(times))
        line["e2e"] = {
            "value": n * e2e_steps / dt, "unit": METRIC,
            "h2d_bytes_per_step": (pts.nbytes + cells64.nbytes) / e2e_steps,
            "d2h_bytes_per_step": (p_out.nbytes + c_out.nbytes) / e2e_steps,
            "call": f"optimize_points_cells(points, cells, {method!r}, 0.0, {e2e_steps}, "
                    f"omega={omega}) on host numpy arrays (float64 points, int64 cells)",
            "seconds": dt, "seconds_all_calls": times, "steps": e2e_steps,
        }
    if not args.no_config5:
        ob._lib.load().om_release_cached_memory(local)
        line["config5"] = config5_block(args, torch, None, ob, 1, 0, local, stream)
        ob._lib.load().om_release_cached_memory(local)
    if not args.no_cpu_baseline:
        v, dt, ns, cs, nb, workers = cpu_step_rate(method, omega, 250000, 2, 0)
        line["cpu_baseline"] = {
            "value": v, "unit": METRIC, "cores": workers, "kind": "port",
            "sample": f"{workers} replicas (one per host core) of a random disk mesh disk({nb}) "
                      f"(Qhull): {ns} vertices / {cs} cells each, 2 steps of {method} incl. "
                      f"limiter and flip-until-Delaunay ({dt:.1f} s)",
            "host_cores_available": os.cpu_count(),
        }
    print(json.dumps(line), flush=True)


def run_multi(args, torch, dist, ob, world, rank, local, stream):
    """N > 1: ONE mesh of N x (per-GPU size) vertices in one address space over the GPUs
    (dist.SharedMesh, csrc/shared.cu): chunk r of every array is resident on GPU r, all chunks
    are mapped on every rank, each rank runs the single-GPU pipeline on its own vertex range
    and work lists, what crosses a chunk boundary is peer memory over NVLink, the ranks meet
    on the device.  The same K steps through the same public loop as at N = 1."""
    from optimesh_b200 import generators as G
    from optimesh_b200.dist import SharedMesh

    method, omega, grid = workload(args)
    total_grid = int(round(grid * np.sqrt(world)))
    t_gen = time.perf_counter()
    full = G.disk_gpu(total_grid, args.rounds, 0, device=local) if args.rounds > 0 else None
    if full is None:
        tp, tc = G.disk_mapped_grid_torch(total_grid, 0.25, 0, device=f"cuda:{local}")
        full = ob.DeviceMesh.from_torch(tp, tc)
        del tp, tc
    n, d, c = full.n, full.dim, full.c
    full.set_method(method, omega)
    sm = SharedMesh.from_complete(full)  # (identical complete mesh on every rank -> chunks)
    full.close()
    ob._lib.load().om_release_cached_memory(local)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    sm.run_prepare()
    info = sm.info()

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    ext = torch.cuda.ExternalStream(sm.stream)  # the stream the library launches on

    def timed_run(k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        steps, last = sm.run(0.0, k)  # one graph launch; returns when this rank has finished
        e1.record(ext)
        e1.synchronize()
        assert steps == k
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), sm.run_totals()

    early_steps = 5
    ms_early, tot_early = timed_run(early_steps)
    if args.warmup > 0:
        timed_run(args.warmup)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = sm.launch_count
    ms, tot = timed_run(args.steps)
    clocks = sampler.stop()
    launches = sm.launch_count - l0
    cnt = torch.tensor([launches], dtype=torch.int64, device="cuda")
    dist.all_reduce(cnt)
    launches = int(cnt[0].item())
    value = n * args.steps / (ms * 1e-3)  # n = vertices of the whole mesh
    k1_ms = sm.time_update(5)
    n_own = info["vertex_hi"] - info["vertex_lo"]
    b_alg = alg_bytes(n_own, int(round(c * n_own / max(n, 1))), d)  # this rank's launch
    roofline = roofline_of(k1_ms, b_alg, n, method, {
        "kernel_share_of_step": k1_ms / (ms / args.steps),
        "kernel_timing": "CUDA events around 5 launches of rank 0's update kernel on its own "
                         "vertex range (om_shared_time_update), after the timed region",
    })
    barrier()
    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": f"{method} omega={omega} on a random disk mesh disk_gpu({total_grid}, "
                        f"rounds={args.rounds}): {n} vertices / {c} cells ({n // world} vertices "
                        f"per GPU), fp64, step = point update + limiter + flip-until-Delaunay",
            "method": method, "omega": omega, "n_vertices": n, "n_cells": c,
            "parallelism": f"{world} GPUs, one mesh in one address space: chunk r of every array "
                           f"resident on GPU r (CUDA VMM), all chunks mapped on every rank, "
                           f"cross-chunk accesses are NVLink peer memory, device-side meetings "
                           f"(barrier + all-reduce through peer atomics), one CUDA graph per rank",
            "chunk_vertices": info["chunk_vertices"],
            "resident_gb_rank0": info["resident_bytes"] / 1e9,
            "mesh_generation_and_distribution_s": t_gen,
            "l2": "inputs (points + ring rows = %.0f MB per GPU) larger than the 126 MB L2"
                  % (48 * n_own / 1e6),
            "timing": "CUDA events on the library's stream around the K-step call (one graph "
                      "launch per rank), host barrier before, max over ranks",
        },
        "steps_per_s": args.steps / (ms * 1e-3),
        "flips_per_step": tot["n_flips"] / args.steps,
        "flip_rounds_per_step": tot["n_flip_rounds"] / args.steps,
        "limited_vertices_per_step": tot["n_limited"] / args.steps,
        "early_phase": {"steps": early_steps, "ms_per_step": ms_early / early_steps,
                        "flips_per_step": tot_early["n_flips"] / early_steps},
        "roofline": roofline,
        "clocks": clocks,
        "gpu_launches": launches,
    }
    sm.close()
    if not args.no_e2e:
        # End to end through the public multi-GPU call with HOST buffers, on config 2's mesh as
        # written (10 M vertices, the same mesh at every N: the inputs of the weak mesh would be
        # N x 5 GB of host arrays per rank).  Every rank passes the same arrays, uploads them,
        # builds the complete mesh on its GPU, keeps its chunks, runs the loop and downloads the
        # result: set-up is replicated, so this number does not grow with N.
        from optimesh_b200.dist import optimize_points_cells_shared

        g1 = args.grid or DEFAULT_GRID
        dm0 = G.disk_gpu(g1, args.rounds, 0, device=local, stream=stream)
        pts, cells64 = dm0.points, dm0.cells(np.int64)
        dm0.close()
        times = []
        p_out = c_out = None
        for call in range(5):
            del p_out, c_out
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            p_out, c_out = optimize_points_cells_shared(pts, cells64, method, 0.0, args.steps,
                                                        omega=omega, device=local)
            torch.cuda.synchronize()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if call >= 2:
                times.append(float(t.item()))
        dt = float(np.median(times))
        line["e2e"] = {
            "value": pts.shape[0] * args.steps / dt, "unit": METRIC,
            "h2d_bytes_per_step": world * (pts.nbytes + cells64.nbytes) / args.steps,
            "d2h_bytes_per_step": world * (p_out.nbytes + c_out.nbytes) / args.steps,
            "call": f"optimize_points_cells_shared(points, cells, {method!r}, 0.0, {args.steps}, "
                    f"omega={omega}) on host numpy arrays, called by all {world} ranks",
            "workload": f"config 2 as written: disk_gpu({g1}, rounds={args.rounds}), "
                        f"{pts.shape[0]} vertices, the SAME mesh at every N (strong)",
            "seconds": dt, "seconds_all_calls": times, "steps": args.steps,
            "note": "two warm-up calls, median of three, max over ranks; upload and set-up are "
                    "replicated on every rank (DESIGN.md section 5), only the loop is divided",
        }
        del pts, cells64, p_out, c_out
    if not args.no_config5:
        ob._lib.load().om_release_cached_memory(local)
        line["config5"] = config5_block(args, torch, dist, ob, world, rank, local, stream)
    if rank == 0:
        print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import optimesh_b200 as ob
    from optimesh_b200.dist import torch_stream_handle

    stream = torch_stream_handle()
    if world == 1:
        run_single(args, torch, ob, local, stream)
    else:
        run_multi(args, torch, dist, ob, world, rank, local, stream)
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl != "reference" and args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # `python bench.py --gpus N` without a launcher: start the N ranks ourselves
        import subprocess

        port = 29500 + os.getpid() % 2000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    # the contract is ONE JSON line on stdout: keep a private copy of stdout for it and send
    # everything else (e.g. NCCL's version banner) to stderr
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
