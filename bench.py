#!/usr/bin/env python
"""Benchmark of the smoothing step (BASELINE.json metric: smoothing steps/s and
vertex-updates/s at ~10M vertices, achieved HBM GB/s against peak).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one iteration of the optimize() loop on the device: fused point update
(K1) + flip-until-Delaunay.  At N=1 the workload is BASELINE.json configs[1]: CVT
block-diagonal on a ~10M-vertex disk mesh, fp64.  Rank 0 prints ONE JSON line.
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
DEFAULT_GRID = 3154  # disk_mapped_grid(3154): 9,947,716 vertices, 19,882,818 cells


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--method", default=None)
    ap.add_argument("--omega", type=float, default=None)
    ap.add_argument("--grid", type=int, default=None, help="n of disk_mapped_grid(n) per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


def make_mesh(grid, seed):
    from optimesh_b200 import generators as G

    return G.disk_mapped_grid(grid, 0.25, seed)


# --------------------------------------------------------------------------- CPU arms
def cpu_step_rate(method, omega, grid, steps, warmup):
    """The oracle (numpy port of the reference's algorithm) timed on the host cores:
    same step (update + limiter + flip-until-Delaunay), bounded sample."""
    import oracle

    pts, cells = make_mesh(grid, 0)
    mesh = oracle.MeshTri(pts, cells)
    mesh.flip_until_delaunay()
    for _ in range(warmup):
        oracle.driver.step(mesh, method, omega=omega)
        mesh.flip_until_delaunay()
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.driver.step(mesh, method, omega=omega)
        mesh.flip_until_delaunay()
    dt = time.perf_counter() - t0
    n = pts.shape[0]
    return n * steps / dt, dt, n, cells.shape[0]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    method, omega, grid = workload(args)
    total = args.steps + args.warmup
    # size the sample so the whole run stays within a few minutes:
    # the numpy port costs ~25 us per vertex per step on one core
    budget_s = 150.0
    n_target = int(min(1.0e6, max(2.0e4, budget_s / (total * 25e-6))))
    sample_grid = min(grid, int(np.sqrt(n_target)))
    v, dt, n, c = cpu_step_rate(method, omega, sample_grid, args.steps, args.warmup)
    sample = (f"disk_mapped_grid({sample_grid}): {n} vertices / {c} cells, {args.steps} steps of "
              f"{method} (omega={omega}) incl. limiter and flip-until-Delaunay")
    line = {
        "impl": "reference",
        "metric": METRIC, "value": v, "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{method} omega={omega} on disk_mapped_grid({grid}) "
                               f"[reference arm timed on a bounded sample]",
                   "method": method, "omega": omega},
        "cpu_baseline": {"value": v, "unit": METRIC, "cores": 1, "kind": "port", "sample": sample,
                         "host_cores_available": os.cpu_count(),
                         "note": "reference source absent (README only, licence-gated): the "
                                 "committed numpy oracle is the stated CPU baseline; numpy "
                                 "gather/bincount is single-threaded"},
        "e2e": {"value": v, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm
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

    method, omega, grid = workload(args)
    from optimesh_b200.dist import torch_stream_handle

    stream = torch_stream_handle()
    if world == 1:
        pts, cells = make_mesh(grid, 0)
        n, d = pts.shape
        c = cells.shape[0]
        dm = ob.DeviceMesh(pts, cells.astype(np.int32), device=local, stream=stream)
        total_grid = grid
    else:
        # one mesh of world x (per-GPU size) vertices, built on every GPU from the same seed
        from optimesh_b200 import generators as G

        total_grid = int(round(grid * np.sqrt(world)))
        tp, tc = G.disk_mapped_grid_torch(total_grid, 0.25, 0, device=f"cuda:{local}")
        n, d = int(tp.shape[0]), int(tp.shape[1])
        c = int(tc.shape[0])
        dm = ob.DeviceMesh.from_torch(tp, tc, stream=stream)
        del tp, tc
        torch.cuda.empty_cache()
    dm.set_method(method, omega)
    band = None
    if world > 1:
        from optimesh_b200.dist import owned_range, partitioned_begin, partitioned_step

        lo, hi = owned_range(n, rank, world)
        band = partitioned_begin(dm)  # own ranges + the loop's initial flip pass (untimed)
    else:
        dm.flip_until_delaunay()  # the loop's initial flip pass (setup, untimed)

    def one_step():
        """One loop iteration; at world > 1 (dist.run_partitioned): own vertex range updated,
        statistics all-reduced, band of coordinates exchanged over NCCL, every flip-check
        round split by cell range with its flagged-edge records all-gathered."""
        if world == 1:
            return dm.step(0.0)
        return partitioned_step(dm, band, 0.0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()
    dm.set_timing(True)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = dm.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flips = 0
    rounds = 0
    limited = 0
    e0.record()
    for _ in range(args.steps):
        st = one_step()
        flips += st["n_flips"]
        rounds += st["n_flip_rounds"]
        limited += st["n_limited"]
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = dm.launch_count - l0
    tim = dm.timing()
    dm.set_timing(False)
    if band is not None:
        from optimesh_b200 import dist as _d

        if _d.PROFILE:
            steps_p = max(_d.PROFILE.get("steps", 1), 1)
            print(f"[rank {rank}] ms/step: " + " ".join(
                f"{k}={1e3 * v / steps_p:.3f}" if isinstance(v, float) else f"{k}={v}"
                for k, v in _d.PROFILE.items()), file=sys.stderr)
        line_band = {"band_vertices_all_ranks": int(sum(band.counts or [0])),
                     "fallback_full_gathers": band.full_gathers,
                     "slow_flip_rounds": band.slow_rounds}
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        cnt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(cnt)
        launches = int(cnt[0].item())
    value = n * args.steps / (ms * 1e-3)  # n = vertices of the whole (sharded) mesh

    # roofline of the dominant kernel (K1, fused step)
    peak, peak_src = measured_peak()
    k1_ms = tim["step_kernel_ms"] / max(tim["step_kernel_launches"], 1)
    n_own = n if world == 1 else (hi - lo)
    b_alg = alg_bytes(n_own, int(round(c * n_own / max(n, 1))), d)  # this rank's launch
    achieved = b_alg / (k1_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("n_vertices") == n and tj.get("method") == method:
            traffic = tj.get("dram_bytes_per_launch")
    roofline = {
        "kernel": "k_step (fused smoothing step, one thread per vertex)",
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": b_alg, "kernel_ms": k1_ms,
        "kernel_share_of_step": tim["step_kernel_ms"] / ms,
        "flip_pass_ms": tim["flip_pass_ms"] / max(tim["flip_passes"], 1),
        "note": "bound by instruction issue (70 % of issue slots, fp64 pipe 47 %): DESIGN.md section 4",
    }

    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": f"{method} omega={omega}, disk_mapped_grid({total_grid}): {n} vertices / "
                        f"{c} cells ({n // world} vertices per GPU), fp64, step = point update + "
                        f"limiter + flip-until-Delaunay",
            "method": method, "omega": omega, "n_vertices": n, "n_cells": c,
            "parallelism": "1 GPU" if world == 1 else
            f"{world} GPUs: vertex ranges of one mesh (topology replicated), update and every "
            f"flip-check round sharded, band of coordinates + flagged-edge records exchanged "
            f"over NCCL each step",
            "l2": "inputs (points+cells+twins = %.0f MB) larger than the 126 MB L2"
                  % ((16 * n + 32 * c) / 1e6),
        },
        "steps_per_s": args.steps / (ms * 1e-3),
        "flips_in_timed_region": flips, "flip_rounds_in_timed_region": rounds,
        "limited_vertex_steps": limited,
        "roofline": roofline,
        "clocks": clocks,
        "gpu_launches": launches,
    }
    if band is not None:
        line["exchange"] = line_band

    dm.close()  # its device memory returns to the pool before the end-to-end calls
    if rank == 0 and world == 1 and not args.no_e2e:
        # end to end through the public API with HOST buffers: upload, setup, K steps,
        # download -- all inside the timed region.  Five calls, the median is reported
        # (a call that has to grow the driver's memory pool is several times slower, and
        # host-side page faulting of the fresh 637 MB result arrays jitters).
        e2e_steps = args.steps
        cells64 = cells  # int64, as numpy produces it
        times = []
        p_out = c_out = None
        for _ in range(5):
            del p_out, c_out
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            p_out, c_out = ob.optimize_points_cells(pts, cells64, method, 0.0, e2e_steps,
                                                    omega=omega, device=local)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        dt = float(np.median(times))
        line["e2e"] = {
            "value": n * e2e_steps / dt, "unit": METRIC,
            "h2d_bytes_per_step": (pts.nbytes + cells64.nbytes) / e2e_steps,
            "d2h_bytes_per_step": (p_out.nbytes + c_out.nbytes) / e2e_steps,
            "call": f"optimize_points_cells(points, cells, {method!r}, 0.0, {e2e_steps}, "
                    f"omega={omega}) on host numpy arrays",
            "seconds": dt, "seconds_all_calls": times, "steps": e2e_steps,
        }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample_grid = 600  # 360,000 vertices: ~10-30 s of single-core numpy
        v, dt, ns, cs = cpu_step_rate(method, omega, sample_grid, 2, 0)
        line["cpu_baseline"] = {
            "value": v, "unit": METRIC, "cores": 1, "kind": "port",
            "sample": f"disk_mapped_grid({sample_grid}): {ns} vertices / {cs} cells, 2 steps of "
                      f"{method} incl. limiter and flip-until-Delaunay ({dt:.1f} s)",
            "host_cores_available": os.cpu_count(),
        }
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    # the contract is ONE JSON line on stdout: keep a private copy of stdout for it and send
    # everything else (e.g. NCCL's version banner) to stderr
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
