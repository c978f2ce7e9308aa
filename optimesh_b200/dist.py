"""Smoothing one mesh on several GPUs of a box (one process per GPU, torch.distributed).

Round-1 decomposition ("replicated topology, sharded update"): every rank builds the SAME
device mesh (identical inputs give identical internal numbering, so vertex ranges mean the
same thing everywhere).  After the spatial renumbering a contiguous vertex range is a
compact region, so rank r updates the range [r*chunk, (r+1)*chunk) with the fused step
kernel -- the fp64-heavy half of a step -- and the updated coordinates are made visible
everywhere by ONE in-place all-gather over NVLink straight on the device point array
(`om_points_device`).  Convergence (max |diff|^2) and the limiter count are all-reduced.
The first round of the flip-until-Delaunay pass -- the only one that scans every cell -- is
split over the ranks by cell range; the flagged-edge records are all-gathered and applied by
every rank, and the remaining work-list rounds (flagged cells only) run replicated: same
data, same deterministic kernels -> same topology everywhere.  Results are bit-identical to
the single-GPU run.

What this does NOT do yet: shard the topology storage, and shrink
the coordinate exchange to the one-ring band.
"""
from __future__ import annotations

import os
import time

import numpy as np

# wall-clock breakdown of sharded_flip, filled when OM_DIST_PROFILE is set (diagnostics)
PROFILE = {} if os.environ.get("OM_DIST_PROFILE") else None

from .mesh import DeviceMesh


def chunk_of(n: int, world: int) -> int:
    return (n + world - 1) // world if n > 0 else 0


def owned_range(n: int, rank: int, world: int):
    c = chunk_of(n, world)
    lo = min(n, rank * c)
    return lo, min(n, lo + c)


class _DevPtr:
    """Exposes a raw device pointer through __cuda_array_interface__ (zero-copy)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
            "strides": None,
        }


def torch_stream_handle() -> int:
    """cudaStream_t of torch's current stream for om_create.  torch's default stream is the
    legacy default stream, whose handle is 0 -- which the C-ABI reads as "make a private
    stream" -- so it is passed as cudaStreamLegacy (1): the library's kernels must be ordered
    with torch's copies and NCCL collectives."""
    import torch

    return torch.cuda.current_stream().cuda_stream or 1


def device_points_tensor(dm: DeviceMesh):
    """torch view [n_alloc, stride] of the handle's internal point array."""
    import torch

    ptr, n_alloc, stride = dm.points_device()
    return torch.as_tensor(_DevPtr(ptr, (n_alloc, stride), "<f8"), device="cuda")


def sharded_flip(dm: DeviceMesh, group=None, tol: float = 0.0, max_steps: int = 100):
    """flip-until-Delaunay with the first round (the only one that scans every cell) split
    over the ranks: each rank examines its range of cells, the flagged-edge records
    (16 bytes each) are all-gathered, every rank applies all of them and runs the remaining
    work-list rounds itself.  Identical topology on every rank, identical to one GPU."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    clo, chi = owned_range(dm.c, rank, world)
    t0 = time.perf_counter() if PROFILE is not None else 0.0
    ptr, n = dm.flip_check_range(clo, chi, tol)
    t1 = time.perf_counter() if PROFILE is not None else 0.0
    counts = torch.zeros(world, dtype=torch.int64, device="cuda")
    counts[rank] = n
    dist.all_reduce(counts, group=group)
    counts = counts.tolist()
    maxc = max(counts)
    t2 = time.perf_counter() if PROFILE is not None else 0.0
    if maxc > 0:
        send = torch.zeros(maxc, 2, dtype=torch.float64, device="cuda")
        if n > 0:
            send[:n] = torch.as_tensor(_DevPtr(ptr, (n, 2), "<f8"), device="cuda")
        out = torch.empty(world * maxc, 2, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(out, send, group=group)
        for r in range(world):
            if counts[r] > 0:
                dm.flip_add_records(out[r * maxc:].data_ptr(), counts[r])
    if PROFILE is None:
        return dm.flip_finish(tol, max_steps)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    res = dm.flip_finish(tol, max_steps)
    t4 = time.perf_counter()
    for key, dt in (("check", t1 - t0), ("counts", t2 - t1), ("records", t3 - t2),
                    ("finish", t4 - t3)):
        PROFILE[key] = PROFILE.get(key, 0.0) + dt
    PROFILE["calls"] = PROFILE.get("calls", 0) + 1
    PROFILE["records_n"] = PROFILE.get("records_n", 0) + sum(counts)
    return res


class GpuShard:
    """Adapter: the operations `run_sharded` needs, on a DeviceMesh."""

    def __init__(self, dm: DeviceMesh, group=None):
        self.dm = dm
        self.n = dm.n
        self.group = group

    def set_method(self, method, omega):
        self.dm.set_method(method, omega)

    def set_owned_range(self, lo, hi):
        self.dm.set_owned_range(lo, hi)

    def flip_until_delaunay(self):
        import torch.distributed as dist

        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            return sharded_flip(self.dm, self.group)
        return self.dm.flip_until_delaunay()

    def update_points(self, tol):
        return self.dm.update_points(tol)

    def project(self):
        return self.dm.project()

    def points_tensor(self):
        return device_points_tensor(self.dm)

    def scalar_device(self):
        return "cuda"


def run_sharded(shard, method: str, tol: float, max_num_steps: int, omega: float = 1.0,
                group=None, log=None):
    """The optimize() loop with the point update sharded over the ranks of `group`.
    `shard` is a GpuShard (or a test double with the same methods).  Returns steps taken."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = shard.n
    chunk = chunk_of(n, world)
    lo, hi = owned_range(n, rank, world)
    replicated = "linear-solve" in method.lower().replace(" ", "-") or world == 1
    shard.set_method(method, omega)
    shard.flip_until_delaunay()
    if not replicated:
        shard.set_owned_range(lo, hi)
    k = 0
    try:
        while True:
            k += 1
            st = shard.update_points(tol)
            max_diff2, n_limited = st["max_diff2"], st["n_limited"]
            if not replicated:
                dev = shard.scalar_device()
                a = torch.tensor([max_diff2], dtype=torch.float64, device=dev)
                b = torch.tensor([n_limited], dtype=torch.int64, device=dev)
                dist.all_reduce(a, op=dist.ReduceOp.MAX, group=group)
                dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group)
                max_diff2, n_limited = float(a.item()), int(b.item())
                if chunk > 0:
                    x = shard.points_tensor()
                    out = x[: world * chunk]  # rank r's chunk is rows [r*chunk, (r+1)*chunk)
                    send = out[rank * chunk:(rank + 1) * chunk].clone()
                    dist.all_gather_into_tensor(out, send, group=group)
            shard.project()
            nf, nr = shard.flip_until_delaunay()
            is_final = (max_diff2 < tol * tol) or k >= max_num_steps
            if log is not None:
                log.append(dict(step=k, max_diff2=max_diff2, n_limited=n_limited, n_flips=nf,
                                n_flip_rounds=nr))
            if is_final:
                break
    finally:
        if not replicated:
            shard.set_owned_range(0, -1)
    return k


def optimize_points_cells_sharded(points, cells, method: str, tol: float, max_num_steps: int,
                                  omega: float = 1.0, implicit_surface=None,
                                  implicit_surface_tol: float = 1.0e-10, device=None, group=None,
                                  log=None):
    """`optimize_points_cells` for a process group: every rank passes the same arrays and
    gets the same result back (README.md:124-126 semantics)."""
    import torch

    from .surfaces import Sphere

    if device is None:
        device = torch.cuda.current_device()
    cells = np.asarray(cells)
    stream = torch_stream_handle()
    with DeviceMesh(points, cells, device=device, stream=stream) as dm:
        if implicit_surface is None:
            dm.clear_surface()
        elif isinstance(implicit_surface, Sphere):
            dm.set_sphere(implicit_surface.center, implicit_surface.radius, implicit_surface_tol)
        else:
            raise NotImplementedError("sharded runs support the built-in Sphere surface only")
        run_sharded(GpuShard(dm, group), method, tol, max_num_steps, omega, group, log)
        return dm.points, dm.cells(cells.dtype)
