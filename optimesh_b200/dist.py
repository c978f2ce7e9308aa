"""Smoothing one mesh on several GPUs of a box (one process per GPU, torch.distributed for
the rendezvous).

Round 2 -- what `bench.py --gpus N` and `optimize_points_cells_shared` run: `SharedMesh`
(bottom of this file; csrc/shared.cu).  ONE mesh in one address space: every array is cut into
N chunks by vertex / cell id, chunk r is memory of GPU r, all chunks are mapped on every rank
(CUDA virtual memory management over NVLink peer memory), so the single-GPU kernels run
unchanged on each rank's vertex range and what crosses a chunk boundary is a peer load, store or
atomic.  The topology is partitioned (memory per rank ~ 1/N), there is no collective and no host
readback on the data path; the ranks meet on the device (k_sync) and the whole loop is one CUDA
graph per rank.  Bit-identical to one GPU.  The host side here is small: the descriptor exchange
(`exchange_fds`) and the calls.

Round 1, kept behind its tests (`run_sharded`, `run_partitioned`): replicated topology, the
point update split by vertex range, coordinates exchanged with NCCL (a band around each range),
flip rounds exchanged as fixed-capacity record slots; every rank applies every flip.  Weak
scaling 0.67 / 0.59 / 0.49 at 2 / 4 / 8 GPUs against 0.84 / 0.82 / 0.81 for the shared address
space.
"""
from __future__ import annotations

import os
import time

import numpy as np

# wall-clock breakdown of sharded_flip, filled when OM_DIST_PROFILE is set (diagnostics)
PROFILE = {} if os.environ.get("OM_DIST_PROFILE") else None
# diagnostics of the last run_partitioned call on this rank
LAST_RUN = {}

from .mesh import DeviceMesh


def _is_solve_method(method: str) -> bool:
    """Methods whose update is a global solve: every rank runs them on the whole mesh."""
    from .mesh import normalize_method_name

    return normalize_method_name(method) in ("cpt-linear-solve", "cpt-quasi-newton")


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


def _tick():
    """Wall clock after a device sync -- only when OM_DIST_PROFILE is set (diagnostics)."""
    if PROFILE is None:
        return 0.0
    import torch

    torch.cuda.synchronize()
    return time.perf_counter()


class BandExchange:
    """Keeps only a BAND of foreign coordinates current on every rank.

    After the update of step k a rank needs, from the others, the vertices within a few edges
    of its own vertex range: the rings of its own vertices (next update) and the cells it
    examines in the flip check.  Every rank therefore publishes the part of ITS range that
    lies within `depth` edges of a foreign vertex (`om_band_build`, recomputed only after
    flips changed the topology); the index lists are all-gathered when they change, the
    coordinates every step.  All ranks hold the same topology and numbering, so ids mean the
    same thing everywhere.  A flip check that would read a coordinate outside own range +
    band reports `stale`; the caller then falls back to one full all-gather."""

    def __init__(self, dm: DeviceMesh, group=None, depth: int = 3):
        import torch.distributed as dist

        self.dm, self.group, self.depth = dm, group, depth
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.dirty = True
        self.counts = None
        self.idx_all = None
        self.full_gathers = 0
        self.band_bytes = 0
        self.slow_rounds = 0
        self.age, self.force, self.refresh = 0, False, 8
        # record-slot capacities: round 0 / later rounds.  They follow the largest count seen;
        # the first guess scales with the mesh (a fresh random mesh flags ~2 % of its cells)
        self.caps = [max(65536, _pow2_at_least(dm.c // 32)), max(4096, _pow2_at_least(dm.c // 128))]
        self.buffers = {}
        # the cells that sit on this rank's vertices (must be asked before any flip)
        lo, hi = owned_range(dm.n, self.rank, self.world)
        if self.rank == self.world - 1:
            hi = dm.n + 1  # the last range also takes what is left
        self.cell_range = dm.cell_range_of_vertices(lo, hi)

    def full_gather(self):
        import torch.distributed as dist

        n = self.dm.n
        chunk = chunk_of(n, self.world)
        x = device_points_tensor(self.dm)
        out = x[: self.world * chunk]
        send = out[self.rank * chunk:(self.rank + 1) * chunk].clone()
        dist.all_gather_into_tensor(out, send, group=self.group)
        self.dm.coords_all_valid()
        self.full_gathers += 1

    def _rebuild(self):
        import torch
        import torch.distributed as dist

        ptr, n = self.dm.band_build(self.depth)
        cnt = torch.zeros(self.world, dtype=torch.int64, device="cuda")
        cnt[self.rank] = n
        dist.all_reduce(cnt, group=self.group)
        self.counts = cnt.tolist()
        maxn = max(max(self.counts), 1)
        send = torch.zeros(maxn, dtype=torch.int32, device="cuda")
        if n > 0:
            send[:n] = torch.as_tensor(_DevPtr(ptr, (n,), "<i4"), device="cuda")
        self.idx_all = torch.empty(self.world * maxn, dtype=torch.int32, device="cuda")
        dist.all_gather_into_tensor(self.idx_all, send, group=self.group)
        self.idx_all = self.idx_all.view(self.world, maxn)
        self.maxn = maxn
        self.dirty = False

    def exchange(self):
        """Call after the update: foreign coordinates become stale, bands are refreshed."""
        import torch
        import torch.distributed as dist

        dm = self.dm
        dm.coords_invalidate()
        # The band only has to be a superset of what the others read; every read is
        # validated (stale -> full gather), so after flips it is rebuilt lazily: at once
        # after a fallback, otherwise every `refresh` steps.
        self.age += 1
        b0 = _tick()
        if self.counts is None or self.force or (self.dirty and self.age >= self.refresh):
            self._rebuild()
            self.age, self.force = 0, False
            if PROFILE is not None:
                PROFILE["band_rebuilds"] = PROFILE.get("band_rebuilds", 0) + 1
        b1 = _tick()
        stride = dm.points_device()[2]
        key = ("band", self.maxn, stride)
        if key not in self.buffers:
            self.buffers = {k: v for k, v in self.buffers.items() if k[0] != "band"}
            self.buffers[key] = (
                torch.zeros(self.maxn, stride, dtype=torch.float64, device="cuda"),
                torch.empty(self.world * self.maxn, stride, dtype=torch.float64, device="cuda"))
        send, recv = self.buffers[key]
        n = self.counts[self.rank]
        if n > 0:
            dm.band_pack(self.idx_all[self.rank].data_ptr(), n, send.data_ptr())
        dist.all_gather_into_tensor(recv, send, group=self.group)
        recv = recv.view(self.world, self.maxn, stride)
        for r in range(self.world):
            if r != self.rank and self.counts[r] > 0:
                dm.band_unpack(self.idx_all[r].data_ptr(), self.counts[r], recv[r].data_ptr())
        self.band_bytes += int(sum(self.counts)) * stride * 8
        b2 = _tick()
        if PROFILE is not None:
            PROFILE["band_rebuild"] = PROFILE.get("band_rebuild", 0.0) + (b1 - b0)
            PROFILE["band_xchg"] = PROFILE.get("band_xchg", 0.0) + (b2 - b1)


def _pow2_at_least(n: int) -> int:
    p = 1
    while p < n:
        p *= 2
    return p


def partitioned_flip(dm: DeviceMesh, band: BandExchange, group=None, tol: float = 0.0,
                     max_steps: int = 100):
    """flip-until-Delaunay with every check round split over the ranks by cell range.  Per
    round: each rank examines its cells (all of them in round 0, its share of the work list
    later); the flagged-edge records travel in fixed-capacity slots (one all-gather, no host
    readback between check and flips); every rank applies ALL records and runs select / flip /
    twin patch itself (integer work on identical topology).  A stale coordinate or a slot
    overflow rejects the round on every rank alike; it is then repeated the slow way."""
    import torch
    import torch.distributed as dist

    world, rank = band.world, band.rank
    clo, chi = band.cell_range
    dm.flip_pass_begin()
    first = True
    for rnd in range(max_steps + 1):
        cap = band.caps[0 if first else 1]
        key = (cap,)
        if key not in band.buffers:
            band.buffers[key] = (torch.zeros(cap + 1, 2, dtype=torch.float64, device="cuda"),
                                 torch.empty(world * (cap + 1), 2, dtype=torch.float64,
                                             device="cuda"))
        send, recv = band.buffers[key]
        r0 = _tick()
        dm.flip_round_check_nofetch(first, clo, chi, tol)
        dm.flip_round_pack(cap, send.data_ptr())
        r1 = _tick()
        dist.all_gather_into_tensor(recv, send, group=group)
        r2 = _tick()
        ncand, _, abort, maxc = dm.flip_round_apply_gathered(recv.data_ptr(), world, cap)
        r3 = _tick()
        if PROFILE is not None:
            tag = "r0" if first else "rN"
            for key2, dt in ((tag + "_check", r1 - r0), (tag + "_gather", r2 - r1),
                             (tag + "_apply", r3 - r2)):
                PROFILE[key2] = PROFILE.get(key2, 0.0) + dt
            PROFILE[tag + "_n"] = PROFILE.get(tag + "_n", 0) + 1
            PROFILE[tag + "_cap"] = cap
        if abort:
            band.slow_rounds += 1
            ncand = _slow_round(dm, band, first, clo, chi, tol, group, abort)
        # next capacity for this kind of round: twice the largest count seen (all ranks agree)
        band.caps[0 if first else 1] = max(1024, _pow2_at_least(2 * max(maxc, 1)))
        first = False
        if ncand == 0:
            break
    else:
        import warnings

        warnings.warn("Maximum number of edge flips reached.")
    nf, nr = dm.flip_pass_end()
    if nf > 0:
        band.dirty = True
    return nf, nr


def _slow_round(dm, band, first, clo, chi, tol, group, abort):
    """One round with host-side counts (exact buffer sizes); used after a rejected round."""
    import torch
    import torch.distributed as dist

    world, rank = band.world, band.rank
    if abort & 1:
        band.full_gather()  # a cell drifted outside own range + band
        band.dirty, band.force = True, True
    while True:
        ptr, n, stale = dm.flip_round_check(first, clo, chi, tol)
        info = torch.zeros(2 * world, dtype=torch.int64, device="cuda")
        info[2 * rank] = n
        info[2 * rank + 1] = int(stale)
        dist.all_reduce(info, group=group)
        info = info.tolist()
        counts, stales = info[0::2], info[1::2]
        if any(stales):
            band.full_gather()
            continue
        break
    total = sum(counts)
    if total == 0:
        return 0
    maxc = max(counts)
    send = torch.zeros(maxc, 2, dtype=torch.float64, device="cuda")
    if n > 0:
        send[:n] = torch.as_tensor(_DevPtr(ptr, (n, 2), "<f8"), device="cuda")
    out = torch.empty(world * maxc, 2, dtype=torch.float64, device="cuda")
    dist.all_gather_into_tensor(out, send, group=group)
    for r in range(world):
        if counts[r] > 0:
            dm.flip_add_records(out[r * maxc:].data_ptr(), counts[r])
    ncand, _ = dm.flip_round_apply(total)
    return ncand


def partitioned_step(dm: DeviceMesh, band: BandExchange, tol: float = 0.0, group=None):
    """One loop iteration with partitioned coordinates: update of the own vertex range (kept
    aside until every rank reports that it read no stale coordinate), statistics all-reduced,
    band exchange, round-wise sharded flip pass."""
    import torch
    import torch.distributed as dist

    t0 = _tick()
    while True:
        st = dm.update_points(tol)
        t1 = _tick()
        # one collective for max |diff|^2 (max), #limited and "stale" (sums)
        mine = torch.tensor([st["max_diff2"], float(st["n_limited"]), float(st["stale"])],
                            dtype=torch.float64, device="cuda")
        allr = torch.empty(band.world * 3, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allr, mine, group=group)
        allr = allr.view(band.world, 3).tolist()
        red = [max(a[0] for a in allr), sum(a[1] for a in allr), sum(a[2] for a in allr)]
        if red[2] > 0:  # a ring reached outside own range + band: refresh, repeat
            band.full_gather()
            band.dirty, band.force = True, True
            continue
        break
    dm.commit_points()
    t2 = _tick()
    band.exchange()
    t3 = _tick()
    nf, nr = partitioned_flip(dm, band, group, 0.0)
    t4 = _tick()
    if PROFILE is not None:
        for key, dt in (("update", t1 - t0), ("reduce+commit", t2 - t1), ("band", t3 - t2),
                        ("flips", t4 - t3)):
            PROFILE[key] = PROFILE.get(key, 0.0) + dt
        PROFILE["steps"] = PROFILE.get("steps", 0) + 1
    return dict(max_diff2=red[0], n_limited=int(red[1]), n_flips=nf, n_flip_rounds=nr)


def partitioned_begin(dm: DeviceMesh, group=None, band_depth: int = 3) -> BandExchange:
    """Sets the handle up for partitioned stepping (call right after om_create, before any
    flip) and runs the loop's initial flip pass."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = owned_range(dm.n, rank, world)
    band = BandExchange(dm, group, band_depth)
    dm.set_owned_range(lo, hi)
    dm.coords_all_valid()
    partitioned_flip(dm, band, group)
    dm.set_deferred_commit(True)
    return band


def partitioned_end(dm: DeviceMesh, band: BandExchange):
    dm.set_deferred_commit(False)
    band.full_gather()  # every rank holds the complete point array again
    dm.set_owned_range(0, -1)


def run_partitioned(dm: DeviceMesh, method: str, tol: float, max_num_steps: int,
                    omega: float = 1.0, group=None, log=None, band_depth: int = 3):
    """optimize() loop with partitioned coordinates.  Bit-identical to one GPU."""
    dm.set_method(method, omega)
    band = partitioned_begin(dm, group, band_depth)
    k = 0
    try:
        while True:
            k += 1
            st = partitioned_step(dm, band, tol, group)
            is_final = (st["max_diff2"] < tol * tol) or k >= max_num_steps
            if log is not None:
                log.append(dict(step=k, **st))
            if is_final:
                break
    finally:
        LAST_RUN.update(steps=k, fallback_full_gathers=band.full_gathers,
                        slow_rounds=band.slow_rounds,
                        band_vertices=int(sum(band.counts or [0])), band_bytes=band.band_bytes)
        partitioned_end(dm, band)
    return k, band

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
    replicated = _is_solve_method(method) or world == 1
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
                                  log=None, exchange: str = "band"):
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
        import torch.distributed as dist

        partition = (exchange == "band" and implicit_surface is None
                     and not _is_solve_method(method)
                     and dist.get_world_size(group) > 1)
        if partition:
            run_partitioned(dm, method, tol, max_num_steps, omega, group, log)
        else:
            run_sharded(GpuShard(dm, group), method, tol, max_num_steps, omega, group, log)
        return dm.points, dm.cells(cells.dtype)


def exchange_fds(mine, group=None):
    """All-to-all of file descriptors between the ranks of one box: every rank passes its list
    `mine` (same length everywhere) to every other rank over Unix datagram sockets with
    SCM_RIGHTS and gets {rank: [descriptors]} back (its own list under its own rank; the
    received descriptors are new ones in this process: the caller closes them).  The host side
    of SharedMesh: the descriptors are CUDA memory handles (cuMemExportToShareableHandle)."""
    import array
    import socket
    import struct

    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = [int(fd) for fd in mine]
    n = len(mine)
    tag = os.environ.get("MASTER_PORT", "0")
    path = lambda r: f"/tmp/om_shared_{tag}_{os.getuid()}_{r}"  # noqa: E731
    sock = socket.socket(socket.AF_UNIX, socket.SOCK_DGRAM)
    try:
        os.unlink(path(rank))
    except FileNotFoundError:
        pass
    sock.bind(path(rank))
    got = {rank: mine}
    try:
        dist.barrier(group)  # every socket is bound
        for r in range(world):
            if r != rank:
                # (socket.send_fds ignores its address argument in CPython 3.12)
                sock.sendmsg([struct.pack("i", rank)],
                             [(socket.SOL_SOCKET, socket.SCM_RIGHTS, array.array("i", mine))],
                             0, path(r))
        sock.settimeout(120.0)
        while len(got) < world:
            msg, rfds, _, _ = socket.recv_fds(sock, 16, max(n, 1))
            if len(rfds) != n:
                raise RuntimeError(f"expected {n} descriptors from a peer, got {len(rfds)}")
            got[struct.unpack("i", msg[:4])[0]] = list(rfds)
        dist.barrier(group)  # everybody has everything: the sockets may go
    finally:
        sock.close()
        try:
            os.unlink(path(rank))
        except FileNotFoundError:
            pass
    return got


# ---------------------------------------------------------------------------------------------
# One mesh in one address space over the GPUs of a box (csrc/shared.cu): memory per rank ~ 1/N,
# no halo buffers, no collective on the data path; the ranks meet on the device.
class SharedMesh(DeviceMesh):
    """A mesh whose arrays are cut into `world` chunks, chunk r resident on rank r's GPU and
    all of them mapped into one virtual range on every rank (CUDA virtual memory management
    over NVLink peer memory).  Built from a complete `DeviceMesh` that every rank holds (same
    inputs everywhere); `run` is the optimize() loop, called by all ranks alike."""

    @classmethod
    def from_complete(cls, full: DeviceMesh, group=None, flip_first: bool = True):
        import ctypes as C

        import torch.distributed as dist

        from . import _lib
        from ._lib import check

        lib = _lib.load()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        if flip_first:
            full.flip_until_delaunay()  # the loop expects a Delaunay mesh (same on every rank)
        h = C.c_void_p()
        fds = (C.c_int32 * 32)()
        n = C.c_int32()
        check(lib.om_shared_begin(full._h, rank, world, C.byref(h), fds, C.byref(n)))
        n = n.value
        mine = [int(fds[i]) for i in range(n)]
        got = exchange_fds(mine, group)
        try:
            flat = (C.c_int32 * (world * n))()
            for r in range(world):
                for i in range(n):
                    flat[r * n + i] = got[r][i]
            check(lib.om_shared_map(h, flat, n))
        finally:
            for r, lst in got.items():
                if r != rank:
                    for fd in lst:
                        os.close(fd)
        dist.barrier(group)  # every rank has copied its share: the complete handles may go
        self = cls.__new__(cls)
        self.n, self.dim, self.c = full.n, full.dim, full.c
        self.cells_dtype = full.cells_dtype
        self._h, self._lib, self._group = h, lib, group
        return self

    def run(self, tol: float, max_num_steps: int):
        import ctypes as C

        from . import _lib
        from ._lib import check

        steps = C.c_int64()
        st = _lib.StepStats()
        check(self._lib.om_shared_run(self._h, float(tol), int(max_num_steps), C.byref(steps),
                                      C.byref(st)))
        if st.flip_cap_hit:
            import warnings

            warnings.warn("Maximum number of edge flips reached.")
        return steps.value, st.as_dict()

    def run_prepare(self):
        from ._lib import check

        check(self._lib.om_shared_prepare(self._h))

    def time_update(self, reps: int = 5) -> float:
        """ms per launch of this rank's update kernel on its own vertex range (CUDA events)."""
        import ctypes as C

        from ._lib import check

        ms = C.c_double()
        check(self._lib.om_shared_time_update(self._h, int(reps), C.byref(ms)))
        return ms.value

    def info(self) -> dict:
        import ctypes as C

        from ._lib import check

        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        check(self._lib.om_shared_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(vertex_lo=a.value, vertex_hi=b.value, chunk_vertices=c.value,
                    resident_bytes=d.value)

    def close(self):
        """Collective: every rank must call it (the chunks are unmapped everywhere)."""
        if getattr(self, "_h", None) is not None and self._h:
            import torch
            import torch.distributed as dist

            torch.cuda.synchronize()
            if dist.is_initialized():
                dist.barrier(self._group)  # nobody reads this rank's chunks any more
            self._lib.om_destroy(self._h)
            self._h = None

    def __del__(self):  # no implicit collective at garbage collection
        pass


def optimize_points_cells_shared(points, cells, method: str, tol: float, max_num_steps: int,
                                 omega: float = 1.0, device=None, group=None):
    """`optimize_points_cells` for a process group with the mesh in the shared address space:
    every rank passes the same arrays and gets the same result back (README.md:124-126)."""
    import torch
    import torch.distributed as dist

    if device is None:
        device = torch.cuda.current_device()
    cells = np.asarray(cells)
    full = DeviceMesh(points, cells, device=device)
    full.set_method(method, omega)
    sm = SharedMesh.from_complete(full, group)
    full.close()
    try:
        sm.run(tol, max_num_steps)
        torch.cuda.synchronize()
        dist.barrier(group)  # every rank's part of the result is final
        return sm.points, sm.cells(cells.dtype)
    finally:
        sm.close()
