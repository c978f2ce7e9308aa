"""world_size-2 gloo test of the sharded loop's host logic (no GPU): ranges, the in-place
all-gather, the all-reduced convergence test.  The device mesh is replaced by a test double
backed by the CPU oracle; the product path itself is covered on GPUs by
tests/test_gpu_parity.py::test_sharded_simulation and `bench.py --gpus N`."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class OracleShard:
    """Same interface as optimesh_b200.dist.GpuShard, numpy/oracle underneath."""

    PAD = 1024

    def __init__(self, points, cells):
        import oracle

        self.oracle = oracle
        self.mesh = oracle.MeshTri(points, cells)
        self.n, self.d = points.shape
        self.buf = np.zeros((self.n + self.PAD, self.d))
        self.buf[: self.n] = points
        self.lo, self.hi = 0, self.n
        self.method, self.omega = None, 1.0

    def set_method(self, method, omega):
        self.method, self.omega = method, omega

    def set_owned_range(self, lo, hi):
        self.lo, self.hi = (0, self.n) if hi < 0 else (lo, hi)

    def flip_until_delaunay(self):
        self.mesh.points = self.buf[: self.n]
        return self.mesh.flip_until_delaunay()

    def update_points(self, tol):
        o = self.oracle
        self.mesh.points = self.buf[: self.n]
        X = self.mesh.points
        new = o.get_new_points(self.mesh, self.method)
        bnd = self.mesh.is_boundary_point
        new[bnd] = X[bnd]
        diff = self.omega * (new - X)
        diff2 = np.einsum("ij,ij->i", diff, diff)
        max_step = np.full(self.n, np.inf)
        np.minimum.at(max_step, self.mesh.cells("points").reshape(-1),
                      np.repeat(self.mesh.cell_inradius, 3))
        max_step *= 0.5
        length = np.sqrt(diff2)
        idx = length > max_step
        diff[idx] *= (max_step / np.where(idx, length, 1.0))[idx, None]
        s = slice(self.lo, self.hi)
        self.buf[s] = X[s] + diff[s]  # only the owned range moves
        return dict(max_diff2=float(diff2[s].max()) if self.hi > self.lo else 0.0,
                    n_limited=int(idx[s].sum()))

    def project(self):
        return 0

    def points_tensor(self):
        return torch.from_numpy(self.buf)

    def scalar_device(self):
        return "cpu"


def _worker(rank, world, port, pts, cells, method, omega, steps, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from optimesh_b200.dist import run_sharded

    shard = OracleShard(pts, cells)
    log = []
    k = run_sharded(shard, method, 1e-6, steps, omega, log=log)
    if rank == 0:
        np.savez(out, points=shard.buf[: shard.n], cells=shard.mesh.cells("points"), steps=k,
                 flips=[l["n_flips"] for l in log], limited=[l["n_limited"] for l in log])
    dist.destroy_process_group()


@pytest.mark.parametrize("method,omega", [("lloyd", 2.0), ("cvt-block-diagonal", 1.0),
                                          ("odt-dp-fp", 1.0), ("cpt-quasi-newton", 1.0)])
def test_sharded_loop_matches_single_process(tmp_path, method, omega):
    import oracle
    from optimesh_b200 import generators as G

    pts, cells = G.disk(40, 2)
    steps = 6
    olog = []
    rp, rc = oracle.optimize_points_cells(pts, cells, method, 1e-6, steps, omega=omega, log=olog)
    out = str(tmp_path / "r0.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, pts, cells, method, omega, steps, out), nprocs=2, join=True)
    z = np.load(out)
    assert int(z["steps"]) == len(olog)
    assert z["flips"].tolist() == [l["n_flips"] for l in olog]
    assert z["limited"].tolist() == [l["n_limited"] for l in olog]
    assert np.array_equal(z["cells"], rc)
    assert np.array_equal(z["points"], rp)  # same arithmetic, only the ownership differs


def test_ranges_cover_everything():
    from optimesh_b200.dist import chunk_of, owned_range

    for n in (0, 1, 5, 1000, 1383):
        for world in (1, 2, 3, 8):
            r = [owned_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert chunk_of(n, world) * world <= n + world


def _fd_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from optimesh_b200.dist import exchange_fds

    # three "memory handles" per rank: anonymous files that name their owner and slot
    mine = []
    for i in range(3):
        fd = os.memfd_create(f"om_test_{rank}_{i}")
        os.write(fd, f"rank {rank} slot {i}".encode())
        mine.append(fd)
    got = exchange_fds(mine, None)
    ok = sorted(got) == list(range(world)) and got[rank] == mine
    for r, fds in got.items():
        for i, fd in enumerate(fds):
            ok = ok and os.pread(fd, 64, 0) == f"rank {r} slot {i}".encode()
            if r != rank:
                ok = ok and fd not in mine  # a new descriptor of this process
                os.close(fd)
    for fd in mine:
        os.close(fd)
    with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
        f.write("1" if ok else "0")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_shared_mesh_descriptor_exchange(tmp_path, world):
    """Host side of the shared-address-space loop (dist.SharedMesh, what `bench.py --gpus N`
    runs): every rank hands the descriptors of its memory chunks to every other rank.  The
    device side needs GPUs: tests/run_shared_gpu.py (N ranks, bit-identity with one GPU) and
    tests/test_gpu_parity.py::test_shared_address_space_loop_world_of_one."""
    port = 29500 + ((os.getpid() + 7 * world) % 2000)
    mp.spawn(_fd_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"ok{r}").read_text() == "1"
