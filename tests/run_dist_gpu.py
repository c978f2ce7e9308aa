"""Multi-GPU check (run under torchrun on a box with >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_dist_gpu.py

The sharded run (update split by vertex range, first flip round split by cell range, NCCL
all-gathers) must reproduce the single-GPU result bit for bit.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    import optimesh_b200 as ob
    from optimesh_b200 import generators as G
    from optimesh_b200.dist import optimize_points_cells_sharded

    ok = True
    for name, (pts, cells), method, omega, kw in (
        ("disk", G.disk(150, 5), "lloyd", 2.0, {}),
        ("grid", G.disk_mapped_grid(400, 0.25, 2, shuffle=True), "cvt-block-diagonal", 1.0, {}),
        ("sphere", G.tetra_sphere(40), "odt-fixed-point", 1.0, {"implicit_surface": ob.Sphere()}),
    ):
        log = []
        p, c = optimize_points_cells_sharded(pts, cells, method, 1e-9, 8, omega=omega, log=log,
                                             device=local, **kw)
        from optimesh_b200 import dist as _d

        diag = dict(_d.LAST_RUN)
        _d.LAST_RUN.clear()
        # the same with the replicated-coordinates exchange
        p2, c2 = optimize_points_cells_sharded(pts, cells, method, 1e-9, 8, omega=omega,
                                               device=local, exchange="allgather", **kw)
        if not (np.array_equal(p2, p) and np.array_equal(c2, c)):
            print(f"[rank {rank}] {name}: band and all-gather exchanges differ", flush=True)
            ok = False
        rlog = []
        rp, rc = ob.optimize_points_cells(pts, cells, method, 1e-9, 8, omega=omega, log=rlog,
                                          device=local, **kw)
        same = (np.array_equal(p, rp) and np.array_equal(c, rc)
                and [l["n_flips"] for l in log] == [l["n_flips"] for l in rlog]
                and [l["n_limited"] for l in log] == [l["n_limited"] for l in rlog])
        flag = torch.tensor([int(same)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"{name:7s} {method:20s} world={world} N={len(pts)} flips={sum(l['n_flips'] for l in log)} "
                  f"bit-identical={bool(flag.item())} {diag}", flush=True)
        ok = ok and bool(flag.item())
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)
    if rank == 0:
        print("DIST OK", flush=True)


if __name__ == "__main__":
    main()
