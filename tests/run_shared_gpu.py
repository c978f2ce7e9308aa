"""Multi-GPU check of the shared-address-space loop (csrc/shared.cu; run under torchrun on a box
with >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_shared_gpu.py

The mesh is cut into chunks of 2^21 vertices, so it takes a few million vertices to span two
GPUs.  The N-GPU result must equal the single-GPU pipelined loop bit for bit (same kernels,
same per-vertex summation order; only the GPU that executes a piece of work differs).
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    from optimesh_b200 import generators as G
    from optimesh_b200.dist import SharedMesh

    grid = int(os.environ.get("OM_SHARED_GRID", "2880"))
    ok = True
    for method, omega, steps in (("cvt-block-diagonal", 1.0, 7), ("lloyd", 2.0, 4),
                                 ("odt-fixed-point", 1.0, 3)):
        ref = G.disk_gpu(grid, 40, 0, device=local)
        ref.set_method(method, omega)
        t0 = time.perf_counter()
        k_ref, last_ref = ref.run(0.0, steps)
        torch.cuda.synchronize()
        t_ref = time.perf_counter() - t0
        tot_ref = ref.run_totals()
        p_ref, c_ref = ref.points, ref.cells()
        ref.close()

        full = G.disk_gpu(grid, 40, 0, device=local)
        full.set_method(method, omega)
        sm = SharedMesh.from_complete(full)
        full.close()
        info = sm.info()
        t0 = time.perf_counter()
        k, last = sm.run(0.0, steps)
        torch.cuda.synchronize()
        t_sh = time.perf_counter() - t0
        tot = sm.run_totals()
        dist.barrier()
        p, c = sm.points, sm.cells()
        # a second run on the same handle (graph reuse, buffer parity)
        k2, _ = sm.run(0.0, 2)
        torch.cuda.synchronize()
        dist.barrier()
        p2 = sm.points
        sm.close()
        same = (k == k_ref and np.array_equal(p, p_ref) and np.array_equal(c, c_ref)
                and tot["n_flips"] == tot_ref["n_flips"] and tot["n_limited"] == tot_ref["n_limited"]
                and last["max_diff2"] == last_ref["max_diff2"] and k2 == 2
                and np.isfinite(p2).all() and not np.array_equal(p2, p))
        flag = torch.tensor([int(same)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            nd = int((np.abs(p - p_ref).max(axis=1) > 0).sum())
            print(f"{method:20s} world={world} N={len(p)} steps={k} flips={tot['n_flips']} "
                  f"(ref {tot_ref['n_flips']}) bit-identical={bool(flag.item())} "
                  f"[points differing: {nd}, cells equal: {np.array_equal(c, c_ref)}] "
                  f"own range [{info['vertex_lo']}, {info['vertex_hi']}) resident "
                  f"{info['resident_bytes'] / 1e9:.2f} GB; {steps} steps: 1 GPU {t_ref * 1e3:.1f} ms, "
                  f"{world} GPUs {t_sh * 1e3:.1f} ms (first call: graph build included)", flush=True)
        ok = ok and bool(flag.item())
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)
    if rank == 0:
        print("SHARED OK", flush=True)


if __name__ == "__main__":
    main()
