"""Config 5, single-GPU leg: Lloyd omega=2 on a ~100M-vertex disk mesh generated on the device."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import optimesh_b200 as ob
from optimesh_b200 import generators as G

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
steps, warm = 10, 3
t0 = time.time()
pts, cells = G.disk_mapped_grid_torch(n)
torch.cuda.synchronize()
t1 = time.time()
dm = ob.DeviceMesh.from_torch(pts, cells, stream=1)
dm.synchronize()
t2 = time.time()
del pts, cells
torch.cuda.empty_cache()
dm.set_method("lloyd", 2.0)
nf, nr = dm.flip_until_delaunay()
out = dict(n_vertices=dm.n, n_cells=dm.c, generate_s=t1 - t0, setup_s=t2 - t1, initial_flips=int(nf))
log = []
for _ in range(warm):
    log.append(dm.step(0.0))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dm.set_timing(True)
torch.cuda.synchronize(); dm.synchronize()
tw = time.perf_counter()
for _ in range(steps):
    log.append(dm.step(0.0))
dm.synchronize()
ms = (time.perf_counter() - tw) * 1e3 / steps
tim = dm.timing()
out.update(ms_per_step=ms, vertex_updates_per_s=dm.n / (ms * 1e-3), k1_ms=tim["step_kernel_ms"] / max(tim["step_kernel_launches"], 1),
           flip_ms=tim["flip_pass_ms"] / max(tim["flip_passes"], 1), flips=[int(l["n_flips"]) for l in log],
           rounds=[int(l["n_flip_rounds"]) for l in log], limited=[int(l["n_limited"]) for l in log],
           mem_gb=torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9)
ah, qh, s = dm.stats()
out["q_avg"] = s["q_avg"]; out["q_min"] = s["q_min"]
# a full check must find nothing to flip right after a step
nf2, _ = dm.flip_until_delaunay()
out["flips_after_pass"] = int(nf2)
print(json.dumps(out))
