"""Where the end-to-end call spends its time (host arrays in, host arrays out).

    python tools/e2e_breakdown.py [grid|random]
"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import optimesh_b200 as ob
from optimesh_b200 import generators as G
kind = sys.argv[1] if len(sys.argv) > 1 else "random"
if kind == "grid":
    pts, cells = G.disk_mapped_grid(3154, 0.25, 0)
else:
    dm0 = G.disk_gpu(3154, 120, 0)
    pts, cells = dm0.points, dm0.cells(np.int64)
    dm0.close()
def T(label, f):
    t=time.perf_counter(); r=f(); dt=time.perf_counter()-t; print(f"{label:34s} {dt*1e3:8.1f} ms", flush=True); return r
for rep in range(4):
    print("--- rep", rep)
    t0=time.perf_counter()
    dm = T("DeviceMesh (H2D+setup)", lambda: ob.DeviceMesh(pts, cells))
    T("set_method", lambda: dm.set_method("cvt-block-diagonal", 1.0))
    T("run_prepare (graph build)", dm.run_prepare)
    T("run 20 steps (incl. flip passes)", lambda: dm.run(0.0, 20))
    p = T("get points", lambda: dm.points)
    c = T("get cells", lambda: dm.cells(np.int64))
    T("close", dm.close)
    print("total", (time.perf_counter()-t0)*1e3)
    del p, c
    t0=time.perf_counter(); r = ob.optimize_points_cells(pts, cells, "cvt-block-diagonal", 0.0, 20); print("optimize_points_cells", (time.perf_counter()-t0)*1e3)
    del r
