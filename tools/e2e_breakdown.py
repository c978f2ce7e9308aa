import sys, time
sys.path.insert(0, '.')
import numpy as np
import optimesh_b200 as ob
from optimesh_b200 import generators as G
pts, cells = G.disk_mapped_grid(3154, 0.25, 0)
def T(label, f):
    t=time.perf_counter(); r=f(); dt=time.perf_counter()-t; print(f"{label:28s} {dt*1e3:8.1f} ms", flush=True); return r
for rep in range(3):
    print("--- rep", rep)
    t0=time.perf_counter()
    dm = T("DeviceMesh (H2D+setup)", lambda: ob.DeviceMesh(pts, cells))
    T("set_method", lambda: dm.set_method("cvt-block-diagonal", 1.0))
    T("run 20 steps (incl. init flips)", lambda: dm.run(0.0, 20))
    p = T("get points", lambda: dm.points)
    c = T("get cells", lambda: dm.cells(np.int64))
    T("close", dm.close)
    print("total", (time.perf_counter()-t0)*1e3)
    t0=time.perf_counter(); ob.optimize_points_cells(pts, cells, "cvt-block-diagonal", 0.0, 20); print("optimize_points_cells", (time.perf_counter()-t0)*1e3)
