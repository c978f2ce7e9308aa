"""cpt-quasi-newton vs cpt-fixed-point on config 3 (5M-vertex jittered square): ms per step,
PCG iterations per step, steps to reach tol."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import optimesh_b200 as ob
from optimesh_b200 import generators as G

pts, cells = G.square(2236, 0.25, 0)
cells = cells.astype(np.int32)
for method in ("cpt-quasi-newton", "cpt-fixed-point"):
    with ob.DeviceMesh(pts, cells) as dm:
        dm.flip_until_delaunay()
        dm.set_method(method)
        for _ in range(2):
            dm.step(0.0)
        dm.synchronize()
        t = time.perf_counter()
        its = []
        for _ in range(10):
            st = dm.step(0.0)
            its.append(int(st["solver_iters"]))
        dm.synchronize()
        dt = (time.perf_counter() - t) / 10
        print(f"{method}: {dt * 1e3:.3f} ms/step, solver iterations {its}, "
              f"max_diff after 12 steps {st['max_diff2'] ** 0.5:.3e}", flush=True)
    with ob.DeviceMesh(pts, cells) as dm:
        dm.set_method(method)
        t = time.perf_counter()
        steps, last = dm.run(1.0e-5, 2000)
        dm.synchronize()
        print(f"{method}: {steps} steps to tol 1e-5 in {time.perf_counter() - t:.2f} s", flush=True)
