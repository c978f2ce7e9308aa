"""BASELINE configs 3 and 4 at full size on one B200 -> gpurun_out/<tag>_config3.json, _config4.json

    python tools/configs34.py [tag]

config 3: CPT on the 5M-vertex jittered square (cpt-fixed-point steps; one cpt-linear-solve to
1e-6 and 1e-10).  config 4: ODT fixed-point on the 2M-vertex sphere, projection every step.
Times are wall clock around synchronised calls (each call is milliseconds to seconds)."""
import json
import os
import sys
import time

sys.path.insert(0, '.')
import numpy as np

import optimesh_b200 as ob
from optimesh_b200 import generators as G

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
os.makedirs("gpurun_out", exist_ok=True)

pts, cells = G.square(2236, 0.25, 0)
c3 = {"workload": f"jittered square, {len(pts)} vertices / {len(cells)} cells, boundary pinned",
      "jacobi_only": os.environ.get("OM_PCG_JACOBI") is not None}
with ob.DeviceMesh(pts, cells.astype(np.int32)) as dm:
    dm.flip_until_delaunay()
    dm.set_method("cpt-fixed-point")
    dm.run(0.0, 3)
    dm.synchronize()
    t = time.perf_counter()
    k, st = dm.run(0.0, 20)
    dm.synchronize()
    dt = time.perf_counter() - t
    c3["cpt_fixed_point"] = {"steps": int(k), "ms_per_step": 1e3 * dt / k,
                            "vertex_updates_per_s": k * len(pts) / dt}
    solves = []
    for rtol in (1e-6, 1e-10):
        dm.points = pts
        dm.flip_until_delaunay()
        dm.synchronize()
        t = time.perf_counter()
        its, res = dm.solve_graph_laplacian(rtol, 200000)
        dm.synchronize()
        dt = time.perf_counter() - t
        solves.append({"rtol": rtol, "iterations": int(its), "relres": float(res),
                       "seconds": dt, "ms_per_iteration": 1e3 * dt / max(its, 1)})
        p = dm.points
        bnd = dm.is_boundary_point
        assert np.array_equal(p[bnd], pts[bnd])
    c3["cpt_linear_solve"] = solves
    # the loop with the solve as its update: 3 steps (flips change the matrix in between)
    dm.points = pts
    dm.set_method("cpt-linear-solve")
    dm.set_solver(1e-10, 200000)
    dm.synchronize()
    t = time.perf_counter()
    k, st = dm.run(0.0, 3)
    dm.synchronize()
    c3["cpt_linear_solve_loop"] = {"steps": int(k), "seconds": time.perf_counter() - t,
                                   "solver_iters_last_step": int(st.get("solver_iters", -1))}
print(json.dumps(c3))
with open(f"gpurun_out/{tag}_config3.json", "w") as f:
    json.dump(c3, f, indent=1)

sp, sc = G.tetra_sphere(1000)
c4 = {"workload": f"tetra-sphere, {len(sp)} vertices / {len(sc)} cells, odt-fixed-point, "
                  "projection onto the unit sphere every step"}
with ob.DeviceMesh(sp, sc.astype(np.int32)) as dm:
    dm.set_method("odt-fixed-point")
    dm.set_sphere()
    dm.flip_until_delaunay()
    for _ in range(3):
        dm.step(0.0)
    dm.synchronize()
    t = time.perf_counter()
    for _ in range(20):
        st = dm.step(0.0)
    dm.synchronize()
    dt = time.perf_counter() - t
    p = dm.points
    c4["odt_fixed_point"] = {"steps": 20, "ms_per_step": 1e3 * dt / 20,
                            "vertex_updates_per_s": 20 * len(sp) / dt,
                            "max_abs_radius_error": float(np.abs(np.linalg.norm(p, axis=1) - 1).max()),
                            "q_avg": float(dm.stats()[2]["q_avg"])}
print(json.dumps(c4))
with open(f"gpurun_out/{tag}_config4.json", "w") as f:
    json.dump(c4, f, indent=1)
