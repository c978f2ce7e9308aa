"""Small cases through every kernel family, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize.py
    compute-sanitizer --tool racecheck python tools/sanitize.py
    compute-sanitizer --tool initcheck python tools/sanitize.py
(OM_NO_GRAPH=1: the sanitizer does not follow conditional graph nodes.)"""
import sys

sys.path.insert(0, '.')
import numpy as np

import optimesh_b200 as ob
from optimesh_b200 import generators as G

pts, cells = G.disk(60, 1)  # random mesh: vertices without a ring row, masked cells, flips
for method, omega in (("lloyd", 2.0), ("cvt-block-diagonal", 1.0), ("odt-fixed-point", 1.0),
                      ("odt-dp-fp", 1.0), ("cpt-fixed-point", 1.0)):
    p, c = ob.optimize_points_cells(pts, cells, method, 0.0, 4, omega=omega)  # pipelined loop
    log = []
    p, c = ob.optimize_points_cells(pts, cells, method, 0.0, 3, omega=omega, log=log)  # stepwise
print("2d ok")
sp, sc = G.tetra_sphere(12)
p, c = ob.optimize_points_cells(sp, sc, "cpt-fixed-point", 0.0, 3, implicit_surface=ob.Sphere())
p, c = ob.optimize_points_cells(sp, sc, "odt-fixed-point", 0.0, 3)
p, c = ob.optimize_points_cells(*G.square(20, 0.25, 0), "cpt-linear-solve", 1e-9, 2)
p, c = ob.optimize_points_cells(*G.square(20, 0.25, 0), "cpt-quasi-newton", 1e-9, 2)
# the multigrid-preconditioned solve (above 20,000 vertices)
big, bc = G.square(150, 0.25, 0)
with ob.DeviceMesh(big, bc) as dm:
    print("mg", dm.solve_graph_laplacian(1e-10, 500))
# hooks on device memory
p, c = ob.optimize_points_cells(pts, cells, "lloyd", 0.0, 2, boundary_step=lambda x: x,
                                device_callables=True)
with ob.DeviceMesh(pts, cells) as dm:
    dm.stats()
    dm.new_points()
    dm.random_walk(2, 1, 0.3)
print("all ok")
