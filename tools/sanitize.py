import sys
sys.path.insert(0, '.')
import numpy as np
import optimesh_b200 as ob
from optimesh_b200 import generators as G
pts, cells = G.disk_mapped_grid(60, 0.25, 0, shuffle=True)
for method, omega in (("lloyd", 2.0), ("cvt-block-diagonal", 1.0), ("odt-fixed-point", 1.0),
                      ("odt-dp-fp", 1.0)):
    p, c = ob.optimize_points_cells(pts, cells, method, 0.0, 4, omega=omega)
print("2d ok")
sp, sc = G.tetra_sphere(12)
p, c = ob.optimize_points_cells(sp, sc, "cpt-fixed-point", 0.0, 3, implicit_surface=ob.Sphere())
p, c = ob.optimize_points_cells(*G.square(20, 0.25, 0), "cpt-linear-solve", 1e-9, 2)
with ob.DeviceMesh(pts, cells) as dm:
    dm.stats(); dm.new_points()
print("all ok")
