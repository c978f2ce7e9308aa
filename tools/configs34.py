import sys, time
sys.path.insert(0, '.')
import numpy as np
import optimesh_b200 as ob
from optimesh_b200 import generators as G

t=time.time(); pts, cells = G.square(2236, 0.25, 0); print("square gen", pts.shape, cells.shape, time.time()-t, flush=True)
with ob.DeviceMesh(pts, cells.astype(np.int32)) as dm:
    t=time.time(); nf = dm.flip_until_delaunay(); dm.synchronize(); print("initial flips", nf, time.time()-t, flush=True)
    dm.set_method("cpt-fixed-point")
    for _ in range(3): dm.step(0.0)
    t=time.time()
    for _ in range(20): st = dm.step(0.0)
    dm.synchronize(); dt=time.time()-t
    print("cpt-fixed-point 20 steps: %.3f ms/step, %.3g vu/s" % (dt/20*1e3, 20*len(pts)/dt), st, flush=True)
    for rtol in (1e-6, 1e-10):
        dm.points = pts
        t=time.time(); its, res = dm.solve_graph_laplacian(rtol, 200000); dm.synchronize(); dt=time.time()-t
        print(f"cpt-linear-solve rtol={rtol}: iters={its} relres={res:.3e} time={dt:.2f}s  {dt/its*1e3:.3f} ms/iter", flush=True)
        p = dm.points
        bnd = dm.is_boundary_point
        assert np.array_equal(p[bnd], pts[bnd])

t=time.time(); sp, sc = G.tetra_sphere(1000); print("sphere gen", sp.shape, sc.shape, time.time()-t, flush=True)
with ob.DeviceMesh(sp, sc.astype(np.int32)) as dm:
    dm.set_method("odt-fixed-point"); dm.set_sphere()
    nf = dm.flip_until_delaunay(); print("initial flips", nf, flush=True)
    for _ in range(3): dm.step(0.0)
    t=time.time()
    for _ in range(20): st = dm.step(0.0)
    dm.synchronize(); dt=time.time()-t
    print("odt sphere 20 steps: %.3f ms/step, %.3g vu/s" % (dt/20*1e3, 20*len(sp)/dt), st, flush=True)
    p = dm.points
    print("max |r-1|", np.abs(np.linalg.norm(p,axis=1)-1).max())
    ah, qh, s = dm.stats(); print(s)
