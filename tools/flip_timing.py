import sys, time
sys.path.insert(0, '.')
import numpy as np
import optimesh_b200 as ob
from optimesh_b200 import generators as G
pts, cells = G.disk_mapped_grid(3154, 0.25, 0)
dm = ob.DeviceMesh(pts, cells.astype(np.int32))
dm.set_method("cvt-block-diagonal", 1.0)
dm.flip_until_delaunay()
for _ in range(6): dm.step(0.0)
# (a) check-only pass (no flips needed): 1 full k_suspect + 1 readback
dm.synchronize()
t=time.perf_counter()
for _ in range(50): r = dm.flip_until_delaunay()
dm.synchronize(); print("check-only pass: %.3f ms" % ((time.perf_counter()-t)/50*1e3), r)
# (b) stats() call = 1 kernel + syncs; launch_count/no-op call latency
t=time.perf_counter()
for _ in range(200): dm.launch_count
print("ctypes call: %.4f ms" % ((time.perf_counter()-t)/200*1e3))
t=time.perf_counter()
for _ in range(200): dm.synchronize()
print("empty sync: %.4f ms" % ((time.perf_counter()-t)/200*1e3))
# (c) K1 alone (update_points) incl. its syncs
t=time.perf_counter()
for _ in range(20): dm.update_points(0.0)
dm.synchronize(); print("update_points: %.3f ms" % ((time.perf_counter()-t)/20*1e3))
# (d) full steps
t=time.perf_counter()
for _ in range(20): st=dm.step(0.0)
dm.synchronize(); print("step: %.3f ms" % ((time.perf_counter()-t)/20*1e3), st)
