"""Launch list of the EARLY phase (steps 1-5 on the fresh random mesh) for ncu:
OM_NO_GRAPH=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none
    --csv --log-file out.csv python tools/early_phase.py"""
import sys

sys.path.insert(0, ".")
import torch

from optimesh_b200 import generators as G

dm = G.disk_gpu(3154, 120, 0, device=0)
dm.set_method("cvt-block-diagonal", 1.0)
dm.flip_until_delaunay()
torch.cuda.synchronize()
torch.cuda.profiler.start()
k, st = dm.run(0.0, 5)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(k, st, dm.run_totals())
