"""cvt-full on the oracle side, part 2: the optimize() loop with the EXACT sparse solve of the
recollected full Hessian (tools/cvt_full_probe.py) against lloyd and cvt-block-diagonal.
Run from the repo root: python tools/cvt_full_loop_probe.py"""
import sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import oracle
from optimesh_b200 import generators as G
from cvt_full_probe import full_hessian

def stats(mesh):
    q = mesh.q_radius_ratio
    return q.mean(), q.min()

for nb in (40, 120):
    pts, cells = G.disk(nb, 0)
    for method in ("lloyd", "cvt-block-diagonal", "cvt-full"):
        mesh = oracle.MeshTri(pts.copy(), cells.copy()); mesh.flip_until_delaunay()
        hist = []
        for k in range(30):
            X = mesh.points
            if method == "cvt-full":
                H, rhs, dead = full_hessian(mesh)
                d = spla.spsolve(H.tocsc(), rhs).reshape(X.shape)
                new = X + d
            else:
                new = oracle.get_new_points(mesh, method)
            bnd = mesh.is_boundary_point
            new[bnd] = X[bnd]
            diff = new - X
            md = np.sqrt((diff**2).sum(1)).max()
            lim = np.full(len(X), np.inf)
            np.minimum.at(lim, mesh.cells("points").reshape(-1), np.repeat(mesh.cell_inradius, 3))
            lim *= 0.5
            L = np.sqrt((diff**2).sum(1)); idx = L > lim
            diff[idx] *= (lim[idx] / L[idx])[:, None]
            mesh.points = X + diff
            mesh.flip_until_delaunay()
            hist.append((md, idx.sum()))
            if md < 1e-5: break
        qa, qm = stats(mesh)
        print(f"disk({nb}) {method:20s} steps {len(hist):3d} last max|diff| {hist[-1][0]:.2e} limited(last) {hist[-1][1]:4d} q_avg {qa:.4f} q_min {qm:.3f}  |diff| trace", " ".join(f"{h[0]:.1e}" for h in hist[:12]))
