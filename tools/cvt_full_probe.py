"""cvt-full on the oracle side: is a Krylov solver with the d x d diagonal
blocks as preconditioner a replacement for the reference's sparse direct solve?

Builds the full Hessian of the midpoint-rule CVT energy (diagonal blocks as in
cvt-block-diagonal, off-diagonal block of edge (i,j): +1/2 sum_cells ce e(x)e), Dirichlet rows
for boundary vertices, and counts iterations.  Run: python tools/cvt_full_probe.py [nb ...]
"""
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, ".")
import oracle  # noqa: E402
from optimesh_b200 import generators as G  # noqa: E402


def full_hessian(mesh):
    X = mesh.points
    n, d = X.shape
    mask = np.any(mesh.ce_ratios < -0.5, axis=0)
    cv = mesh.get_control_volumes(cell_mask=mask)
    cen = mesh.get_control_volume_centroids(cell_mask=mask)
    idx = mesh.idx_hierarchy[:, :, ~mask]
    e = mesh.half_edge_coords[:, ~mask]
    ce = mesh.ce_ratios[:, ~mask]
    rows, cols, vals = [], [], []
    ar = np.arange(n)
    for k in range(d):
        rows.append(d * ar + k)
        cols.append(d * ar + k)
        vals.append(2 * cv)
    for k in range(3):
        m = -0.5 * ce[k][:, None, None] * np.einsum("ci,cj->cij", e[k], e[k])
        i0, i1 = idx[0, k], idx[1, k]
        for a in range(d):
            for b in range(d):
                for (r, c, sgn) in ((i0, i0, 1), (i1, i1, 1), (i0, i1, -1), (i1, i0, -1)):
                    rows.append(d * r + a)
                    cols.append(d * c + b)
                    vals.append(sgn * m[:, a, b])
    H = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(d * n, d * n)).tocsr()
    rhs = (-2 * (X - cen) * cv[:, None])
    dead = mesh.is_boundary_point | (cv == 0.0)
    free = np.repeat(~dead, d)
    keep = sp.diags(free.astype(float))
    H = keep @ H @ keep + sp.diags((~free).astype(float))
    rhs[dead] = 0.0
    return H.tocsr(), rhs.reshape(-1), dead


def block_precond(H, n, d):
    B = np.zeros((n, d, d))
    for a in range(d):
        for b in range(d):
            B[:, a, b] = H[np.arange(n) * d + a, np.arange(n) * d + b].A1 if hasattr(H, "A1") else \
                np.asarray(H[np.arange(n) * d + a, np.arange(n) * d + b]).ravel()
    Binv = np.linalg.inv(B)
    ev = np.linalg.eigvalsh(B)
    def apply(v):
        return np.einsum("nij,nj->ni", Binv, v.reshape(n, d)).reshape(-1)
    return spla.LinearOperator(H.shape, matvec=apply), ev


def count(solver, H, rhs, M, **kw):
    it = [0]
    def cb(*a):
        it[0] += 1
    t = time.time()
    x, info = solver(H, rhs, M=M, callback=cb, **kw)
    res = np.linalg.norm(H @ x - rhs) / np.linalg.norm(rhs)
    return it[0], info, res, time.time() - t


for nb in [int(a) for a in sys.argv[1:]] or [40, 120]:
    pts, cells = G.disk(nb, 0)
    mesh = oracle.MeshTri(pts, cells)
    mesh.flip_until_delaunay()
    for phase in ("fresh", "after 5 lloyd steps"):
        if phase != "fresh":
            for _ in range(5):
                oracle.driver.step(mesh, "lloyd", omega=1.0)
                mesh.flip_until_delaunay()
        H, rhs, dead = full_hessian(mesh)
        n, d = mesh.points.shape
        M, ev = block_precond(H, n, d)
        asym = abs(H - H.T).max()
        t = time.time()
        xd = spla.spsolve(H.tocsc(), rhs)
        td = time.time() - t
        k = min(6, H.shape[0] - 2)
        try:
            lo = spla.eigsh(H, k=k, sigma=0.0, which="LM", return_eigenvectors=False)
        except Exception as ex:  # noqa: BLE001
            lo = [repr(ex)]
        hi = spla.eigsh(H, k=1, which="LA", return_eigenvectors=False)
        print(f"disk({nb}) {phase}: n={n} asym={asym:.1e} block eig min={ev.min():.3e} "
              f"max={ev.max():.3e}; eig nearest 0: {np.sort(lo)}; largest {hi}; spsolve {td:.2f}s")
        for name, solver, kw in (("minres+blockdiag", spla.minres, dict(rtol=1e-10, maxiter=5000)),
                                 ("gmres(60)+blockdiag", spla.gmres,
                                  dict(rtol=1e-10, restart=60, maxiter=100,
                                       callback_type="pr_norm"))):
            if name.startswith("minres") and ev.min() <= 0:
                print("   ", name, "skipped: preconditioner not SPD")
                continue
            it, info, res, tt = count(solver, H, rhs, M, **kw)
            print(f"    {name}: {it} iterations info={info} relres={res:.2e} ({tt:.1f}s)")
