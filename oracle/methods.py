"""Per-method ``get_new_points`` (oracle, test infrastructure).

Method names: /root/reference/README.md:80 (lloyd, cvt-block-diagonal, cvt-full),
:90 (cpt-linear-solve, cpt-fixed-point, cpt-quasi-newton), :104 (odt-dp-fp,
odt-fixed-point, odt-bfgs), :125 (display form "CVT (block-diagonal)"), :194 (legacy
alias cvt-uniform-qnf).  Arithmetic: SURVEY.md Appendix A.4, A.8-A.10 (published
algorithms README.md:244-251).  Parity unpinned -- see oracle/__init__.py.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse
import scipy.sparse.linalg

from .meshtri import MeshTri

NOT_IMPLEMENTED = (
    "cvt-full",
    "cvt-uniform-qnf",
    "odt-bfgs",
)

# ODT: cells with a boundary edge contribute their barycenter instead of their circumcenter
# (the circumcenter of such a cell may lie outside the domain).  This is how the upstream
# package's ODT fixed-point update treats them as far as it could be recollected; SURVEY.md
# A.8 as written uses circumcenters everywhere -- set False for that variant.  The GPU library
# has the same switch (om_set_odt_boundary_barycenters) and the tests cover both settings.
ODT_BOUNDARY_BARYCENTERS = True


def normalize_method_name(name: str) -> str:
    # "CVT (block-diagonal)" -> "cvt-block-diagonal"   (README.md:80 vs :125)
    return "-".join(name.lower().replace("(", "").replace(")", "").split())


def lloyd(mesh: MeshTri) -> np.ndarray:
    """A.4: centroid of the Voronoi control volume; cells with an angle > 135 deg masked."""
    mask = np.any(mesh.ce_ratios < -0.5, axis=0)
    X = mesh.get_control_volume_centroids(cell_mask=mask)
    idx = np.any(np.isnan(X), axis=1)
    X[idx] = mesh.points[idx]
    return X


def _volume_averaged(mesh: MeshTri, reference_points: np.ndarray) -> np.ndarray:
    """A.8: x_i = sum_c |c| r_c / sum_c |c| over adjacent cells; boundary pinned."""
    vol = mesh.cell_volumes
    scaled = reference_points * vol[:, None]
    num = np.zeros(mesh.points.shape)
    den = np.zeros(mesh.n)
    for i in mesh.cells("points").T:
        np.add.at(num, i, scaled)
        np.add.at(den, i, vol)
    idx = mesh.is_interior_point
    new = mesh.points.copy()
    new[idx] = num[idx] / den[idx][:, None]
    return new


def cpt_fixed_point(mesh: MeshTri) -> np.ndarray:
    return _volume_averaged(mesh, mesh.cell_barycenters)


def _count_averaged(mesh: MeshTri, reference_points: np.ndarray) -> np.ndarray:
    """Density-preserving variant: density ~ 1/|c| makes every adjacent cell count the same,
    x_i = mean_c r_c over adjacent cells; boundary pinned."""
    num = np.zeros(mesh.points.shape)
    den = np.zeros(mesh.n)
    for i in mesh.cells("points").T:
        np.add.at(num, i, reference_points)
        np.add.at(den, i, 1.0)
    idx = mesh.is_interior_point
    new = mesh.points.copy()
    new[idx] = num[idx] / den[idx][:, None]
    return new


def _odt_reference_points(mesh: MeshTri) -> np.ndarray:
    ref = mesh.cell_circumcenters.copy()
    if ODT_BOUNDARY_BARYCENTERS:
        bc = mesh.is_boundary_cell
        ref[bc] = mesh.cell_barycenters[bc]
    return ref


def odt_fixed_point(mesh: MeshTri) -> np.ndarray:
    return _volume_averaged(mesh, _odt_reference_points(mesh))


def odt_dp_fp(mesh: MeshTri) -> np.ndarray:
    """README.md:104: density-preserving ODT fixed-point iteration (count averaged)."""
    return _count_averaged(mesh, _odt_reference_points(mesh))


def cvt_block_diagonal(mesh: MeshTri) -> np.ndarray:
    """A.9: quasi-Newton with the d x d diagonal blocks of the CVT Hessian."""
    X = mesh.points
    n, d = X.shape
    mask = np.any(mesh.ce_ratios < -0.5, axis=0)
    cv = mesh.get_control_volumes(cell_mask=mask)
    cen = mesh.get_control_volume_centroids(cell_mask=mask)
    blocks = np.zeros((n, d, d))
    for k in range(d):
        blocks[:, k, k] += 2 * cv
    idx = mesh.idx_hierarchy[:, :, ~mask]  # (2,3,C')
    e = mesh.half_edge_coords[:, ~mask]  # (3,C',d)
    ce = mesh.ce_ratios[:, ~mask]  # (3,C')
    for k in range(3):
        m = -0.5 * ce[k][:, None, None] * np.einsum("ci,cj->cij", e[k], e[k])
        np.add.at(blocks, idx[0, k], m)
        np.add.at(blocks, idx[1, k], m)
    rhs = -2 * (X - cen) * cv[:, None]
    bnd = mesh.is_boundary_point
    # vertices without any unmasked cell (cv == 0, centroid 0/0) and orphans stay put
    dead = bnd | (cv == 0.0)
    blocks[dead] = 0.0
    for k in range(d):
        blocks[dead, k, k] = 1.0
    rhs[dead] = 0.0
    return X + np.linalg.solve(blocks, rhs[..., None])[..., 0]


def laplacian_matrix(mesh: MeshTri):
    """A.10: L = sum_cells sum_edges [[1,-1],[-1,1]]; Dirichlet rows -> identity."""
    cells = mesh.cells("points")
    n = mesh.n
    rows, cols, vals = [], [], []
    one = np.ones(cells.shape[0])
    for a, b in ((0, 1), (1, 2), (2, 0)):
        u, v = cells[:, a], cells[:, b]
        rows += [u, v, u, v]
        cols += [u, v, v, u]
        vals += [one, one, -one, -one]
    L = scipy.sparse.coo_matrix(
        (np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)
    ).tocsr()
    fixed = ~mesh.is_interior_point  # boundary + orphan vertices
    keep = scipy.sparse.diags((~fixed).astype(float))
    L = keep @ L + scipy.sparse.diags(fixed.astype(float))
    return L.tocsr(), fixed


def cpt_linear_solve(mesh: MeshTri) -> np.ndarray:
    L, fixed = laplacian_matrix(mesh)
    rhs = np.zeros(mesh.points.shape)
    rhs[fixed] = mesh.points[fixed]
    out = scipy.sparse.linalg.spsolve(L.tocsc(), rhs)
    return np.asarray(out).reshape(mesh.points.shape)


def cpt_quasi_newton(mesh: MeshTri) -> np.ndarray:
    """README.md:90, :97-98: uniform-density CPT, one quasi-Newton step x - H^-1 dE.

    Chen-Holst: dE_i = 2/(d+1) sum_{t in star(i)} |t| (x_i - b_t).  Differentiating with the
    cell areas held fixed gives  d_ii E = 2/(d+1) |w_i| - 2/(d+1)^2 |w_i|  and
    d_ij E = -2/(d+1)^2 (sum of the areas of the cells on edge ij); the approximate Hessian
    drops the negative part of the diagonal (it hurts convergence), keeping
    H_ii = 2/(d+1) |w_i|.  d = 2 also on surfaces.  Boundary rows: identity, right-hand side 0.
    The same scalar matrix is solved for every coordinate."""
    X = mesh.points
    n = mesh.n
    cells = mesh.cells("points")
    vol = mesh.cell_volumes
    dim = 2
    jac = np.zeros(X.shape)
    bary = mesh.cell_barycenters
    for k in range(3):
        np.add.at(jac, cells[:, k], (X[cells[:, k]] - bary) * vol[:, None])
    jac *= 2.0 / (dim + 1)
    rows, cols, vals = [], [], []
    for k in range(3):
        rows.append(cells[:, k])
        cols.append(cells[:, k])
        vals.append(2.0 / (dim + 1) * vol)
        for kk in ((k + 1) % 3, (k + 2) % 3):
            rows.append(cells[:, k])
            cols.append(cells[:, kk])
            vals.append(-2.0 / (dim + 1) ** 2 * vol)
    H = scipy.sparse.coo_matrix(
        (np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)
    ).tocsr()
    fixed = ~mesh.is_interior_point  # boundary + orphan vertices
    keep = scipy.sparse.diags((~fixed).astype(float))
    H = (keep @ H + scipy.sparse.diags(fixed.astype(float))).tocsc()
    rhs = jac.copy()
    rhs[fixed] = 0.0
    step = np.asarray(scipy.sparse.linalg.spsolve(H, rhs)).reshape(X.shape)
    return X - step


METHODS = {
    "lloyd": lloyd,
    "cvt-block-diagonal": cvt_block_diagonal,
    "cpt-fixed-point": cpt_fixed_point,
    "cpt-linear-solve": cpt_linear_solve,
    "odt-fixed-point": odt_fixed_point,
    "odt-dp-fp": odt_dp_fp,
    "cpt-quasi-newton": cpt_quasi_newton,
}


def get_new_points(mesh: MeshTri, method: str) -> np.ndarray:
    """README.md:141: ``optimesh.get_new_points(mesh, "CVT (block-diagonal)")``."""
    name = normalize_method_name(method)
    if name in NOT_IMPLEMENTED:
        raise NotImplementedError(f"method {name!r} is outside the hot-path scope")
    if name not in METHODS:
        raise KeyError(f"unknown method {method!r}; valid: {sorted(METHODS)}")
    return METHODS[name](mesh)
