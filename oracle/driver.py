"""Driver loop (oracle, test infrastructure).

Restates ``optimesh.optimize`` / ``optimize_points_cells``
(/root/reference/README.md:124-133, kwargs :166-175, surface protocol :157-162) per
SURVEY.md Appendix A.5.  Parity unpinned -- see oracle/__init__.py.
"""
from __future__ import annotations

import numpy as np

from .meshtri import MeshTri
from .methods import get_new_points


def project_to_surface(X, surface, tol=1.0e-10, max_iter=100):
    """Newton projection onto f = 0: x -= grad * f / |grad|^2 (README.md:157-162)."""
    x = X.T.copy()
    for _ in range(max_iter):
        fval = surface.f(x)
        if np.all(np.abs(fval) <= tol):
            break
        grad = surface.grad(x)
        grad_dot_grad = np.einsum("ij,ij->j", grad, grad)
        x = x - grad * (fval / grad_dot_grad)
    return x.T.copy()


def step(mesh: MeshTri, method: str, omega=1.0, implicit_surface=None,
         implicit_surface_tol=1.0e-10, limiter=True):
    """One smoothing step without the flip pass.

    Returns (max_i |diff_i|^2 before limiting, number of limited vertices).
    """
    X = mesh.points
    new = get_new_points(mesh, method)
    bnd = mesh.is_boundary_point
    new[bnd] = X[bnd]
    diff = omega * (new - X)
    diff2 = np.einsum("ij,ij->i", diff, diff)
    max_diff2 = float(diff2.max()) if diff2.size else 0.0
    n_limited = 0
    if limiter:
        max_step = np.full(mesh.n, np.inf)
        np.minimum.at(
            max_step, mesh.cells("points").reshape(-1), np.repeat(mesh.cell_inradius, 3)
        )
        max_step *= 0.5
        step_lengths = np.sqrt(diff2)
        idx = step_lengths > max_step
        diff[idx] *= (max_step / np.where(idx, step_lengths, 1.0))[idx, None]
        n_limited = int(idx.sum())
    Xn = X + diff
    if implicit_surface is not None:
        Xn = project_to_surface(Xn, implicit_surface, implicit_surface_tol)
    mesh.points = Xn
    return max_diff2, n_limited


def optimize(mesh: MeshTri, method: str, tol: float, max_num_steps: int, omega: float = 1.0,
             verbose: bool = False, callback=None, implicit_surface=None,
             implicit_surface_tol: float = 1.0e-10, log=None):
    """A.5.  Mutates ``mesh``.  Returns the number of steps taken."""
    mesh.flip_until_delaunay()
    k = 0
    while True:
        k += 1
        max_diff2, n_limited = step(mesh, method, omega, implicit_surface, implicit_surface_tol)
        is_final = (max_diff2 < tol * tol) or k >= max_num_steps
        nflips, nrounds = mesh.flip_until_delaunay()
        if log is not None:
            log.append(dict(step=k, max_diff2=max_diff2, n_limited=n_limited,
                            n_flips=nflips, n_rounds=nrounds))
        if callback is not None:
            callback(k, mesh)
        if is_final:
            break
    return k


def optimize_points_cells(points, cells, method, tol, max_num_steps, **kwargs):
    """README.md:124-126: returns ``(points, cells)``; inputs are not mutated."""
    mesh = MeshTri(points, cells)
    optimize(mesh, method, tol, max_num_steps, **kwargs)
    return mesh.points, mesh.cells("points").astype(np.asarray(cells).dtype)
