"""The reference's Python API for the smoothing path, backed by the CUDA library.

    optimize_points_cells(points, cells, method, tol, max_num_steps, omega=...)
        /root/reference/README.md:124-126 (kwargs :166-175)
    optimize(mesh, method, tol, max_num_steps, ...)        README.md:131-132
    get_new_points(mesh, method)                           README.md:141

Loop semantics: flip-until-Delaunay once, then per step: new points,
pin boundary, diff = omega (new - x), is_final = all |diff|^2 < tol^2 or k >=
max_num_steps, step limiter, implicit-surface projection, flip-until-Delaunay, stop if
is_final.  Everything per-step runs on the GPU; Python only keeps the signature,
argument checking and the optional hooks (callback, step dumps, generic surfaces).
"""
from __future__ import annotations

import numpy as np

from .helpers import print_stats
from .mesh import DeviceMesh, MeshTri, method_id, normalize_method_name
from .surfaces import Sphere


def _project_host(dm: DeviceMesh, surface, tol: float, max_iter: int = 100):
    """Generic implicit surface (objects with f/grad, README.md:157-162): the user's
    Python callables run on the host between the point update and the flips."""
    x = dm.points.T.copy()
    changed = False
    for _ in range(max_iter):
        fval = surface.f(x)
        if np.all(np.abs(fval) <= tol):
            break
        grad = surface.grad(x)
        x = x - grad * (fval / np.einsum("ij,ij->j", grad, grad))
        changed = True
    if changed:
        dm.points = x.T


class _DeviceHooks:
    """The per-step hooks on DEVICE memory: the user's callables get torch
    CUDA tensors that alias the library's arrays, so a generic implicit surface
    (README.md:157-162: ``f(x)``, ``grad(x)`` on ``(3, n)``) and a ``boundary_step`` callback
    (README.md:146-149: ``(d, n_boundary)`` targets in, moved targets out) run between the
    device kernels without a host round trip.  Opt-in (``device_callables=True`` or an
    ``on_device = True`` attribute on the object): the callables must be written for
    array-API style arguments (``x[0] ** 2``, ``-2 * x``, ``torch.sqrt`` ...), numpy functions
    do not accept CUDA tensors.  The boundary targets arrive in the caller's vertex order."""

    def __init__(self, dm: DeviceMesh, device: int):
        import torch

        from .dist import _DevPtr

        self.torch, self._DevPtr = torch, _DevPtr
        self.dm, self.dev = dm, torch.device("cuda", device)
        self.n, self.dim = dm.n, dm.dim
        _, _, perm_ptr, self.stride = dm.device_ptrs()
        bnd = np.nonzero(dm.is_boundary_point)[0]  # caller ids, ascending
        ids = torch.as_tensor(bnd, device=self.dev, dtype=torch.int64)
        if perm_ptr:
            perm = torch.as_tensor(_DevPtr(perm_ptr, (self.n,), "<i4"), device=self.dev).long()
            inv = torch.empty(self.n, dtype=torch.int64, device=self.dev)
            inv[perm] = torch.arange(self.n, device=self.dev)
            ids = inv[ids]
        self.bidx = ids  # internal ids of the boundary vertices, in the caller's order

    def _view(self, ptr):
        t = self.torch.as_tensor(self._DevPtr(ptr, (self.n, self.stride), "<f8"), device=self.dev)
        return t[:, : self.dim]

    def _fence(self):
        # the library runs on its own stream, torch on its current one
        self.dm.synchronize()

    def update(self, omega: float, tol: float, boundary_step) -> dict:
        torch = self.torch
        ptr = self.dm.targets_device()
        if self.bidx.numel():
            self._fence()
            T = self._view(ptr)
            moved = boundary_step(T[self.bidx].T.contiguous())
            if not torch.is_tensor(moved):
                moved = torch.as_tensor(np.asarray(moved), device=self.dev)
            if tuple(moved.shape) != (self.dim, self.bidx.numel()):
                raise ValueError("boundary_step must return an array of the shape it was given")
            T[self.bidx] = moved.T.to(torch.float64)
            torch.cuda.current_stream(self.dev).synchronize()
        return self.dm.update_from_targets(ptr, tol)

    def project(self, surface, tol: float, max_iter: int = 100) -> int:
        torch = self.torch
        self._fence()
        ptr = self.dm.device_ptrs()[0]
        X = self._view(ptr)
        sweeps = 0
        for _ in range(max_iter):
            x = X.T
            fval = surface.f(x)
            if not bool((fval.abs() <= tol).all()):
                grad = surface.grad(x)
                X -= (grad * (fval / (grad * grad).sum(0))).T
                sweeps += 1
            else:
                break
        torch.cuda.current_stream(self.dev).synchronize()
        return sweeps


def _wants_device(obj, flag) -> bool:
    return obj is not None and (bool(flag) or bool(getattr(obj, "on_device", False)))


def _half_min_inradius(x: np.ndarray, cells: np.ndarray) -> np.ndarray:
    """Per vertex: half of the smallest inradius over its cells (the loop's step limit)."""
    p0, p1, p2 = x[cells[:, 0]], x[cells[:, 1]], x[cells[:, 2]]
    e0, e1, e2 = p2 - p1, p0 - p2, p1 - p0
    l0, l1, l2 = (np.sqrt(np.einsum("ij,ij->i", e, e)) for e in (e0, e1, e2))
    # |e1 x e2|^2 = |e1|^2 |e2|^2 - (e1.e2)^2 in any embedding dimension
    d12 = np.einsum("ij,ij->i", e1, e2)
    area = 0.5 * np.sqrt(np.maximum(l1 * l1 * l2 * l2 - d12 * d12, 0.0))
    r_in = 2.0 * area / (l0 + l1 + l2)
    out = np.full(x.shape[0], np.inf)
    np.minimum.at(out, cells.reshape(-1), np.repeat(r_in, 3))
    return 0.5 * out


def _host_update(dm: DeviceMesh, omega: float, tol: float, boundary_step) -> dict:
    """One point update with a ``boundary_step`` callback: the targets come from the device,
    the callback moves the boundary targets (e.g. back onto the domain boundary) on the
    host, then relaxation and the step limiter are applied to every vertex as the reference's
    loop does it.  This is the slow, hook-driven path; without the callback the whole update
    is one kernel."""
    x = dm.points
    new = dm.new_points()
    bnd = dm.is_boundary_point
    if bnd.any():
        moved = np.asarray(boundary_step(new[bnd].T), dtype=np.float64).T
        if moved.shape != new[bnd].shape:
            raise ValueError("boundary_step must return an array of the shape it was given")
        new[bnd] = moved
    diff = omega * (new - x)
    len2 = np.einsum("ij,ij->i", diff, diff)
    max_diff2 = float(len2.max()) if len2.size else 0.0
    limit = _half_min_inradius(x, np.asarray(dm.cells(np.int64)))
    length = np.sqrt(len2)
    idx = length > limit
    diff[idx] *= (limit[idx] / length[idx])[:, None]
    dm.points = x + diff
    return {"max_diff2": max_diff2, "n_limited": int(idx.sum()), "n_flips": 0, "n_flip_rounds": 0,
            "flip_cap_hit": 0, "is_final": int(max_diff2 < tol * tol), "solver_iters": 0,
            "surface_sweeps": 0, "stale": 0}


def _run_loop(dm: DeviceMesh, method: str, tol: float, max_num_steps: int, omega: float = 1.0,
              verbose: bool = False, callback=None, step_filename_format=None,
              implicit_surface=None, implicit_surface_tol: float = 1.0e-10, boundary_step=None,
              cells_dtype=None, log=None, odt_boundary_barycenters: bool = True,
              device_callables: bool = False, device: int = 0):
    if boundary_step is not None and not callable(boundary_step):
        raise TypeError("boundary_step must be callable: (d, n) array -> (d, n) array")
    if max_num_steps < 1:
        raise ValueError("max_num_steps must be >= 1")
    dm.set_method(method, omega)
    dm.set_odt_boundary_barycenters(odt_boundary_barycenters)
    host_surface = None
    if implicit_surface is None:
        dm.clear_surface()
    elif isinstance(implicit_surface, Sphere):
        dm.set_sphere(implicit_surface.center, implicit_surface.radius, implicit_surface_tol)
    else:
        if not (hasattr(implicit_surface, "f") and hasattr(implicit_surface, "grad")):
            raise TypeError("implicit_surface must provide f(x) and grad(x)")
        dm.clear_surface()
        host_surface = implicit_surface
    dev_surface = _wants_device(host_surface, device_callables)
    dev_boundary = _wants_device(boundary_step, device_callables)
    hooks_dev = _DeviceHooks(dm, device) if (dev_surface or dev_boundary) else None

    if verbose:
        print("Before:")
        print_stats(*dm.stats())

    hooks = callback is not None or step_filename_format is not None or host_surface is not None \
        or log is not None or boundary_step is not None
    if not hooks:
        steps, last = dm.run(tol, max_num_steps)
    else:
        dm.flip_until_delaunay()
        steps = 0
        if callback is not None:
            callback(0, _snapshot(dm, cells_dtype))
        while True:
            steps += 1
            if host_surface is None and boundary_step is None:
                st = dm.step(tol)
            else:
                if boundary_step is None:
                    st = dm.update_points(tol)
                elif dev_boundary:
                    st = hooks_dev.update(omega, tol, boundary_step)
                else:
                    st = _host_update(dm, omega, tol, boundary_step)
                if host_surface is not None and dev_surface:
                    st["surface_sweeps"] = hooks_dev.project(host_surface, implicit_surface_tol)
                elif host_surface is not None:
                    _project_host(dm, host_surface, implicit_surface_tol)
                elif boundary_step is not None:
                    dm.project()  # the built-in surface, if one is set
                nf, nr = dm.flip_until_delaunay()
                st["n_flips"], st["n_flip_rounds"] = nf, nr
            is_final = bool(st["is_final"]) or steps >= max_num_steps
            if log is not None:
                log.append(dict(step=steps, **st))
            if step_filename_format is not None:
                from . import io

                io.write(step_filename_format.format(steps), dm.points, dm.cells(cells_dtype))
            if callback is not None:
                callback(steps, _snapshot(dm, cells_dtype))
            if is_final:
                break
    if verbose:
        print(f"\nFinal ({steps} steps):")
        print_stats(*dm.stats())
    return steps


def _snapshot(dm: DeviceMesh, cells_dtype):
    return MeshTri(dm.points, dm.cells(cells_dtype))


def optimize_points_cells(points, cells, method: str, tol: float, max_num_steps: int,
                          omega: float = 1.0, verbose: bool = False, callback=None,
                          step_filename_format=None, implicit_surface=None,
                          implicit_surface_tol: float = 1.0e-10, boundary_step=None,
                          method_name=None, device: int = 0, log=None,
                          odt_boundary_barycenters: bool = True, device_callables: bool = False):
    """Returns ``(points, cells)``; the inputs are not modified (README.md:124-126).

    ``device``, ``log`` (per-step statistics are appended to the list) and
    ``odt_boundary_barycenters`` (see ``DeviceMesh.set_odt_boundary_barycenters``) are
    additions of this build."""
    method_id(method)  # validate before touching the device
    cells = np.asarray(cells)
    with DeviceMesh(points, cells, device=device) as dm:
        _run_loop(dm, method, tol, max_num_steps, omega, verbose, callback, step_filename_format,
                  implicit_surface, implicit_surface_tol, boundary_step, cells.dtype, log,
                  odt_boundary_barycenters, device_callables, device)
        # large results come back in blocks of the library's pinned result cache: one DMA at
        # PCIe speed instead of a staged copy into 155k fresh pages (DeviceMesh.get_points)
        return dm.get_points(), dm.cells(cells.dtype)


def optimize(mesh, method: str, tol: float, max_num_steps: int, **kwargs):
    """In-place variant on an object exposing ``.points`` and ``.cells``
    (``meshplex.MeshTri`` or ``optimesh_b200.MeshTri``), README.md:131-133."""
    cells = _mesh_cells(mesh)
    points, new_cells = optimize_points_cells(mesh.points, cells, method, tol, max_num_steps,
                                              **kwargs)
    _mesh_assign(mesh, points, new_cells)
    return mesh


def get_new_points(mesh, method: str, device: int = 0,
                   odt_boundary_barycenters: bool = True) -> np.ndarray:
    """One un-relaxed update, ``(N, d)`` array (README.md:141)."""
    method_id(method)
    with DeviceMesh(mesh.points, _mesh_cells(mesh), device=device) as dm:
        dm.set_method(method, 1.0)
        dm.set_odt_boundary_barycenters(odt_boundary_barycenters)
        return dm.new_points()


def _mesh_cells(mesh):
    cells = mesh.cells
    if callable(cells) and not isinstance(cells, np.ndarray):
        cells = cells("points")
    elif isinstance(cells, dict):
        cells = cells["points"]
    return np.asarray(cells)


def _mesh_assign(mesh, points, cells):
    if isinstance(mesh, MeshTri):
        mesh.points = points
        mesh._set_cells(cells)
        return
    # foreign mesh classes (meshplex): rebuild through the constructor protocol
    try:
        mesh.__init__(points, cells)
    except Exception:
        mesh.points = points
        if isinstance(getattr(mesh, "cells", None), dict):
            mesh.cells["points"] = cells


__all__ = ["optimize_points_cells", "optimize", "get_new_points", "normalize_method_name"]
