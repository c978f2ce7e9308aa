"""Mesh file I/O for the CLI and the ``step_filename_format`` dumps
(/root/reference/README.md:49-53, :174).

The reference delegates to meshio (README.md:21-22), which is imported lazily when it
is installed.  Without it, two formats are built in: legacy VTK (ASCII, triangles) and
``.npz`` (arrays ``points``, ``cells``).
"""
from __future__ import annotations

import os

import numpy as np


def _have_meshio():
    try:
        import meshio  # noqa: F401

        return True
    except Exception:
        return False


def read(path: str):
    """Returns (points (N,d) float64, triangle cells (C,3) int64)."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npz":
        with np.load(path) as z:
            return np.asarray(z["points"], dtype=np.float64), np.asarray(z["cells"], dtype=np.int64)
    if _have_meshio():
        import meshio

        m = meshio.read(path)
        tris = [c.data for c in m.cells if c.type == "triangle"]
        if not tris:
            raise ValueError(f"{path}: no triangle cells")
        return np.asarray(m.points, dtype=np.float64), np.concatenate(tris).astype(np.int64)
    if ext == ".vtk":
        return _read_vtk(path)
    raise ValueError(
        f"cannot read {path!r}: meshio is not installed; built-in formats are .vtk (legacy "
        "ASCII) and .npz")


def write(path: str, points, cells):
    points = np.asarray(points, dtype=np.float64)
    cells = np.asarray(cells)
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npz":
        np.savez(path, points=points, cells=cells)
        return
    if _have_meshio():
        import meshio

        meshio.write_points_cells(path, points, [("triangle", cells)])
        return
    if ext == ".vtk":
        _write_vtk(path, points, cells)
        return
    raise ValueError(
        f"cannot write {path!r}: meshio is not installed; built-in formats are .vtk (legacy "
        "ASCII) and .npz")


def _write_vtk(path, points, cells):
    n, d = points.shape
    p3 = np.zeros((n, 3))
    p3[:, :d] = points
    with open(path, "w") as f:
        f.write("# vtk DataFile Version 3.0\noptimesh_b200\nASCII\nDATASET UNSTRUCTURED_GRID\n")
        f.write(f"POINTS {n} double\n")
        np.savetxt(f, p3, fmt="%.17g")
        c = cells.shape[0]
        f.write(f"CELLS {c} {4 * c}\n")
        np.savetxt(f, np.column_stack([np.full(c, 3), cells]), fmt="%d")
        f.write(f"CELL_TYPES {c}\n")
        np.savetxt(f, np.full(c, 5), fmt="%d")


def _read_vtk(path):
    with open(path) as f:
        tok = f.read().split()
    up = [t.upper() for t in tok]
    if "ASCII" not in up[:40]:
        raise ValueError(f"{path}: only ASCII legacy VTK is built in (install meshio for more)")
    i = up.index("POINTS")
    n = int(tok[i + 1])
    pts = np.array(tok[i + 3:i + 3 + 3 * n], dtype=np.float64).reshape(n, 3)
    i = up.index("CELLS")
    c, total = int(tok[i + 1]), int(tok[i + 2])
    flat = np.array(tok[i + 3:i + 3 + total], dtype=np.int64)
    cells = []
    k = 0
    while k < total:
        m = flat[k]
        if m == 3:
            cells.append(flat[k + 1:k + 4])
        k += m + 1
    if not cells:
        raise ValueError(f"{path}: no triangle cells")
    if np.all(pts[:, 2] == 0.0):
        pts = pts[:, :2]
    return np.ascontiguousarray(pts), np.array(cells, dtype=np.int64)
