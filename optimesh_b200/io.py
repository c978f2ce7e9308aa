"""Mesh file I/O for the CLI and the ``step_filename_format`` dumps
(/root/reference/README.md:49-53, :174).

The reference delegates to meshio (README.md:21-22), which is imported lazily when it
is installed.  Without it, these formats are built in (triangles only): legacy VTK
(ASCII), Gmsh MSH 2.2 (ASCII; the physical and elementary tags become the cell fields
``gmsh:physical`` and ``gmsh:geometrical``, the names meshio uses), OFF, Wavefront OBJ and
``.npz`` (arrays ``points``, ``cells``, plus per-cell arrays).
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


def read(path: str, with_cell_data: bool = False):
    """Returns (points (N,d) float64, triangle cells (C,3) int64); with ``with_cell_data``
    also a dict of per-triangle arrays (e.g. the subdomain field of ``optimesh -s NAME``)."""
    points, cells, cell_data = _read(path)
    return (points, cells, cell_data) if with_cell_data else (points, cells)


def _read(path: str):
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npz":
        with np.load(path) as z:
            cells = np.asarray(z["cells"], dtype=np.int64)
            data = {k: np.asarray(z[k]) for k in z.files
                    if k not in ("points", "cells") and z[k].shape[:1] == cells.shape[:1]}
            return np.asarray(z["points"], dtype=np.float64), cells, data
    if _have_meshio():
        import meshio

        m = meshio.read(path)
        blocks = [i for i, c in enumerate(m.cells) if c.type == "triangle"]
        if not blocks:
            raise ValueError(f"{path}: no triangle cells")
        cells = np.concatenate([m.cells[i].data for i in blocks]).astype(np.int64)
        data = {k: np.concatenate([np.asarray(v[i]) for i in blocks])
                for k, v in (m.cell_data or {}).items()}
        return np.asarray(m.points, dtype=np.float64), cells, data
    if ext in _READERS:
        return _READERS[ext](path)
    raise ValueError(
        f"cannot read {path!r}: meshio is not installed; built-in formats are "
        f"{', '.join(sorted(_READERS))} and .npz")


def write(path: str, points, cells, cell_data=None):
    points = np.asarray(points, dtype=np.float64)
    cells = np.asarray(cells)
    cell_data = {k: np.asarray(v) for k, v in (cell_data or {}).items()}
    for k, v in cell_data.items():
        if v.shape[:1] != cells.shape[:1]:
            raise ValueError(f"cell data {k!r} has {v.shape[0]} entries for {cells.shape[0]} cells")
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npz":
        np.savez(path, points=points, cells=cells, **cell_data)
        return
    if _have_meshio():
        import meshio

        meshio.write_points_cells(path, points, [("triangle", cells)],
                                  cell_data={k: [v] for k, v in cell_data.items()})
        return
    if ext in _WRITERS:
        _WRITERS[ext](path, points, cells, cell_data)
        return
    raise ValueError(
        f"cannot write {path!r}: meshio is not installed; built-in formats are "
        f"{', '.join(sorted(_WRITERS))} and .npz")


def _write_vtk(path, points, cells, cell_data=None):
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
        if cell_data:
            f.write(f"CELL_DATA {c}\n")
            for name, v in cell_data.items():
                if v.ndim != 1:
                    raise ValueError("built-in VTK writer: scalar cell data only")
                is_int = np.issubdtype(v.dtype, np.integer)
                f.write(f"SCALARS {name.replace(' ', '_')} {'int' if is_int else 'double'} 1\n")
                f.write("LOOKUP_TABLE default\n")
                np.savetxt(f, v, fmt="%d" if is_int else "%.17g")


_VTK_INT_TYPES = ("bit", "char", "unsigned_char", "short", "unsigned_short", "int",
                  "unsigned_int", "long", "unsigned_long", "vtktypeint32", "vtktypeint64",
                  "vtkidtype")


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
    cells, is_tri = [], []
    k = 0
    while k < total:
        m = flat[k]
        is_tri.append(m == 3)
        if m == 3:
            cells.append(flat[k + 1:k + 4])
        k += m + 1
    if not cells:
        raise ValueError(f"{path}: no triangle cells")
    is_tri = np.array(is_tri)
    data = {}
    if "CELL_DATA" in up:
        j = up.index("CELL_DATA")
        nc = int(tok[j + 1])
        j += 2
        while j < len(tok) and up[j] not in ("POINT_DATA",):
            if up[j] == "SCALARS":
                name, typ = tok[j + 1], tok[j + 2].lower()
                j += 3
                ncomp = 1
                if j < len(tok) and tok[j].isdigit():
                    ncomp = int(tok[j])
                    j += 1
                if up[j] == "LOOKUP_TABLE":
                    j += 2
                vals = np.array(tok[j:j + nc * ncomp],
                                dtype=np.int64 if typ in _VTK_INT_TYPES else np.float64)
                j += nc * ncomp
                vals = vals.reshape(nc, ncomp)[is_tri[:nc]] if nc == len(is_tri) else None
                if vals is not None:
                    data[name] = vals[:, 0] if ncomp == 1 else vals
            elif up[j] == "FIELD":
                narr = int(tok[j + 2])
                j += 3
                for _ in range(narr):
                    name, ncomp, ntup, typ = tok[j], int(tok[j + 1]), int(tok[j + 2]), tok[j + 3].lower()
                    j += 4
                    vals = np.array(tok[j:j + ncomp * ntup],
                                    dtype=np.int64 if typ in _VTK_INT_TYPES else np.float64)
                    j += ncomp * ntup
                    if ntup == len(is_tri):
                        vals = vals.reshape(ntup, ncomp)[is_tri]
                        data[name] = vals[:, 0] if ncomp == 1 else vals
            else:
                j += 1
    if np.all(pts[:, 2] == 0.0):
        pts = pts[:, :2]
    return np.ascontiguousarray(pts), np.array(cells, dtype=np.int64), data


def _flatten_if_planar(pts):
    if pts.shape[1] == 3 and np.all(pts[:, 2] == 0.0):
        pts = pts[:, :2]
    return np.ascontiguousarray(pts, dtype=np.float64)


def _pad3(points):
    p3 = np.zeros((points.shape[0], 3))
    p3[:, :points.shape[1]] = points
    return p3


# ---- Gmsh MSH 2.2 ASCII
def _read_msh(path):
    with open(path) as f:
        lines = [l.strip() for l in f]
    def section(name):
        try:
            a = lines.index("$" + name)
            b = lines.index("$End" + name)
        except ValueError:
            return None
        return lines[a + 1:b]
    fmt = section("MeshFormat")
    if not fmt or not fmt[0].startswith("2") or fmt[0].split()[1] != "0":
        raise ValueError(f"{path}: only Gmsh MSH 2.x ASCII is built in (install meshio for more)")
    nodes = section("Nodes")
    n = int(nodes[0])
    rows = np.array([l.split() for l in nodes[1:1 + n]], dtype=np.float64)
    ids = rows[:, 0].astype(np.int64)
    lookup = np.full(ids.max() + 1, -1, dtype=np.int64)
    lookup[ids] = np.arange(n)
    cells, phys, geom = [], [], []
    for l in section("Elements")[1:]:
        t = l.split()
        if int(t[1]) != 2:  # 2 = 3-node triangle
            continue
        ntags = int(t[2])
        tags = [int(x) for x in t[3:3 + ntags]]
        phys.append(tags[0] if ntags > 0 else 0)
        geom.append(tags[1] if ntags > 1 else 0)
        cells.append([int(x) for x in t[3 + ntags:6 + ntags]])
    if not cells:
        raise ValueError(f"{path}: no triangle cells")
    cells = lookup[np.array(cells, dtype=np.int64)]
    data = {"gmsh:physical": np.array(phys, dtype=np.int64),
            "gmsh:geometrical": np.array(geom, dtype=np.int64)}
    return _flatten_if_planar(rows[:, 1:4]), cells, data


def _write_msh(path, points, cells, cell_data=None):
    cell_data = cell_data or {}
    c = cells.shape[0]
    phys = np.asarray(cell_data.get("gmsh:physical", np.ones(c)), dtype=np.int64)
    geom = np.asarray(cell_data.get("gmsh:geometrical", phys), dtype=np.int64)
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n")
        f.write(f"{points.shape[0]}\n")
        p3 = _pad3(points)
        for i, p in enumerate(p3):
            f.write(f"{i + 1} {p[0]:.17g} {p[1]:.17g} {p[2]:.17g}\n")
        f.write(f"$EndNodes\n$Elements\n{c}\n")
        for i in range(c):
            a, b, d = cells[i] + 1
            f.write(f"{i + 1} 2 2 {phys[i]} {geom[i]} {a} {b} {d}\n")
        f.write("$EndElements\n")


# ---- OFF
def _read_off(path):
    with open(path) as f:
        tok = [t for l in f for t in l.split("#")[0].split()]
    if not tok or tok[0].upper() != "OFF":
        raise ValueError(f"{path}: not an OFF file")
    n, c = int(tok[1]), int(tok[2])
    pts = np.array(tok[4:4 + 3 * n], dtype=np.float64).reshape(n, 3)
    k = 4 + 3 * n
    cells = []
    for _ in range(c):
        m = int(tok[k])
        if m == 3:
            cells.append([int(x) for x in tok[k + 1:k + 4]])
        k += m + 1
    if not cells:
        raise ValueError(f"{path}: no triangle cells")
    return _flatten_if_planar(pts), np.array(cells, dtype=np.int64), {}


def _write_off(path, points, cells, cell_data=None):
    with open(path, "w") as f:
        f.write(f"OFF\n{points.shape[0]} {cells.shape[0]} 0\n")
        np.savetxt(f, _pad3(points), fmt="%.17g")
        np.savetxt(f, np.column_stack([np.full(cells.shape[0], 3), cells]), fmt="%d")


# ---- Wavefront OBJ (vertices and triangular faces)
def _read_obj(path):
    pts, cells = [], []
    with open(path) as f:
        for l in f:
            t = l.split()
            if not t:
                continue
            if t[0] == "v":
                pts.append([float(x) for x in t[1:4]])
            elif t[0] == "f" and len(t) == 4:
                cells.append([int(x.split("/")[0]) for x in t[1:4]])
    if not cells:
        raise ValueError(f"{path}: no triangle cells")
    cells = np.array(cells, dtype=np.int64)
    cells = np.where(cells < 0, cells + len(pts), cells - 1)  # negative = relative to the end
    return _flatten_if_planar(np.array(pts, dtype=np.float64)), cells, {}


def _write_obj(path, points, cells, cell_data=None):
    with open(path, "w") as f:
        for p in _pad3(points):
            f.write(f"v {p[0]:.17g} {p[1]:.17g} {p[2]:.17g}\n")
        for a, b, d in cells + 1:
            f.write(f"f {a} {b} {d}\n")


_READERS = {".vtk": _read_vtk, ".msh": _read_msh, ".off": _read_off, ".obj": _read_obj}
_WRITERS = {".vtk": _write_vtk, ".msh": _write_msh, ".off": _write_off, ".obj": _write_obj}
