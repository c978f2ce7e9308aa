"""MeshTri-lite: per-cell geometry, vertex aggregates, edges and Delaunay flips.

Oracle (test infrastructure).  Restates the parts of meshplex's ``MeshTri`` that the
optimesh step touches (constructor + ``.points``/``.cells`` shown at
/root/reference/README.md:128-133; everything else per SURVEY.md Appendix A.1-A.7,
parity unpinned -- see oracle/__init__.py).

Conventions (A.1): cell ``c`` has vertices ``cells[c] = (p0, p1, p2)``; local edge
``k`` is opposite local vertex ``k`` and runs from vertex ``(k+1)%3`` to ``(k+2)%3``.
A half-edge is the pair (cell c, local edge k), numbered ``h = 3*c + k``.
"""
from __future__ import annotations

import warnings

import numpy as np


class DegenerateCellsError(ValueError):
    """A cell has zero area (upstream: ``MeshplexError("Degenerate cells.")``)."""


class MeshTri:
    def __init__(self, points, cells):
        self._points = np.array(points, dtype=np.float64, order="C")
        cells = np.asarray(cells)
        if cells.ndim != 2 or cells.shape[1] != 3:
            raise ValueError("cells must have shape (C, 3)")
        self._cells = np.array(cells, dtype=np.int64, order="C")
        self._twin = None
        self._geo = None

    # ------------------------------------------------------------------ state
    @property
    def points(self):
        return self._points

    @points.setter
    def points(self, new):
        # A.16 / upstream setter: drops every cached geometric quantity.
        self._points = np.array(new, dtype=np.float64, order="C")
        self._geo = None

    def cells(self, which="points"):
        assert which == "points"
        return self._cells

    @property
    def n(self):
        return self._points.shape[0]

    # --------------------------------------------------------------- geometry
    def _geometry(self):
        """A.1-A.3 for all cells at once."""
        if self._geo is not None:
            return self._geo
        X = self._points
        c = self._cells
        # idx[0,k] = cells[:,(k+1)%3], idx[1,k] = cells[:,(k+2)%3]
        idx = np.array(
            [[c[:, 1], c[:, 2], c[:, 0]], [c[:, 2], c[:, 0], c[:, 1]]]
        )  # (2,3,C)
        e = X[idx[1]] - X[idx[0]]  # (3,C,d)
        ee = np.einsum("kcd,kcd->kc", e, e)
        ed = np.array(
            [
                np.einsum("cd,cd->c", e[1], e[2]),
                np.einsum("cd,cd->c", e[2], e[0]),
                np.einsum("cd,cd->c", e[0], e[1]),
            ]
        )
        vol2 = 0.25 * (ed[2] * ed[0] + ed[0] * ed[1] + ed[1] * ed[2])
        vol2 = np.where(vol2 < 0.0, 0.0, vol2)
        vol = np.sqrt(vol2)
        if np.any(vol == 0.0):
            raise DegenerateCellsError("Degenerate cells.")
        ce = -ed * 0.25 / vol[None]
        part = 0.25 * ee * ce
        alpha = ee * ed
        beta = alpha / (alpha[0] + alpha[1] + alpha[2])[None]
        Xc = X[c.T]  # (3,C,d)
        a = Xc * beta[..., None]
        cc = a[0] + a[1] + a[2]
        self._geo = dict(idx=idx, e=e, ee=ee, ed=ed, vol=vol, ce=ce, part=part, cc=cc, Xc=Xc)
        return self._geo

    @property
    def idx_hierarchy(self):
        return self._geometry()["idx"]

    @property
    def half_edge_coords(self):
        return self._geometry()["e"]

    @property
    def ei_dot_ei(self):
        return self._geometry()["ee"]

    @property
    def ei_dot_ej(self):
        return self._geometry()["ed"]

    @property
    def cell_volumes(self):
        return self._geometry()["vol"]

    @property
    def ce_ratios(self):
        return self._geometry()["ce"]

    @property
    def cell_partitions(self):
        return self._geometry()["part"]

    @property
    def cell_circumcenters(self):
        return self._geometry()["cc"]

    @property
    def cell_barycenters(self):
        Xc = self._geometry()["Xc"]
        return (Xc[0] + Xc[1] + Xc[2]) / 3.0

    cell_centroids = cell_barycenters

    @property
    def edge_lengths(self):
        return np.sqrt(self.ei_dot_ei)

    @property
    def cell_inradius(self):
        abc = np.sqrt(self.ei_dot_ei)
        return 2 * self.cell_volumes / np.sum(abc, axis=0)

    @property
    def cell_circumradius(self):
        a, b, c = np.sqrt(self.ei_dot_ei)
        return (a * b * c) / (4.0 * self.cell_volumes)

    @property
    def q_radius_ratio(self):
        a, b, c = np.sqrt(self.ei_dot_ei)
        return (-a + b + c) * (a - b + c) * (a + b - c) / (a * b * c)

    @property
    def angles(self):
        """Angle at local vertex k, radians, shape (3,C) (A.3)."""
        l = np.sqrt(self.ei_dot_ei)
        ed = self.ei_dot_ej
        cosv = np.array(
            [-ed[0] / (l[1] * l[2]), -ed[1] / (l[2] * l[0]), -ed[2] / (l[0] * l[1])]
        )
        return np.arccos(np.clip(cosv, -1.0, 1.0))

    # ------------------------------------------------------- vertex aggregates
    def get_control_volumes(self, cell_mask=None):
        """A.4: cv_i = sum over unmasked cells of the partitions of the 2 edges at i."""
        g = self._geometry()
        part, idx = g["part"], g["idx"]
        if cell_mask is not None:
            part = part[:, ~cell_mask]
            idx = idx[:, :, ~cell_mask]
        vals = np.array([part, part])
        return np.bincount(idx.reshape(-1), vals.reshape(-1), minlength=self.n)

    def get_control_volume_centroids(self, cell_mask=None):
        """A.4: centroid of the (mesh-clipped) Voronoi control volume of each vertex."""
        g = self._geometry()
        part, idx, cc = g["part"], g["idx"], g["cc"]
        corner = self._points[idx]  # (2,3,C,d)
        mid = 0.5 * (corner[0] + corner[1])
        average = (corner + mid[None] + cc[None, None]) / 3.0
        contribs = part[None, :, :, None] * average
        if cell_mask is not None:
            idx = idx[:, :, ~cell_mask]
            contribs = contribs[:, :, ~cell_mask]
        d = self._points.shape[1]
        flat = idx.reshape(-1)
        num = np.array(
            [np.bincount(flat, contribs[..., k].reshape(-1), minlength=self.n) for k in range(d)]
        ).T
        cv = self.get_control_volumes(cell_mask)
        with np.errstate(invalid="ignore", divide="ignore"):
            return num / cv[:, None]

    @property
    def control_volumes(self):
        return self.get_control_volumes()

    # ------------------------------------------------------------- topology
    def _build_twins(self):
        """A.6: half-edge twin table.  twin[h] = partner half-edge or -1 (boundary)."""
        c = self._cells
        C = c.shape[0]
        u = np.stack([c[:, 1], c[:, 2], c[:, 0]], axis=1).reshape(-1)
        v = np.stack([c[:, 2], c[:, 0], c[:, 1]], axis=1).reshape(-1)
        lo = np.minimum(u, v)
        hi = np.maximum(u, v)
        key = lo * np.int64(self.n) + hi
        order = np.argsort(key, kind="stable")
        ks = key[order]
        same_next = np.zeros(3 * C, dtype=bool)
        same_next[:-1] = ks[1:] == ks[:-1]
        same_prev = np.zeros(3 * C, dtype=bool)
        same_prev[1:] = same_next[:-1]
        if np.any(same_next & same_prev):
            raise ValueError("non-manifold edge (more than 2 adjacent cells)")
        twin = np.full(3 * C, -1, dtype=np.int64)
        first = np.nonzero(same_next)[0]
        twin[order[first]] = order[first + 1]
        twin[order[first + 1]] = order[first]
        self._twin = twin
        return twin

    @property
    def twins(self):
        if self._twin is None:
            self._build_twins()
        return self._twin

    @property
    def is_boundary_point(self):
        twin = self.twins
        c = self._cells
        bh = np.nonzero(twin < 0)[0]
        cc_, k = bh // 3, bh % 3
        flag = np.zeros(self.n, dtype=bool)
        flag[c[cc_, (k + 1) % 3]] = True
        flag[c[cc_, (k + 2) % 3]] = True
        return flag

    @property
    def is_boundary_cell(self):
        """Cells with at least one boundary edge (meshplex: ``is_boundary_cell``)."""
        return np.any(self.twins.reshape(-1, 3) < 0, axis=1)

    @property
    def is_interior_point(self):
        # vertices that belong to at least one cell and are not on the boundary
        used = np.zeros(self.n, dtype=bool)
        used[self._cells.reshape(-1)] = True
        return used & ~self.is_boundary_point

    def flip_round_survey(self, tol=0.0):
        """One flip round with the selection rule of SURVEY.md A.7 AS WRITTEN: flag all
        non-Delaunay interior edges; while some cell touches two or more flagged edges, every
        such cell keeps only its most negative one (an edge dropped by either of its cells is
        unflagged); flip what remains.  Returns the number of flips.

        `flip_round` below implements this build's own single-pass rule instead (the upstream
        rule cannot be pinned, see oracle/__init__.py).  Both rules only ever flip
        non-Delaunay edges, one per cell and round, so they walk to the same fixed point --
        the Delaunay triangulation, unique for points in general position -- possibly in a
        different number of rounds and with different cell rows; `tests/test_oracle.py`
        checks that the canonical cells agree."""
        twin = self.twins
        C = self._cells.shape[0]
        ceh = self.ce_ratios.T.reshape(-1)
        interior = twin >= 0
        s = np.full(3 * C, np.inf)
        s[interior] = ceh[interior] + ceh[twin[interior]]
        flagged = s < -tol
        if not flagged.any():
            return 0
        while True:
            per_cell = flagged.reshape(C, 3).sum(axis=1)
            crit = np.nonzero(per_cell > 1)[0]
            if crit.size == 0:
                break
            sv = np.where(flagged, s, np.inf).reshape(C, 3)
            for c in crit:
                keep = int(np.argmin(sv[c]))
                for k in range(3):
                    if k != keep and flagged[3 * c + k]:
                        flagged[3 * c + k] = False
                        flagged[twin[3 * c + k]] = False
        h0 = np.nonzero(flagged)[0]
        h1 = twin[h0]
        sel = h0 < h1
        h0, h1 = h0[sel], h1[sel]
        if h0.size == 0:
            return 0
        self._apply_flips(h0, h1)
        return int(h0.size)

    def _apply_flips(self, h0, h1):
        a0, k0 = h0 // 3, h0 % 3
        a1, k1 = h1 // 3, h1 % 3
        cells = self._cells
        v0 = cells[a0, k0]
        v1 = cells[a1, k1]
        v2 = cells[a0, (k0 + 1) % 3]
        v3 = cells[a0, (k0 + 2) % 3]
        cells[a0] = np.stack([v0, v1, v2], axis=1)
        cells[a1] = np.stack([v0, v1, v3], axis=1)
        self._twin = None
        self._geo = None

    def flip_until_delaunay_survey(self, tol=0.0, max_steps=100):
        total = rounds = 0
        for _ in range(max_steps):
            n = self.flip_round_survey(tol)
            if n == 0:
                return total, rounds
            total += n
            rounds += 1
        return total, rounds

    def flip_round(self, tol=0.0):
        """One simultaneous flip round (A.7).  Returns the number of flips.

        Rule: flag interior edges with s = ce_k0(c0) + ce_k1(c1) < -tol; every cell
        keeps only its most negative flagged edge (ties: lowest local index); an edge
        is flipped iff it is kept by both adjacent cells.  (a0,k0) is the half-edge
        with the smaller id 3c+k.
        """
        twin = self.twins
        ce = self.ce_ratios  # (3,C)
        C = self._cells.shape[0]
        ceh = ce.T.reshape(-1)  # index by h = 3c+k
        interior = twin >= 0
        s = np.full(3 * C, np.inf)
        s[interior] = ceh[interior] + ceh[twin[interior]]
        flagged = s < -tol
        if not flagged.any():
            return 0
        sv = np.where(flagged, s, np.inf).reshape(C, 3)
        best = np.argmin(sv, axis=1)
        has = flagged.reshape(C, 3).any(axis=1)
        chosen = np.zeros((C, 3), dtype=bool)
        chosen[np.arange(C)[has], best[has]] = True
        chosen = chosen.reshape(-1)
        h0 = np.nonzero(chosen)[0]
        h1 = twin[h0]
        sel = chosen[h1] & (h0 < h1)
        h0, h1 = h0[sel], h1[sel]
        if h0.size == 0:
            return 0
        a0, k0 = h0 // 3, h0 % 3
        a1, k1 = h1 // 3, h1 % 3
        cells = self._cells
        v0 = cells[a0, k0]
        v1 = cells[a1, k1]
        v2 = cells[a0, (k0 + 1) % 3]
        v3 = cells[a0, (k0 + 2) % 3]
        cells[a0] = np.stack([v0, v1, v2], axis=1)
        cells[a1] = np.stack([v0, v1, v3], axis=1)
        self._twin = None
        self._geo = None
        return int(h0.size)

    def flip_until_delaunay(self, tol=0.0, max_steps=100):
        """A.7.  Returns (number of flips, number of rounds that flipped)."""
        total = 0
        rounds = 0
        for _ in range(max_steps):
            n = self.flip_round(tol)
            if n == 0:
                return total, rounds
            total += n
            rounds += 1
        if self.num_delaunay_violations(tol) > 0:
            warnings.warn("Maximum number of edge flips reached.")
        return total, rounds

    def num_delaunay_violations(self, tol=0.0):
        twin = self.twins
        ceh = self.ce_ratios.T.reshape(-1)
        interior = twin >= 0
        s = ceh[interior] + ceh[twin[interior]]
        return int(np.count_nonzero(s < -tol) // 2)


def canonical_cells(cells):
    """Rows sorted, then rows lexsorted: topology independent of slot order/row order."""
    c = np.sort(np.asarray(cells, dtype=np.int64), axis=1)
    order = np.lexsort((c[:, 2], c[:, 1], c[:, 0]))
    return c[order]
