"""Host model of the step kernel's arithmetic (test infrastructure).

`optimesh_b200/csrc/chain.cuh` evaluates the star of a vertex as a chain of SPOKES
d_q = x[n_q] - x[v] (n_0 .. n_{k-1}: the one-ring in walk order).  Cell q lies between the
spokes q and q+1; everything a cell contributes to its vertex is a combination of those two
spokes, so the per-vertex sums of SURVEY.md A.4/A.8/A.9 regroup into one coefficient per
spoke fed by the two cells next to it.  This file restates that regrouping in plain Python,
one vertex at a time, so that `tests/test_chain_model.py` can check it against the oracle on
the CPU (no GPU needed): a wrong sign or factor in the regrouping shows up here, before the
CUDA transcription is ever run.

Scaling (division free, one rsqrt per cell): with V4 = L_q L_{q+1} - c_q^2 = (2A)^2 and
rs = 1/sqrt(V4), t''_k = ed_k rs = 2 t_k where t_k = ed_k / (4A) = -ce_k (A.2).
"""
from __future__ import annotations

import numpy as np


def _hi(x):
    """High 32 bits of a positive double, as the kernel reads them (__double2hiint)."""
    return int(np.float64(x).view(np.int64) >> 32)


def _from_hi(h):
    return float(np.int64(int(h) << 32).view(np.float64))


def vertex_rings(cells, n):
    """Ordered one-rings: for every vertex with a CLOSED fan the neighbour ids in walk order."""
    star = [[] for _ in range(n)]
    for c, tri in enumerate(cells):
        for j in range(3):
            star[tri[j]].append((tri[(j + 1) % 3], tri[(j + 2) % 3]))
    rings = {}
    for v in range(n):
        pairs = star[v]
        if not pairs:
            continue
        # neighbour -> the two cells (as unordered pairs) it belongs to
        nxt = {}
        for a, b in pairs:
            nxt.setdefault(a, []).append(b)
            nxt.setdefault(b, []).append(a)
        if any(len(w) != 2 for w in nxt.values()):
            continue  # open fan (boundary vertex)
        start = pairs[0][0]
        ring = [start, pairs[0][1]]
        while True:
            a, b = nxt[ring[-1]]
            new = b if a == ring[-2] else a
            if new == ring[0]:
                break
            ring.append(new)
        if len(ring) == len(pairs):
            rings[v] = ring
    return rings


def chain_vertex(P0, R, method, omega=1.0, bary=None):
    """One vertex: P0 (d,), ring coordinates R (k, d) in walk order.

    Returns dict(d=offset of the relaxed, unlimited update, rmin=smallest incident inradius,
    min_q (lazy limiter bound, a difference of high words), flags=[suspicious spoke q], degenerate)."""
    k = R.shape[0]
    d = R - P0
    L = np.einsum("ij,ij->i", d, d)
    dim = P0.shape[0]
    W = 0.0
    NUM = np.zeros(dim)
    H = np.zeros((dim, dim))
    t1 = np.zeros(k)  # t''_1 of cell q: angle at n_q (opposite spoke q+1)
    t2 = np.zeros(k)  # t''_2 of cell q: angle at n_{q+1} (opposite spoke q)
    s1 = np.zeros(k)
    s2 = np.zeros(k)
    t1raw = np.zeros(k)
    t2raw = np.zeros(k)
    masked = np.zeros(k, dtype=bool)
    mflag = set()
    rmin_num, rmin_den = np.inf, 1.0
    min_q = 0x7fffffff
    lens = np.sqrt(L)
    for q in range(k):
        r = (q + 1) % k
        c = float(d[q] @ d[r])
        V4 = L[q] * L[r] - c * c
        if not V4 > 0.0:
            return dict(degenerate=True)
        rs = 1.0 / np.sqrt(V4)
        # limiter
        min_q = min(min_q, _hi(V4) - max(_hi(L[q]), _hi(L[r])))
        A2 = V4 * rs
        per = lens[q] + lens[r] + np.sqrt(L[q] + L[r] - 2.0 * c)
        if A2 * rmin_den < rmin_num * per:
            rmin_num, rmin_den = A2, per
        m0, m1, m2 = -c, c - L[q], c - L[r]
        T0, T1, T2 = m0 * rs, m1 * rs, m2 * rs
        t1raw[q], t2raw[q] = T1, T2
        # a cell with |T| >= 2^29 (T hugely negative: an angle below 4e-9 rad) names both its
        # spokes: the absolute threshold of the check below does not cover its rounding
        if min(T1, T2) <= -2.0 ** 29:
            mflag.add(q)
            mflag.add(r)
        if method in ("lloyd", "cvt-block-diagonal"):
            if max(T0, T1, T2) > 1.0:
                masked[q] = True
                # the masked cell names the spoke opposite its > 135 deg angle itself
                if T2 > 1.0:
                    mflag.add(q)
                if T1 > 1.0:
                    mflag.add(r)
                continue
            w1, w2 = L[r] * T1, L[q] * T2
            uu = rs * (w1 + w2)
            t1[q], t2[q] = T1, T2
            s2[q] = -uu * w1 + w2
            s1[q] = -uu * w2 + w1
        elif method == "cpt-fixed-point":
            W += A2
            s1[q] = s2[q] = A2
        elif method in ("odt-fixed-point", "odt-dp-fp"):
            dp = method == "odt-dp-fp"
            W += 1.0 if dp else A2
            if bary is not None and bary[q]:
                s1[q] = s2[q] = 1.0 if dp else A2
            else:
                f = -1.5 * (rs if dp else 1.0)
                s2[q] = f * L[r] * T1
                s1[q] = f * L[q] * T2
        else:
            raise KeyError(method)
    flags = []
    for q in range(k):
        p = (q - 1) % k
        cH = t2[q] + t1[p]
        cN = s2[q] + s1[p]
        if method in ("lloyd", "cvt-block-diagonal"):
            W += L[q] * cH
            H += cH * np.outer(d[q], d[q])
        NUM += cN * d[q]
        # the kernel sees the masked (zeroed) t: conservative by the argument in chain.cuh;
        # kept: cH >= 0, or negative with a high word of at most that of -2^-20
        if q in mflag or cH >= 0.0 or _hi(-cH) <= 0x3EB00000:
            flags.append(q)
    if W == 0.0:
        off = np.zeros(dim)
    elif method == "lloyd":
        off = NUM / (6.0 * W)
    elif method == "cvt-block-diagonal":
        M = W * np.eye(dim) - H
        off = np.linalg.solve(M, NUM / 6.0) if np.linalg.det(M) != 0.0 else np.zeros(dim)
    else:
        off = NUM / (3.0 * W)
    return dict(d=omega * off, rmin=rmin_num / rmin_den, min_q=min_q,
                flags=flags, degenerate=False)


def step_model(points, cells, method, omega=1.0, is_boundary=None, boundary_cells=None):
    """The update of every free vertex with a closed fan (boundary vertices pinned), with the
    exact limiter.  Returns (new points, max diff^2, n_limited, lazy_ok, flagged edges)."""
    X = np.asarray(points, dtype=np.float64)
    n = X.shape[0]
    rings = vertex_rings(cells, n)
    new = X.copy()
    max_diff2 = 0.0
    n_limited = 0
    lazy_ok = True
    flagged = set()
    bc_lookup = None
    if boundary_cells is not None:
        bc_lookup = {tuple(sorted(cells[c])) for c in np.nonzero(boundary_cells)[0]}
    for v, ring in rings.items():
        if is_boundary is not None and is_boundary[v]:
            continue
        bary = None
        if bc_lookup is not None:
            k = len(ring)
            bary = [tuple(sorted((v, ring[q], ring[(q + 1) % k]))) in bc_lookup for q in range(k)]
        out = chain_vertex(X[v], X[ring], method, omega, bary)
        assert not out["degenerate"]
        d = out["d"]
        diff2 = float(d @ d)
        max_diff2 = max(max_diff2, diff2)
        limited = np.sqrt(diff2) > 0.5 * out["rmin"]
        # the lazy bound must never declare a limited vertex "not limited"
        dq = _from_hi(out["min_q"] - 1 + 0x3FF00000)
        proves_free = 81.2 * diff2 <= dq
        if proves_free and limited:
            lazy_ok = False
        if limited:
            d = d * (0.5 * out["rmin"] / np.sqrt(diff2))
            n_limited += 1
        new[v] = X[v] + d
        for q in out["flags"]:
            flagged.add((min(v, ring[q]), max(v, ring[q])))
    return new, max_diff2, n_limited, lazy_ok, flagged
