"""Synthetic input meshes (host side, numpy/scipy; not on the hot path).

The reference's examples use a "randomly generated disk mesh"
(/root/reference/README.md:70-74) and ``meshzoo.tetra_sphere(20)`` (README.md:151-154);
neither generator is available offline, so the workloads named in BASELINE.json are
produced here.
"""
from __future__ import annotations

import numpy as np


def disk(nb: int, seed: int = 0):
    """Random disk mesh: ``nb`` equispaced boundary points on the unit circle, uniform
    random interior points, triangulated by Qhull.  disk(120) -> 1,383 / 2,643."""
    import scipy.spatial

    h = 2 * np.pi / nb
    a_dom = np.pi - nb * 0.5 * (h - np.sin(h))
    a_cell = np.sqrt(3.0) / 4.0 * h * h
    m = int(0.5 * (a_dom / a_cell + nb) + 1 - nb)
    rs = np.random.RandomState(seed)
    u1 = rs.rand(m)
    u2 = rs.rand(m)
    t = 2 * np.pi * np.arange(nb) / nb
    bnd = np.stack([np.cos(t), np.sin(t)], axis=1)
    r = np.sqrt(u1)
    inner = np.stack([r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)], axis=1)
    pts = np.concatenate([bnd, inner])
    cells = scipy.spatial.Delaunay(pts).simplices.astype(np.int64)
    return np.ascontiguousarray(pts), np.ascontiguousarray(cells)


def square(n: int, jitter: float = 0.25, seed: int = 0, shuffle: bool = False):
    """n x n grid on [0,1]^2, each quad split along the same diagonal; interior vertices
    jittered uniformly by +-jitter*h per coordinate (jitter < 0.3 keeps all cells valid)."""
    h = 1.0 / (n - 1)
    g = np.arange(n, dtype=np.float64) * h
    X, Y = np.meshgrid(g, g, indexing="xy")
    pts = np.stack([X.reshape(-1), Y.reshape(-1)], axis=1)
    rs = np.random.RandomState(seed)
    jit = (rs.rand(n * n, 2) * 2.0 - 1.0) * (jitter * h)
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    interior = ((ii > 0) & (ii < n - 1) & (jj > 0) & (jj < n - 1)).reshape(-1)
    pts[interior] += jit[interior]
    cells = _grid_cells(n)
    if shuffle:
        pts, cells = shuffle_vertices(pts, cells, seed)
    return pts, cells


def _grid_cells(n: int):
    i, j = np.meshgrid(np.arange(n - 1, dtype=np.int64), np.arange(n - 1, dtype=np.int64),
                       indexing="xy")
    a = (j * n + i).reshape(-1)
    b = a + 1
    c = a + n
    d = c + 1
    lower = np.stack([a, b, d], axis=1)
    upper = np.stack([a, d, c], axis=1)
    return np.ascontiguousarray(np.stack([lower, upper], axis=1).reshape(-1, 3))


def disk_mapped_grid(n: int, jitter: float = 0.25, seed: int = 0, shuffle: bool = False):
    """Disk mesh for sizes Qhull cannot reach on the host (10M+ vertices): a jittered
    n x n grid on [-1,1]^2 pushed to the unit disk by the elliptical-grid map
    (u, v) = (x sqrt(1 - y^2/2), y sqrt(1 - x^2/2)).  The result is a valid but
    non-Delaunay triangulation; the first flip-until-Delaunay pass repairs it."""
    pts, cells = square(n, jitter, seed)
    x = 2.0 * pts[:, 0] - 1.0
    y = 2.0 * pts[:, 1] - 1.0
    u = x * np.sqrt(1.0 - 0.5 * y * y)
    v = y * np.sqrt(1.0 - 0.5 * x * x)
    # boundary vertices exactly on the unit circle
    bnd = (np.abs(x) == 1.0) | (np.abs(y) == 1.0)
    rr = np.sqrt(u[bnd] ** 2 + v[bnd] ** 2)
    u[bnd] /= rr
    v[bnd] /= rr
    pts = np.ascontiguousarray(np.stack([u, v], axis=1))
    if shuffle:
        pts, cells = shuffle_vertices(pts, cells, seed)
    return pts, cells


def tetra_sphere(n: int):
    """Each face of a regular tetrahedron split into n^2 triangles, shared vertices
    merged (N = 2 n^2 + 2, C = 4 n^2), all vertices normalised to the unit sphere."""
    corners = np.array(
        [[1.0, 1.0, 1.0], [1.0, -1.0, -1.0], [-1.0, 1.0, -1.0], [-1.0, -1.0, 1.0]]
    ) / np.sqrt(3.0)
    faces = [(0, 1, 2), (0, 3, 1), (0, 2, 3), (1, 3, 2)]
    i, j = np.meshgrid(np.arange(n + 1, dtype=np.int64), np.arange(n + 1, dtype=np.int64),
                       indexing="ij")
    ok = (i + j) <= n
    i, j = i[ok], j[ok]
    k = n - i - j
    # local index of grid point (i, j) inside one face
    lid = -np.ones((n + 1, n + 1), dtype=np.int64)
    lid[i, j] = np.arange(i.size)
    # local cells
    iu, ju = np.meshgrid(np.arange(n, dtype=np.int64), np.arange(n, dtype=np.int64), indexing="ij")
    up = (iu + ju) <= n - 1
    a_i, a_j = iu[up], ju[up]
    up_cells = np.stack([lid[a_i, a_j], lid[a_i + 1, a_j], lid[a_i, a_j + 1]], axis=1)
    dn = (iu + ju) <= n - 2
    b_i, b_j = iu[dn], ju[dn]
    dn_cells = np.stack([lid[b_i + 1, b_j], lid[b_i + 1, b_j + 1], lid[b_i, b_j + 1]], axis=1)
    local_cells = np.concatenate([up_cells, dn_cells])
    keys = []
    cells = []
    m = i.size
    base = np.int64(n + 1)
    for f, (a, b, c) in enumerate(faces):
        w = np.zeros((m, 4), dtype=np.int64)
        w[:, a] = k
        w[:, b] = i
        w[:, c] = j
        keys.append(((w[:, 0] * base + w[:, 1]) * base + w[:, 2]) * base + w[:, 3])
        cells.append(local_cells + f * m)
    keys = np.concatenate(keys)
    cells = np.concatenate(cells)
    uniq, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    w3 = uniq % base
    w2 = (uniq // base) % base
    w1 = (uniq // (base * base)) % base
    w0 = uniq // (base * base * base)
    W = np.stack([w0, w1, w2, w3], axis=1).astype(np.float64) / n
    pts = W @ corners
    pts /= np.sqrt((pts * pts).sum(axis=1))[:, None]
    cells = inv.reshape(-1)[cells]
    return np.ascontiguousarray(pts), np.ascontiguousarray(cells.astype(np.int64))


def shuffle_vertices(points, cells, seed: int = 0):
    """Random vertex renumbering (exercises the spatial renumbering in setup)."""
    rs = np.random.RandomState(seed + 12345)
    perm = rs.permutation(points.shape[0])  # new -> old
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    return np.ascontiguousarray(points[perm]), np.ascontiguousarray(inv[cells])


SIMPLE1 = (
    np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0], [0.4, 0.5]]),
    np.array([[0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4]]),
)


def disk_mapped_grid_torch(n: int, jitter: float = 0.25, seed: int = 0, device="cuda"):
    """`disk_mapped_grid` built on the device with torch (same construction, torch's RNG):
    float64 (n*n, 2) points and int32 (2 (n-1)^2, 3) cells as CUDA tensors.  For meshes that
    are too large to build on the host within the bench's time budget."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    h = 1.0 / (n - 1)
    g = torch.arange(n, dtype=torch.float64, device=device) * h
    X = g.repeat(n)                      # x varies fastest
    Y = g.repeat_interleave(n)
    ii = torch.arange(n, device=device).repeat(n)
    jj = torch.arange(n, device=device).repeat_interleave(n)
    interior = (ii > 0) & (ii < n - 1) & (jj > 0) & (jj < n - 1)
    jit = (torch.rand(n * n, 2, dtype=torch.float64, device=device, generator=gen) * 2 - 1) \
        * (jitter * h)
    X = X + jit[:, 0] * interior
    Y = Y + jit[:, 1] * interior
    del jit
    x = 2.0 * X - 1.0
    y = 2.0 * Y - 1.0
    u = x * torch.sqrt(1.0 - 0.5 * y * y)
    v = y * torch.sqrt(1.0 - 0.5 * x * x)
    bnd = ~interior
    rr = torch.sqrt(u[bnd] ** 2 + v[bnd] ** 2)
    u[bnd] = u[bnd] / rr
    v[bnd] = v[bnd] / rr
    pts = torch.stack([u, v], dim=1).contiguous()
    i = torch.arange(n - 1, device=device, dtype=torch.int32).repeat(n - 1)
    j = torch.arange(n - 1, device=device, dtype=torch.int32).repeat_interleave(n - 1)
    a = j * n + i
    b = a + 1
    c = a + n
    d = c + 1
    lower = torch.stack([a, b, d], dim=1)
    upper = torch.stack([a, d, c], dim=1)
    cells = torch.stack([lower, upper], dim=1).reshape(-1, 3).contiguous()
    return pts, cells


def disk_gpu(n: int, rounds: int = 120, seed: int = 0, device: int = 0, stream=None,
             jitter: float = 0.25):
    """Random disk mesh for sizes Qhull cannot reach on the host ("a randomly generated disk
    mesh", README.md:70-74): the mapped grid `disk_mapped_grid(n)` is
    built on the device, then `rounds` random moves of every interior vertex, each bounded by
    half the smallest incident inradius (no cell can invert) and followed by
    flip-until-Delaunay, turn it into a random Delaunay triangulation of the unit disk
    (`om_random_walk`).  120 rounds give the vertex-degree histogram of `disk()`: about 4 % of
    the interior vertices have more than 8 cells.  Returns the resident `DeviceMesh`."""
    import torch

    from .mesh import DeviceMesh

    with torch.cuda.device(device):
        tp, tc = disk_mapped_grid_torch(n, jitter, seed, device=f"cuda:{device}")
        torch.cuda.synchronize()
        dm = DeviceMesh.from_torch(tp, tc, stream=stream)
        del tp, tc
        torch.cuda.empty_cache()
    dm.random_walk(rounds, seed)
    return dm
