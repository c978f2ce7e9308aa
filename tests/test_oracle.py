"""Pins the CPU oracle (no GPU needed).

The reference tree holds no tests or golden vectors, so the oracle
is pinned by: the recollected upstream known-answer literals on the 5-point "simple1"
mesh, Qhull as an independent topology oracle, scipy's direct solver, and invariants.
"""
import json
import os

import numpy as np
import pytest
import scipy.spatial

import oracle
from oracle.meshtri import MeshTri, canonical_cells, DegenerateCellsError
from optimesh_b200 import generators as G

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def norms(p):
    return np.abs(p).sum(), np.sqrt((p * p).sum()), np.abs(p).max()


def test_simple1_lloyd_known_answer():
    X, cells = G.SIMPLE1
    p, c = oracle.optimize_points_cells(X, cells, "lloyd", 1.0e-2, 100)
    n1, n2, ninf = norms(p)
    # upstream literal, rel. tol 1e-12 as upstream
    assert abs(n1 - 4.986335452622451) <= 1e-12 * n1
    assert abs(n2 - 2.1181412069258942) <= 1e-12 * n2
    assert ninf == 1.0


def test_simple1_lloyd_iterates():
    X, cells = G.SIMPLE1
    expect = [0.43194444, 0.45420143, 0.46933845, 0.47952039, 0.48633545]
    mesh = MeshTri(X, cells)
    for k in range(5):
        oracle.driver.step(mesh, "lloyd")
        assert abs(mesh.points[4, 0] - expect[k]) < 5e-9
        assert abs(mesh.points[4, 1] - 0.5) < 1e-15


@pytest.mark.parametrize("method", ["cpt-fixed-point", "cpt-linear-solve", "odt-fixed-point"])
def test_simple1_cpt_odt_known_answer(method):
    X, cells = G.SIMPLE1
    p, c = oracle.optimize_points_cells(X, cells, method, 1.0e-2, 100)
    n1, n2, ninf = norms(p)
    assert abs(n1 - 5.0) < 1e-12
    assert abs(n2 - 2.1213203435596424) < 1e-12
    assert ninf == 1.0


def test_simple1_block_diagonal_converges_faster():
    X, cells = G.SIMPLE1
    steps = {}
    for m in ("lloyd", "cvt-block-diagonal"):
        mesh = MeshTri(X, cells)
        steps[m] = [oracle.optimize(MeshTri(X, cells), m, tol, 100) for tol in (1e-2, 1e-3, 1e-5)]
    assert steps["lloyd"] == [5, 10, 22]
    assert steps["cvt-block-diagonal"] == [3, 5, 10]


def test_method_name_normaliser():
    assert oracle.normalize_method_name("CVT (block-diagonal)") == "cvt-block-diagonal"
    assert oracle.normalize_method_name("Lloyd") == "lloyd"
    assert oracle.normalize_method_name("CVT (full)") == "cvt-full"
    with pytest.raises(NotImplementedError):
        oracle.get_new_points(MeshTri(*G.SIMPLE1), "CVT (full)")
    with pytest.raises(KeyError):
        oracle.get_new_points(MeshTri(*G.SIMPLE1), "nope")


def test_control_volumes_sum_to_area():
    pts, cells = G.disk(60, 1)
    m = MeshTri(pts, cells)
    assert abs(m.control_volumes.sum() - m.cell_volumes.sum()) < 1e-12
    assert np.allclose(2 * m.cell_partitions.sum(axis=0), m.cell_volumes, rtol=1e-12)


def test_circumcenter_equidistant_and_quality():
    pts, cells = G.disk(40, 2)
    m = MeshTri(pts, cells)
    cc = m.cell_circumcenters
    d = np.stack([np.linalg.norm(pts[cells[:, k]] - cc, axis=1) for k in range(3)])
    assert np.allclose(d[0], d[1], rtol=1e-9) and np.allclose(d[0], d[2], rtol=1e-9)
    assert np.allclose(d[0], m.cell_circumradius, rtol=1e-9)
    assert np.allclose(m.q_radius_ratio, 2 * m.cell_inradius / m.cell_circumradius, rtol=1e-10)
    eq = MeshTri(np.array([[0, 0], [1, 0], [0.5, np.sqrt(3) / 2]]), np.array([[0, 1, 2]]))
    assert abs(eq.q_radius_ratio[0] - 1.0) < 1e-14
    assert np.allclose(eq.angles, np.pi / 3)


def test_hex_patch_is_fixed_point():
    # regular hexagonal patch: the centre vertex is a fixed point of every method
    t = np.arange(6) * np.pi / 3
    pts = np.concatenate([[[0.0, 0.0]], np.stack([np.cos(t), np.sin(t)], axis=1)])
    cells = np.array([[0, 1 + k, 1 + (k + 1) % 6] for k in range(6)])
    for meth in oracle.METHODS:
        new = oracle.get_new_points(MeshTri(pts, cells), meth)
        assert np.allclose(new[0], 0.0, atol=1e-14), meth


def test_odt_boundary_cells_use_barycenters(monkeypatch):
    """Cells with a boundary edge contribute their barycenter to the ODT updates."""
    import oracle.methods as om

    pts, cells = G.disk(40, 3)
    mesh = MeshTri(pts, cells)
    bc = mesh.is_boundary_cell
    # a cell is a boundary cell iff one of its edges has no twin
    assert bc.sum() == (mesh.twins < 0).sum() == 40
    ref = mesh.cell_circumcenters.copy()
    ref[bc] = mesh.cell_barycenters[bc]
    assert np.array_equal(om.odt_fixed_point(mesh), om._volume_averaged(mesh, ref))
    assert np.array_equal(om.odt_dp_fp(mesh), om._count_averaged(mesh, ref))
    with_sub = om.odt_fixed_point(mesh)
    monkeypatch.setattr(om, "ODT_BOUNDARY_BARYCENTERS", False)
    plain = om.odt_fixed_point(mesh)
    assert np.array_equal(plain, om._volume_averaged(mesh, mesh.cell_circumcenters))
    # only vertices of boundary cells are affected
    touched = np.zeros(mesh.n, dtype=bool)
    touched[cells[bc].reshape(-1)] = True
    assert np.array_equal(with_sub[~touched], plain[~touched])
    assert not np.allclose(with_sub[touched & mesh.is_interior_point],
                           plain[touched & mesh.is_interior_point])


def test_odt_dp_fp_simple1():
    # every cell of simple1 is a boundary cell: the interior vertex moves to
    # x/3 + 2/3 (0.5, 0.5) per step (mean of the four barycenters)
    X, cells = G.SIMPLE1
    new = oracle.get_new_points(MeshTri(X, cells), "odt-dp-fp")
    assert np.allclose(new[4], X[4] / 3 + np.array([1.0, 1.0]) / 3, atol=1e-15)
    p, _ = oracle.optimize_points_cells(X, cells, "odt-dp-fp", 1.0e-6, 100)
    assert np.allclose(p[4], [0.5, 0.5], atol=1e-6)


def test_cpt_quasi_newton_matrix_properties():
    """The approximate Hessian is strictly diagonally dominant (off-diagonal row sum = 2/3 of
    the diagonal), hence SPD and well conditioned; on a single free vertex the step equals
    the CPT fixed-point step; it needs fewer steps than the fixed-point iteration."""
    X, cells = G.SIMPLE1
    qn = oracle.get_new_points(MeshTri(X, cells), "cpt-quasi-newton")
    fp = oracle.get_new_points(MeshTri(X, cells), "cpt-fixed-point")
    assert np.allclose(qn, fp, atol=1e-15)
    pts, cells = G.disk(40, 3)
    mesh = MeshTri(pts, cells)
    new = oracle.get_new_points(mesh, "cpt-quasi-newton")
    assert np.array_equal(new[mesh.is_boundary_point], pts[mesh.is_boundary_point])
    # residual of the defining system on interior rows: H (x - new) = dE
    vol, bary = mesh.cell_volumes, mesh.cell_barycenters
    step = pts - new
    res = np.zeros(pts.shape)
    for k in range(3):
        i = cells[:, k]
        np.add.at(res, i, (2 / 3) * vol[:, None] * step[i])
        for kk in ((k + 1) % 3, (k + 2) % 3):
            np.add.at(res, i, -(2 / 9) * vol[:, None] * step[cells[:, kk]])
        np.add.at(res, i, -(2 / 3) * vol[:, None] * (pts[i] - bary))
    inner = mesh.is_interior_point
    assert np.abs(res[inner]).max() < 1e-14
    n_qn, n_fp = [], []
    oracle.optimize_points_cells(pts, cells, "cpt-quasi-newton", 1.0e-6, 200, log=n_qn)
    oracle.optimize_points_cells(pts, cells, "cpt-fixed-point", 1.0e-6, 200, log=n_fp)
    assert len(n_qn) < len(n_fp)


@pytest.mark.parametrize("method", sorted(oracle.METHODS))
def test_methods_are_similarity_and_relabelling_equivariant(method):
    """new(s R x + t) = s R new(x) + t, and renumbering vertices / reordering or reorienting
    cells does not change the geometry of the result: the restated formulas depend on the
    mesh only, not on coordinates or labels."""
    pts, cells = G.disk(30, 5)
    base = oracle.get_new_points(MeshTri(pts, cells), method)
    th, s, t = 0.7, 2.5, np.array([3.0, -1.5])
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    moved = oracle.get_new_points(MeshTri(s * pts @ R.T + t, cells), method)
    assert np.allclose(moved, s * base @ R.T + t, rtol=0, atol=5e-12)
    rs = np.random.RandomState(11)
    perm = rs.permutation(len(pts))           # new label of old vertex i is inv[i]
    inv = np.argsort(perm)
    c2 = inv[cells][rs.permutation(len(cells))]
    flip = rs.rand(len(c2)) < 0.5             # mixed orientation, rotated slots
    c2[flip] = c2[flip][:, ::-1]
    c2 = np.where((rs.rand(len(c2)) < 0.5)[:, None], np.roll(c2, 1, axis=1), c2)
    relabelled = oracle.get_new_points(MeshTri(pts[perm], c2), method)
    assert np.allclose(relabelled, base[perm], rtol=0, atol=5e-13)


def test_odt_update_is_the_gradient_of_the_odt_energy(monkeypatch):
    """Chen-Holst (README.md:244-251): with the triangulation fixed, the ODT energy is
    E = 1/(d+1) sum_i |x_i|^2 |w_i| + const, and dE/dx_i = 2/(d+1) |w_i| (x_i - x_i*) where
    x_i* is the area-weighted mean of the circumcenters of the star.  Checked by central
    differences on an energy evaluated independently of the oracle (d = 2)."""
    import oracle.methods as om

    monkeypatch.setattr(om, "ODT_BOUNDARY_BARYCENTERS", False)  # the formula of the paper
    pts, cells = G.disk(24, 9)

    def areas(x):
        a, b = x[cells[:, 1]] - x[cells[:, 0]], x[cells[:, 2]] - x[cells[:, 0]]
        return 0.5 * np.abs(a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0])

    def energy(x):
        star = np.zeros(len(x))
        np.add.at(star, cells.reshape(-1), np.repeat(areas(x), 3))
        return np.sum(np.einsum("ij,ij->i", x, x) * star) / 3.0

    mesh = MeshTri(pts, cells)
    target = om.odt_fixed_point(mesh)
    star = np.zeros(len(pts))
    np.add.at(star, cells.reshape(-1), np.repeat(areas(pts), 3))
    inner = np.nonzero(mesh.is_interior_point)[0]
    h = 1.0e-6
    for i in inner[:: max(1, len(inner) // 12)]:
        for k in range(2):
            xp, xm = pts.copy(), pts.copy()
            xp[i, k] += h
            xm[i, k] -= h
            fd = (energy(xp) - energy(xm)) / (2 * h)
            want = (2.0 / 3.0) * star[i] * (pts[i, k] - target[i, k])
            assert abs(fd - want) <= 1e-8 * max(1.0, abs(want)), (i, k, fd, want)


def test_control_volumes_are_voronoi_cells():
    """A.4 against an independent construction: on a Delaunay mesh the control volume of an
    interior vertex whose star has no boundary cell is its Voronoi cell (Qhull via
    scipy.spatial.Voronoi); areas and centroids must agree, and with them the Lloyd target and
    the gradient 2 |V_i| (x_i - c_i) of the CVT energy (Du-Faber-Gunzburger)."""
    from scipy.spatial import Voronoi

    pts, cells = G.disk(40, 2)
    mesh = MeshTri(pts, cells)
    mask = np.any(mesh.ce_ratios < -0.5, axis=0)
    cv = mesh.get_control_volumes(cell_mask=mask)
    cen = mesh.get_control_volume_centroids(cell_mask=mask)
    touched = np.zeros(mesh.n, dtype=bool)  # vertices of boundary cells or masked cells
    touched[cells[mesh.is_boundary_cell | mask].reshape(-1)] = True
    vor = Voronoi(pts)
    checked = 0
    for i in np.nonzero(mesh.is_interior_point & ~touched)[0]:
        region = vor.regions[vor.point_region[i]]
        if not region or -1 in region:
            continue
        poly = vor.vertices[region]
        if np.linalg.norm(poly, axis=1).max() > 0.98:  # leaves the meshed domain
            continue
        # order the polygon around its vertex, then shoelace area and centroid
        ang = np.arctan2(poly[:, 1] - pts[i, 1], poly[:, 0] - pts[i, 0])
        poly = poly[np.argsort(ang)]
        x, y = poly[:, 0], poly[:, 1]
        xn, yn = np.roll(x, -1), np.roll(y, -1)
        cross = x * yn - xn * y
        area = 0.5 * cross.sum()
        cx = ((x + xn) * cross).sum() / (6 * area)
        cy = ((y + yn) * cross).sum() / (6 * area)
        assert abs(cv[i] - area) <= 1e-12 * area
        assert np.allclose(cen[i], [cx, cy], rtol=0, atol=1e-12)
        checked += 1
    assert checked > 50
    lloyd = oracle.get_new_points(mesh, "lloyd")
    assert np.array_equal(lloyd[mesh.is_interior_point], cen[mesh.is_interior_point])


def test_degenerate_cell_raises():
    pts = np.array([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0]])
    with pytest.raises(DegenerateCellsError):
        MeshTri(pts, np.array([[0, 1, 2]])).cell_volumes


@pytest.mark.parametrize("seed", range(6))
def test_flips_reach_qhull_delaunay(seed):
    # Delaunay mesh, interior points jittered -> flips must restore Qhull's triangulation
    pts, cells = G.disk(60, seed)
    pts, cells = oracle.optimize_points_cells(pts, cells, "cpt-fixed-point", 0.0, 5)
    rs = np.random.RandomState(seed)
    mesh = MeshTri(pts, cells)
    bnd = mesh.is_boundary_point
    pts2 = pts.copy()
    rmin = np.full(len(pts), np.inf)
    np.minimum.at(rmin, cells.reshape(-1), np.repeat(mesh.cell_inradius, 3))
    step = rs.uniform(-1, 1, size=pts.shape) * (0.45 * rmin / np.sqrt(2))[:, None]
    pts2[~bnd] += step[~bnd]
    mesh = MeshTri(pts2, cells)
    assert np.all(_signed_areas(pts2, cells) * np.sign(_signed_areas(pts, cells)) > 0)
    nflips, nrounds = mesh.flip_until_delaunay()
    assert nflips > 0
    assert mesh.num_delaunay_violations() == 0
    ref = scipy.spatial.Delaunay(pts2).simplices
    assert np.array_equal(canonical_cells(mesh.cells("points")), canonical_cells(ref))


def _signed_areas(p, c):
    a, b, cc = p[c[:, 0]], p[c[:, 1]], p[c[:, 2]]
    return 0.5 * ((b[:, 0] - a[:, 0]) * (cc[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (cc[:, 0] - a[:, 0]))


def test_sphere_flips_match_convex_hull():
    pts, cells = G.tetra_sphere(8)
    assert pts.shape[0] == 2 * 64 + 2 and cells.shape[0] == 4 * 64
    rs = np.random.RandomState(0)
    p2 = pts + rs.normal(scale=0.03, size=pts.shape)
    p2 /= np.linalg.norm(p2, axis=1)[:, None]
    mesh = MeshTri(p2, cells)
    mesh.flip_until_delaunay()
    hull = scipy.spatial.ConvexHull(p2).simplices
    assert np.array_equal(canonical_cells(mesh.cells("points")), canonical_cells(hull))


def test_cpt_linear_solve_is_harmonic():
    pts, cells = G.square(12, 0.25, 0)
    mesh = MeshTri(pts, cells)
    new = oracle.get_new_points(mesh, "cpt-linear-solve")
    bnd = mesh.is_boundary_point
    assert np.allclose(new[bnd], pts[bnd], atol=1e-14)
    # interior vertices are the mean of their neighbours
    nbrs = [set() for _ in range(len(pts))]
    for a, b, c in cells:
        nbrs[a] |= {b, c}
        nbrs[b] |= {a, c}
        nbrs[c] |= {a, b}
    for i in np.nonzero(~bnd)[0]:
        assert np.allclose(new[i], new[list(nbrs[i])].mean(axis=0), atol=1e-12)


def test_generators_sizes():
    pts, cells = G.disk(120, 0)
    assert pts.shape == (1383, 2) and cells.shape == (2643, 3)
    pts, cells = G.square(10)
    assert pts.shape == (100, 2) and cells.shape == (162, 3)
    assert np.all(_signed_areas(pts, cells) > 0)
    pts, cells = G.disk_mapped_grid(40)
    assert np.all(_signed_areas(pts, cells) > 0)
    m = MeshTri(pts, cells)
    assert m.is_boundary_point.sum() == 4 * 39
    assert np.allclose(np.linalg.norm(pts[m.is_boundary_point], axis=1), 1.0)
    pts, cells = G.tetra_sphere(5)
    m = MeshTri(pts, cells)
    assert not m.is_boundary_point.any()
    assert abs(m.cell_volumes.sum() - 4 * np.pi) < 0.15 * 4 * np.pi


def test_config1_trajectory_golden():
    """Config 1 (BASELINE.json configs[0]): Lloyd omega=1, disk(120), 50 steps, tol 1e-5.
    Golden produced by tests/golden/make_golden.py from this oracle."""
    with open(os.path.join(GOLDEN, "config1_lloyd.json")) as f:
        gold = json.load(f)
    pts, cells = G.disk(120, 0)
    log = []
    p, c = oracle.optimize_points_cells(pts, cells, "lloyd", 1.0e-5, 50, log=log)
    assert len(log) == gold["steps"] == 50
    assert [l["n_flips"] for l in log] == gold["n_flips"]
    assert [l["n_limited"] for l in log] == gold["n_limited"]
    assert np.allclose(norms(p), gold["norms"], rtol=1e-12)
    ah, qh, s = oracle.stats(MeshTri(p, c))
    assert ah.tolist() == gold["angle_hist"] and qh.tolist() == gold["q_hist"]


def test_survey_flip_rule_reaches_the_same_triangulation():
    """The per-round flip selection is this build's own rule ("an edge flips iff it is the most
    negative flagged edge of BOTH its cells"); SURVEY.md A.7 words the upstream rule differently
    ("while a cell touches two flagged edges it keeps its most negative one").  Neither can be
    pinned against the reference, so row-for-row cell equality between the GPU and this oracle
    is agreement on OUR rule.  What is independent of the rule is the fixed point: both rules
    end in the same (canonical) triangulation -- Qhull's."""
    import scipy.spatial

    from oracle.meshtri import MeshTri, canonical_cells
    from optimesh_b200 import generators as G

    rs = np.random.RandomState(11)
    done = 0
    for trial in range(8):
        pts, cells = G.disk(int(rs.randint(20, 70)), int(rs.randint(0, 100)))
        m0 = MeshTri(pts, cells)
        bnd = m0.is_boundary_point
        # random moves bounded by 0.45 x the smallest incident inradius (no cell can invert),
        # a few of them in a row so that plenty of edges stop being Delaunay
        moved = pts.copy()
        for _ in range(4):
            mm = MeshTri(moved, cells)
            rmin = np.full(len(pts), np.inf)
            np.minimum.at(rmin, cells.reshape(-1), np.repeat(mm.cell_inradius, 3))
            ang = rs.rand(len(pts)) * 2 * np.pi
            step = 0.45 * rmin[:, None] * np.stack([np.cos(ang), np.sin(ang)], axis=1)
            step[bnd] = 0.0
            moved = moved + step
        a, b = MeshTri(moved, cells), MeshTri(moved, cells)
        if np.any(_signed_areas(moved, cells) * _signed_areas(pts, cells) <= 0):
            continue  # the random move inverted a cell: not a valid mesh
        done += 1
        fa, ra = a.flip_until_delaunay()
        fb, rb = b.flip_until_delaunay_survey()
        assert fa > 0 and fb > 0
        assert a.num_delaunay_violations() == 0 and b.num_delaunay_violations() == 0
        ca, cb = canonical_cells(a.cells("points")), canonical_cells(b.cells("points"))
        assert np.array_equal(ca, cb)
        assert np.array_equal(ca, canonical_cells(scipy.spatial.Delaunay(moved).simplices))
    assert done >= 4


def _signed_areas(pts, cells):
    p0, p1, p2 = pts[cells[:, 0]], pts[cells[:, 1]], pts[cells[:, 2]]
    return (p1[:, 0] - p0[:, 0]) * (p2[:, 1] - p0[:, 1]) - (p1[:, 1] - p0[:, 1]) * (p2[:, 0] - p0[:, 0])
