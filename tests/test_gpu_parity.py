"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle.

Bar (BASELINE.json north_star): cell arrays bit-exact on non-degenerate meshes, point
coordinates within 1e-10 relative per step (1e-8 after convergence), angle/quality
histograms equal.
"""
import json
import os

import numpy as np
import pytest
import scipy.spatial

import oracle
from oracle.meshtri import MeshTri as OMesh, canonical_cells

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
STEP_TOL = 1.0e-10  # relative to max |x|, per step
METHODS = ["lloyd", "cvt-block-diagonal", "cpt-fixed-point", "odt-fixed-point", "odt-dp-fp"]
SOLVE_METHODS = ["cpt-linear-solve", "cpt-quasi-newton"]  # iterative solve: 1e-9 per step


@pytest.fixture(scope="module")
def ob():
    import optimesh_b200

    return optimesh_b200


@pytest.fixture(scope="module")
def G():
    from optimesh_b200 import generators

    return generators


def rel_err(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def norms(p):
    return np.abs(p).sum(), np.sqrt((p * p).sum()), np.abs(p).max()


# ------------------------------------------------------------------ known answers
def test_simple1_lloyd_known_answer(ob, G):
    p, c = ob.optimize_points_cells(*G.SIMPLE1, "lloyd", 1.0e-2, 100)
    n1, n2, ninf = norms(p)
    assert abs(n1 - 4.986335452622451) <= 1e-12 * n1
    assert abs(n2 - 2.1181412069258942) <= 1e-12 * n2
    assert ninf == 1.0
    assert np.array_equal(c, G.SIMPLE1[1])


@pytest.mark.parametrize("method", ["cpt-fixed-point", "cpt-linear-solve", "odt-fixed-point"])
def test_simple1_cpt_odt_known_answer(ob, G, method):
    p, c = ob.optimize_points_cells(*G.SIMPLE1, method, 1.0e-2, 100)
    n1, n2, ninf = norms(p)
    assert abs(n1 - 5.0) < 1e-11 and abs(n2 - 2.1213203435596424) < 1e-11 and ninf == 1.0


def test_simple1_step_counts(ob, G):
    for method, expect in (("lloyd", [5, 10, 22]), ("CVT (block-diagonal)", [3, 5, 10])):
        for tol, n in zip((1e-2, 1e-3, 1e-5), expect):
            log = []
            ob.optimize_points_cells(*G.SIMPLE1, method, tol, 100, log=log)
            assert len(log) == n, (method, tol)


# ------------------------------------------------------------------ single step
def _meshes(G):
    out = {}
    out["disk40"] = G.disk(40, 3)
    out["disk120"] = G.disk(120, 0)
    out["square"] = G.square(30, 0.25, 1)
    sp, sc = G.tetra_sphere(6)
    rs = np.random.RandomState(5)
    sp = sp + rs.normal(scale=0.02, size=sp.shape)
    sp /= np.linalg.norm(sp, axis=1)[:, None]
    m = OMesh(sp, sc)
    m.flip_until_delaunay()
    out["sphere6"] = (sp, m.cells("points").copy())
    sp, sc = G.tetra_sphere(24)
    out["sphere24"] = (sp, sc)
    return out


@pytest.mark.parametrize("name", ["disk40", "disk120", "square", "sphere6", "sphere24"])
@pytest.mark.parametrize("method", METHODS + SOLVE_METHODS)
@pytest.mark.parametrize("renumber", [True, False])
def test_get_new_points_matches_oracle(ob, G, name, method, renumber):
    pts, cells = _meshes(G)[name]
    if method == "cpt-linear-solve" and name.startswith("sphere"):
        pytest.skip("closed surface: the graph Laplacian is singular")
    ref = oracle.get_new_points(OMesh(pts, cells), method)
    with ob.DeviceMesh(pts, cells, renumber=renumber) as dm:
        dm.set_method(method)
        got = dm.new_points()
    tol = 1e-9 if method in SOLVE_METHODS else STEP_TOL
    assert rel_err(got, ref) <= tol


@pytest.mark.parametrize("name", ["disk40", "disk120", "square"])
@pytest.mark.parametrize("method", ["odt-fixed-point", "odt-dp-fp"])
def test_odt_circumcenters_everywhere_option(ob, G, name, method, monkeypatch):
    """SURVEY.md A.8 as written: no barycenter substitution in cells with a boundary edge.
    The default (substitution on) is covered by every other ODT test."""
    import oracle.methods as om

    pts, cells = _meshes(G)[name]
    default = oracle.get_new_points(OMesh(pts, cells), method)
    monkeypatch.setattr(om, "ODT_BOUNDARY_BARYCENTERS", False)
    ref = oracle.get_new_points(OMesh(pts, cells), method)
    assert rel_err(default, ref) > 1e-6  # the two variants really differ on these meshes
    with ob.DeviceMesh(pts, cells) as dm:
        dm.set_method(method)
        dm.set_odt_boundary_barycenters(False)
        got = dm.new_points()
        dm.set_odt_boundary_barycenters(True)
        got_default = dm.new_points()
    assert rel_err(got, ref) <= STEP_TOL
    assert rel_err(got_default, default) <= STEP_TOL
    got2 = ob.get_new_points(ob.MeshTri(pts, cells), method, odt_boundary_barycenters=False)
    assert np.array_equal(got2, got)


def test_odt_boundary_cells_after_flips(ob, G):
    """Flips change which cells carry a boundary edge: the ring rows must follow."""
    pts, cells = G.disk(60, 4)
    for method in ("odt-fixed-point", "odt-dp-fp"):
        rp, rc = oracle.optimize_points_cells(pts, cells, method, 0.0, 12, omega=1.5)
        p, c = ob.optimize_points_cells(pts, cells, method, 0.0, 12, omega=1.5)
        assert np.array_equal(canonical_cells(c), canonical_cells(rc))
        assert rel_err(p, rp) <= 1e-8


@pytest.mark.parametrize("name", ["disk120", "square", "sphere24"])
@pytest.mark.parametrize("omega", [1.0, 1.5])
def test_quasi_newton_step_and_trajectory(ob, G, name, omega):
    """cpt-quasi-newton: one driver step (pin, omega, limiter) and a short run with flips."""
    pts, cells = _meshes(G)[name]
    om = OMesh(pts, cells)
    max_diff2, n_limited = oracle.driver.step(om, "cpt-quasi-newton", omega=omega)
    with ob.DeviceMesh(pts, cells) as dm:
        dm.set_method("cpt-quasi-newton", omega)
        st = dm.update_points(0.0)
        got = dm.points
    assert rel_err(got, om.points) <= 1e-9
    assert st["n_limited"] == n_limited
    assert abs(st["max_diff2"] - max_diff2) <= 1e-8 * max_diff2
    assert 0 < st["solver_iters"] <= 100  # condition number <= 5: a few dozen iterations
    rp, rc = oracle.optimize_points_cells(pts, cells, "cpt-quasi-newton", 0.0, 6, omega=omega)
    p, c = ob.optimize_points_cells(pts, cells, "cpt-quasi-newton", 0.0, 6, omega=omega)
    assert np.array_equal(canonical_cells(c), canonical_cells(rc))
    assert rel_err(p, rp) <= 1e-8


def test_quasi_newton_simple1_and_legacy_entry(ob, G):
    # a single free vertex: one quasi-Newton step is the CPT fixed-point step
    p, c = ob.cpt.quasi_newton(*G.SIMPLE1, 1.0e-2, 100)
    n1, n2, ninf = norms(p)
    assert abs(n1 - 5.0) < 1e-11 and abs(n2 - 2.1213203435596424) < 1e-11 and ninf == 1.0


def test_get_new_points_matches_golden_fixture(ob, G):
    z = np.load(os.path.join(GOLDEN, "single_steps.npz"))
    pts, cells = G.disk(40, 3)
    for m in METHODS:
        got = ob.get_new_points(ob.MeshTri(pts, cells), m)
        assert rel_err(got, z[f"disk40_{m}"]) <= STEP_TOL
    for m in SOLVE_METHODS:
        got = ob.get_new_points(ob.MeshTri(pts, cells), m)
        assert rel_err(got, z[f"disk40_{m}"]) <= 1e-9
    sp, sc = z["sphere6_points"], z["sphere6_cells"]
    for m in METHODS:
        got = ob.get_new_points(ob.MeshTri(sp, sc), m)
        assert rel_err(got, z[f"sphere6_{m}"]) <= STEP_TOL


@pytest.mark.parametrize("name", ["disk120", "square", "sphere24"])
@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("omega", [1.0, 2.0])
def test_update_points_matches_oracle_step(ob, G, name, method, omega):
    """Pin + omega + limiter, from identical input state."""
    pts, cells = _meshes(G)[name]
    om = OMesh(pts, cells)
    max_diff2, n_limited = oracle.driver.step(om, method, omega=omega)
    with ob.DeviceMesh(pts, cells) as dm:
        dm.set_method(method, omega)
        st = dm.update_points(0.0)
        got = dm.points
    assert rel_err(got, om.points) <= STEP_TOL
    assert st["n_limited"] == n_limited
    assert abs(st["max_diff2"] - max_diff2) <= 1e-9 * max_diff2
    bnd = om.is_boundary_point
    assert np.array_equal(got[bnd], pts[bnd])


# ------------------------------------------------------------------ flips
def _jittered(G, nb, seed, frac=0.45):
    pts, cells = G.disk(nb, seed)
    pts, cells = oracle.optimize_points_cells(pts, cells, "cpt-fixed-point", 0.0, 5)
    m = OMesh(pts, cells)
    rs = np.random.RandomState(seed)
    rmin = np.full(len(pts), np.inf)
    np.minimum.at(rmin, cells.reshape(-1), np.repeat(m.cell_inradius, 3))
    step = rs.uniform(-1, 1, size=pts.shape) * (frac * rmin / np.sqrt(2))[:, None]
    bnd = m.is_boundary_point
    p2 = pts.copy()
    p2[~bnd] += step[~bnd]
    return p2, cells


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("renumber", [True, False])
def test_flips_bit_exact_vs_oracle_and_qhull(ob, G, seed, renumber):
    pts, cells = _jittered(G, 80, seed)
    om = OMesh(pts, cells)
    nf, nr = om.flip_until_delaunay()
    with ob.DeviceMesh(pts, cells, renumber=renumber) as dm:
        gf, gr = dm.flip_until_delaunay()
        got = dm.cells()
        # the twin table must still be consistent: a second pass finds nothing,
        # and a smoothing step walks every star without error
        assert dm.flip_until_delaunay() == (0, 0)
        dm.set_method("lloyd")
        dm.update_points(0.0)
    assert (gf, gr) == (nf, nr) and nf > 0
    assert np.array_equal(got, om.cells("points"))  # row for row, slot for slot
    ref = scipy.spatial.Delaunay(pts).simplices
    assert np.array_equal(canonical_cells(got), canonical_cells(ref))


def test_flips_on_sphere_match_convex_hull(ob, G):
    pts, cells = G.tetra_sphere(12)
    rs = np.random.RandomState(0)
    p2 = pts + rs.normal(scale=0.02, size=pts.shape)
    p2 /= np.linalg.norm(p2, axis=1)[:, None]
    om = OMesh(p2, cells)
    om.flip_until_delaunay()
    with ob.DeviceMesh(p2, cells) as dm:
        dm.flip_until_delaunay()
        got = dm.cells()
    assert np.array_equal(got, om.cells("points"))
    hull = scipy.spatial.ConvexHull(p2).simplices
    assert np.array_equal(canonical_cells(got), canonical_cells(hull))


def test_mapped_grid_flips_to_delaunay(ob, G):
    pts, cells = G.disk_mapped_grid(120, 0.25, 0, shuffle=True)
    om = OMesh(pts, cells)
    nf, nr = om.flip_until_delaunay()
    with ob.DeviceMesh(pts, cells) as dm:
        gf, gr = dm.flip_until_delaunay()
        got = dm.cells()
    assert (gf, gr) == (nf, nr)
    assert np.array_equal(got, om.cells("points"))


# ------------------------------------------------------------------ trajectories
def test_config1_trajectory(ob, G):
    """BASELINE.json configs[0]: Lloyd omega=1, disk(120) (1,383 vertices), 50 steps,
    tol 1e-5 -- cells identical, points within 1e-8, histograms equal."""
    with open(os.path.join(GOLDEN, "config1_lloyd.json")) as f:
        gold = json.load(f)
    fin = np.load(os.path.join(GOLDEN, "config1_lloyd_final.npz"))
    pts, cells = G.disk(120, 0)
    log = []
    p, c = ob.optimize_points_cells(pts, cells, "lloyd", 1.0e-5, 50, log=log)
    assert len(log) == 50
    assert [l["n_flips"] for l in log] == gold["n_flips"]
    assert [l["n_flip_rounds"] for l in log] == gold["n_rounds"]
    assert [l["n_limited"] for l in log] == gold["n_limited"]
    assert np.array_equal(c, fin["cells"])
    assert rel_err(p, fin["points"]) <= 1e-8
    # same result without the per-step hooks (whole loop inside om_run)
    p2, c2 = ob.optimize_points_cells(pts, cells, "lloyd", 1.0e-5, 50)
    assert np.array_equal(p2, p) and np.array_equal(c2, c)
    with ob.DeviceMesh(p, c) as dm:
        ah, qh, s = dm.stats()
    assert np.abs(ah - np.array(gold["angle_hist"])).max() <= 1
    assert np.abs(qh - np.array(gold["q_hist"])).max() <= 1
    assert ah.sum() == 3 * len(c) and qh.sum() == len(c)
    for k, v in gold["summary"].items():
        assert abs(s[k] - v) <= 1e-7 * max(1.0, abs(v)), k


@pytest.mark.parametrize("method,omega", [("lloyd", 2.0), ("cvt-block-diagonal", 1.0),
                                          ("cpt-fixed-point", 1.0), ("odt-fixed-point", 1.0)])
def test_short_trajectories(ob, G, method, omega):
    pts, cells = G.disk(60, 7)
    olog, glog = [], []
    rp, rc = oracle.optimize_points_cells(pts, cells, method, 1e-6, 12, omega=omega, log=olog)
    p, c = ob.optimize_points_cells(pts, cells, method, 1e-6, 12, omega=omega, log=glog)
    assert [l["n_flips"] for l in glog] == [l["n_flips"] for l in olog]
    assert np.array_equal(c, rc)
    assert rel_err(p, rp) <= 1e-8


def test_sphere_odt_with_projection(ob, G):
    """Config 4 in miniature: ODT fixed-point on a sphere, projection every step."""
    pts, cells = G.tetra_sphere(16)
    olog, glog = [], []
    rp, rc = oracle.optimize_points_cells(pts, cells, "odt-fixed-point", 1e-4, 8,
                                          implicit_surface=ob.Sphere(), log=olog)
    p, c = ob.optimize_points_cells(pts, cells, "odt-fixed-point", 1e-4, 8,
                                    implicit_surface=ob.Sphere(), log=glog)
    assert np.abs(np.linalg.norm(p, axis=1) - 1.0).max() < 1e-10
    assert [l["n_flips"] for l in glog] == [l["n_flips"] for l in olog]
    assert np.array_equal(c, rc)
    assert rel_err(p, rp) <= 1e-8

    class UserSphere:  # the README's object: generic host path
        def f(self, x):
            return 1.0 - (x[0] ** 2 + x[1] ** 2 + x[2] ** 2)

        def grad(self, x):
            return -2 * x

    p2, c2 = ob.optimize_points_cells(pts, cells, "odt-fixed-point", 1e-4, 8,
                                      implicit_surface=UserSphere())
    assert np.array_equal(c2, rc) and rel_err(p2, rp) <= 1e-8


def test_hooks_on_device_match_host_hooks(ob, G):
    """a generic implicit surface (README.md:157-162) and a boundary_step
    callback (README.md:146-149) given torch-style callables run on device memory
    (`device_callables=True`): same trajectory as the host-hook path of the same library
    (numpy callables between the device phases), which in turn follows the oracle."""
    import torch

    # 1. generic surface: the README's object works on numpy arrays and on CUDA tensors alike
    class UserSphere:
        def f(self, x):
            return 1.0 - (x[0] ** 2 + x[1] ** 2 + x[2] ** 2)

        def grad(self, x):
            return -2 * x

    pts, cells = G.tetra_sphere(16)
    rp, rc = oracle.optimize_points_cells(pts, cells, "odt-fixed-point", 1e-4, 8,
                                          implicit_surface=ob.Sphere())
    hlog, dlog = [], []
    hp, hc = ob.optimize_points_cells(pts, cells, "odt-fixed-point", 1e-4, 8,
                                      implicit_surface=UserSphere(), log=hlog)
    dp, dc = ob.optimize_points_cells(pts, cells, "odt-fixed-point", 1e-4, 8,
                                      implicit_surface=UserSphere(), log=dlog,
                                      device_callables=True)
    assert np.array_equal(dc, hc) and np.array_equal(dc, rc)
    assert rel_err(dp, hp) <= 1e-12 and rel_err(dp, rp) <= 1e-8
    assert np.abs(np.linalg.norm(dp, axis=1) - 1.0).max() < 1e-10
    assert [l["n_flips"] for l in dlog] == [l["n_flips"] for l in hlog]
    assert any(l["surface_sweeps"] > 0 for l in dlog)

    # 2. boundary_step: Lloyd proposes the control-volume centroid for boundary vertices too;
    #    the callback puts it back onto the circle.  The device variant sees a CUDA tensor in
    #    the caller's vertex order.
    seen = {}

    def to_circle_np(x):
        return x / np.sqrt(np.einsum("ij,ij->j", x, x))

    def to_circle_torch(x):
        seen["cuda"] = x.is_cuda
        seen["shape"] = tuple(x.shape)
        return x / torch.sqrt((x * x).sum(0))

    to_circle_torch.on_device = True  # (the attribute opts in, like device_callables=True)
    pts, cells = G.disk(60, 3)
    nb = int(OMesh(pts, cells).is_boundary_point.sum())
    for method, omega in (("lloyd", 1.0), ("cvt-block-diagonal", 1.0), ("lloyd", 2.0)):
        hlog, dlog = [], []
        hp, hc = ob.optimize_points_cells(pts, cells, method, 0.0, 8, omega=omega,
                                          boundary_step=to_circle_np, log=hlog)
        dp, dc = ob.optimize_points_cells(pts, cells, method, 0.0, 8, omega=omega,
                                          boundary_step=to_circle_torch, log=dlog)
        assert seen == {"cuda": True, "shape": (2, nb)}
        assert np.array_equal(dc, hc)
        assert rel_err(dp, hp) <= 1e-11
        assert [l["n_flips"] for l in dlog] == [l["n_flips"] for l in hlog]
        assert [l["n_limited"] for l in dlog] == [l["n_limited"] for l in hlog]
    bnd = OMesh(pts, cells).is_boundary_point
    assert not np.allclose(dp[bnd], pts[bnd])  # the boundary vertices slid along the circle
    assert np.abs(np.linalg.norm(dp[bnd], axis=1) - 1.0).max() < 0.02
    with pytest.raises(ValueError):
        ob.optimize_points_cells(pts, cells, "lloyd", 0.0, 1, boundary_step=lambda x: x[:, :-1],
                                 device_callables=True)


def test_solver_failures_are_reported(ob, G):
    """An iterative solve that does not reach its tolerance is an error, never a silent OK;
    a closed surface (no fixed vertex: singular graph Laplacian) is refused."""
    from optimesh_b200._lib import OptimeshError

    pts, cells = G.square(60, 0.25, 0)
    with ob.DeviceMesh(pts, cells) as dm:
        with pytest.raises(OptimeshError, match="stopped after"):
            dm.solve_graph_laplacian(1e-14, 5)
    big, bc = G.disk(900, 5)  # above the multigrid threshold
    with ob.DeviceMesh(big, bc) as dm:
        with pytest.raises(OptimeshError, match="stopped after"):
            dm.solve_graph_laplacian(1e-14, 3)
    sp, sc = G.tetra_sphere(8)
    with ob.DeviceMesh(sp, sc) as dm:
        with pytest.raises(ValueError, match="needs a boundary"):
            dm.solve_graph_laplacian(1e-10, 100)
    with pytest.raises(ValueError, match="needs a boundary"):
        ob.optimize_points_cells(sp, sc, "cpt-linear-solve", 1e-6, 2)


def test_cpt_linear_solve_vs_spsolve(ob, G):
    """Config 3 in miniature: Laplacian smoothing on a jittered square, boundary pinned."""
    pts, cells = G.square(60, 0.25, 0)
    ref = oracle.get_new_points(OMesh(pts, cells), "cpt-linear-solve")
    with ob.DeviceMesh(pts, cells) as dm:
        its, res = dm.solve_graph_laplacian(1e-14, 20000)
        got = dm.points
    assert res <= 1e-13 and its > 0
    assert rel_err(got, ref) <= 1e-10
    bnd = OMesh(pts, cells).is_boundary_point
    assert np.array_equal(got[bnd], pts[bnd])


def test_cpt_linear_solve_multigrid_vs_spsolve(ob, G):
    """Above 20,000 vertices the solve is preconditioned by aggregation multigrid (pcg.cu):
    same answer as the sparse direct solve, a mesh-independent-ish iteration count, and the
    same bits from run to run (integer-valued Galerkin weights, fixed summation orders)."""
    pts, cells = G.disk(900, 5)  # random (Qhull)
    assert pts.shape[0] > 60000
    ref = oracle.get_new_points(OMesh(pts, cells), "cpt-linear-solve")
    runs = []
    for _ in range(2):
        with ob.DeviceMesh(pts, cells) as dm:
            its, res = dm.solve_graph_laplacian(1e-12, 2000)
            runs.append((its, dm.points))
    its, got = runs[0]
    assert res <= 1e-12 and 0 < its <= 120, its
    assert rel_err(got, ref) <= 1e-9
    bnd = OMesh(pts, cells).is_boundary_point
    assert np.array_equal(got[bnd], pts[bnd])
    assert runs[1][0] == its and runs[1][1].tobytes() == got.tobytes()
    # the loop with the solve as its update (flips change the matrix between steps)
    rp, rc = oracle.optimize_points_cells(pts, cells, "cpt-linear-solve", 0.0, 2)
    gp, gc = ob.optimize_points_cells(pts, cells, "cpt-linear-solve", 0.0, 2)
    assert np.array_equal(gc, rc) and rel_err(gp, rp) <= 1e-8


# ------------------------------------------------------------------ properties, edge cases
def test_bitwise_deterministic(ob, G):
    pts, cells = G.disk(100, 4)
    a = ob.optimize_points_cells(pts, cells, "lloyd", 0.0, 10, omega=2.0)
    b = ob.optimize_points_cells(pts, cells, "lloyd", 0.0, 10, omega=2.0)
    assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes()


def test_inputs_untouched_and_dtypes(ob, G):
    pts, cells = G.disk(40, 1)
    for dt in (np.int32, np.int64, np.uint32):
        p0, c0 = pts.copy(), cells.astype(dt)
        p, c = ob.optimize_points_cells(p0, c0, "cpt-fixed-point", 1e-3, 5)
        assert np.array_equal(p0, pts) and np.array_equal(c0, cells.astype(dt))
        assert c.dtype == dt and p.dtype == np.float64 and p.shape == pts.shape


def test_vertex_numbering_is_preserved(ob, G):
    pts, cells = G.disk(50, 2)
    ps, cs = G.shuffle_vertices(pts, cells, 3)
    p1, c1 = ob.optimize_points_cells(pts, cells, "lloyd", 0.0, 6)
    p2, c2 = ob.optimize_points_cells(ps, cs, "lloyd", 0.0, 6)
    rs = np.random.RandomState(3 + 12345)
    perm = rs.permutation(len(pts))
    assert rel_err(p2, p1[perm]) < 1e-9
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    assert np.array_equal(canonical_cells(c2), canonical_cells(inv[c1]))


def test_mixed_orientation_and_orphans(ob, G):
    pts, cells = G.square(12, 0.2, 3)
    cells = cells.copy()
    cells[::3] = cells[::3][:, [0, 2, 1]]  # flip the orientation of every third cell
    pts = np.concatenate([pts, [[5.0, 5.0]]])  # an orphan vertex
    for m in METHODS:
        ref = oracle.get_new_points(OMesh(pts, cells), m)
        got = ob.get_new_points(ob.MeshTri(pts, cells), m)
        assert rel_err(got, ref) <= STEP_TOL
        assert np.array_equal(got[-1], [5.0, 5.0])


def test_single_triangle_and_masked_cells(ob):
    pts = np.array([[0.0, 0.0], [1.0, 0.0], [0.3, 0.8]])
    cells = np.array([[0, 1, 2]])
    p, c = ob.optimize_points_cells(pts, cells, "lloyd", 1e-3, 3)
    assert np.array_equal(p, pts) and np.array_equal(c, cells)
    # interior vertex whose cells are all obtuse (> 135 deg) -> fully masked -> stays
    pts = np.array([[0.0, 0.0], [10.0, 0.0], [5.0, 0.3], [5.0, -0.3], [5.0, 0.0]])
    cells = np.array([[0, 4, 2], [4, 1, 2], [0, 3, 4], [3, 1, 4]])
    for m in ("lloyd", "cvt-block-diagonal"):
        ref = oracle.get_new_points(OMesh(pts, cells), m)
        got = ob.get_new_points(ob.MeshTri(pts, cells), m)
        assert np.allclose(got, ref, rtol=0, atol=1e-12)


def test_errors(ob, G):
    from optimesh_b200._lib import DegenerateCellsError, MeshTopologyError

    pts = np.array([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0], [1.0, 1.0]])
    with pytest.raises(DegenerateCellsError):
        ob.optimize_points_cells(pts, np.array([[0, 1, 2], [0, 1, 3]]), "lloyd", 1e-3, 2)
    with pytest.raises(MeshTopologyError):
        ob.optimize_points_cells(pts, np.array([[0, 1, 9]]), "lloyd", 1e-3, 2)
    with pytest.raises(KeyError):
        ob.optimize_points_cells(*G.SIMPLE1, "nope", 1e-3, 2)
    with pytest.raises(NotImplementedError):
        ob.optimize_points_cells(*G.SIMPLE1, "CVT (full)", 1e-3, 2)
    tri3 = np.array([[0, 1, 3], [0, 1, 2], [0, 1, 3]])
    with pytest.raises(MeshTopologyError):
        ob.DeviceMesh(np.array([[0.0, 0], [1, 0], [0, 1], [1, 1.0]]),
                      np.array([[0, 1, 2], [0, 1, 3], [1, 0, 3]]))
    del tri3


def test_stats_match_oracle(ob, G):
    for pts, cells in (G.disk(90, 5), G.tetra_sphere(10)):
        ah, qh, s = oracle.stats(OMesh(pts, cells))
        with ob.DeviceMesh(pts, cells) as dm:
            gah, gqh, gs = dm.stats()
        assert np.abs(gah - ah).max() <= 1 and np.abs(gqh - qh).max() <= 1
        assert gah.sum() == ah.sum() and gqh.sum() == qh.sum()
        for k, v in s.items():
            assert abs(gs[k] - v) <= 1e-9 * max(1.0, abs(v)), k


# ------------------------------------------------------------------ larger sizes
def test_single_step_100k(ob, G):
    pts, cells = G.disk_mapped_grid(330, 0.25, 1, shuffle=True)  # 108,900 vertices
    om = OMesh(pts, cells)
    om.flip_until_delaunay()
    c0 = om.cells("points").copy()
    for method in ("lloyd", "cvt-block-diagonal"):
        om2 = OMesh(pts, c0)
        oracle.driver.step(om2, method)
        with ob.DeviceMesh(pts, c0) as dm:
            dm.set_method(method)
            dm.update_points(0.0)
            got = dm.points
        assert rel_err(got, om2.points) <= STEP_TOL


def test_full_size_properties(ob, G):
    """~1M vertices: properties that need no oracle run at this size."""
    pts, cells = G.disk_mapped_grid(1000, 0.25, 0)
    with ob.DeviceMesh(pts, cells) as dm:
        nf, nr = dm.flip_until_delaunay()
        assert nf > 0 and dm.flip_until_delaunay() == (0, 0)  # idempotent
        c = dm.cells()
        # flips conserve the cell count, every vertex's presence and the total area
        assert c.shape == cells.shape and np.array_equal(np.unique(c), np.arange(len(pts)))
        dm.set_method("lloyd", 1.0)
        bnd = dm.is_boundary_point
        assert bnd.sum() == 4 * 999
        for _ in range(3):
            st = dm.step(0.0)
        p = dm.points
        assert np.array_equal(p[bnd], pts[bnd])  # boundary pinned
        ah, qh, s = dm.stats()
        assert ah.sum() == 3 * len(cells) and qh.sum() == len(cells)
        def total_area(q, cc):
            u, w = q[cc[:, 1]] - q[cc[:, 0]], q[cc[:, 2]] - q[cc[:, 0]]
            return 0.5 * np.abs(u[:, 0] * w[:, 1] - u[:, 1] * w[:, 0]).sum()

        area, area0 = total_area(p, dm.cells()), total_area(pts, cells)
        assert abs(area - area0) < 1e-9 * area0


# ------------------------------------------------------------------ sharded update
@pytest.mark.parametrize("method,omega", [("lloyd", 2.0), ("cvt-block-diagonal", 1.0)])
def test_sharded_simulation(ob, G, method, omega):
    """Partition simulator: P handles hold the same mesh in one process, each updates its own
    internal vertex range, ranges are copied across (what the NCCL all-gather does between
    GPUs).  Must be bit-identical to the single-handle run."""
    import torch

    from optimesh_b200.dist import device_points_tensor, owned_range

    pts, cells = G.disk(90, 11)
    P, steps = 3, 6
    ref_p, ref_c = ob.optimize_points_cells(pts, cells, method, 0.0, steps, omega=omega)
    hs = [ob.DeviceMesh(pts, cells) for _ in range(P)]
    try:
        for r, h in enumerate(hs):
            h.set_method(method, omega)
            h.flip_until_delaunay()
            h.set_owned_range(*owned_range(h.n, r, P))
        for _ in range(steps):
            for h in hs:
                h.update_points(0.0)
            xs = [device_points_tensor(h) for h in hs]
            torch.cuda.synchronize()
            for r in range(P):
                lo, hi = owned_range(hs[r].n, r, P)
                for s in range(P):
                    if s != r:
                        xs[s][lo:hi] = xs[r][lo:hi]
            torch.cuda.synchronize()
            # first flip round sharded by cell range, records applied by every handle
            recs = [h.flip_check_range(*owned_range(h.c, r, P)) for r, h in enumerate(hs)]
            for h in hs:
                for ptr, cnt in recs:
                    if cnt:
                        h.flip_add_records(ptr, cnt)
            flips = [h.flip_finish() for h in hs]
            assert all(f == flips[0] for f in flips)
        for h in hs:
            assert np.array_equal(h.points, ref_p) and np.array_equal(h.cells(), ref_c)
    finally:
        for h in hs:
            h.close()


# ------------------------------------------------------------------ BASELINE configs 3, 4 at full size
def test_config3_full_size_square_cpt(ob, G):
    """configs[2]: CPT linear-solve and CPT fixed-point on the 5M-vertex jittered square,
    boundary pinned.  Size-independent checks: the solve is harmonic on interior rows
    (independent scipy matvec), boundary untouched, fixed-point steps keep the boundary."""
    import scipy.sparse

    n = 2236
    pts, cells = G.square(n, 0.25, 0)
    assert pts.shape[0] == 4999696 and cells.shape[0] == 9990450
    with ob.DeviceMesh(pts, cells.astype(np.int32)) as dm:
        bnd = dm.is_boundary_point
        assert bnd.sum() == 4 * (n - 1)
        dm.flip_until_delaunay()
        c = dm.cells()
        its, res = dm.solve_graph_laplacian(1e-10, 100000)
        x = dm.points
        assert res <= 1e-10 and its > 0
        assert np.array_equal(x[bnd], pts[bnd])
        # residual of the Dirichlet graph Laplacian, computed independently on the host
        i = np.concatenate([c[:, 0], c[:, 1], c[:, 2], c[:, 1], c[:, 2], c[:, 0]])
        j = np.concatenate([c[:, 1], c[:, 2], c[:, 0], c[:, 0], c[:, 1], c[:, 2]])
        A = scipy.sparse.coo_matrix((np.ones(i.size), (i, j)), shape=(len(pts), len(pts))).tocsr()
        deg = np.asarray(A.sum(axis=1)).ravel()
        r = (deg[:, None] * x - A @ x)[~bnd]
        assert np.abs(r).max() <= 1e-9 * deg.max()
        dm.points = pts
        dm.set_method("cpt-fixed-point")
        for _ in range(3):
            st = dm.step(0.0)
        p = dm.points
        assert np.array_equal(p[bnd], pts[bnd]) and np.isfinite(p).all()
        assert dm.flip_until_delaunay() == (0, 0)


def test_config4_full_size_sphere_odt(ob, G):
    """configs[3]: ODT fixed-point on the 2M-vertex tetra-sphere with projection each step."""
    pts, cells = G.tetra_sphere(1000)
    assert pts.shape[0] == 2000002 and cells.shape[0] == 4000000
    with ob.DeviceMesh(pts, cells.astype(np.int32)) as dm:
        assert not dm.is_boundary_point.any()
        dm.set_method("odt-fixed-point")
        dm.set_sphere()
        dm.flip_until_delaunay()
        q0 = dm.stats()[2]["q_avg"]
        for _ in range(5):
            st = dm.step(0.0)
        p, c = dm.points, dm.cells()
        assert np.abs(np.linalg.norm(p, axis=1) - 1.0).max() <= 1e-10
        assert c.shape[0] == 2 * p.shape[0] - 4  # closed genus-0 surface: F = 2V - 4
        assert np.array_equal(np.unique(c), np.arange(len(p)))
        assert dm.stats()[2]["q_avg"] > q0  # smoothing improves the average quality
        assert dm.flip_until_delaunay() == (0, 0)


# ------------------------------------------------------------------ more edge cases
def test_empty_and_tiny_inputs(ob):
    p, c = ob.optimize_points_cells(np.zeros((0, 2)), np.zeros((0, 3), dtype=np.int64), "lloyd",
                                    1e-3, 3)
    assert p.shape == (0, 2) and c.shape == (0, 3)
    # points but no cells: nothing moves
    pts = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
    p, c = ob.optimize_points_cells(pts, np.zeros((0, 3), dtype=np.int64), "cpt-fixed-point",
                                    1e-3, 3)
    assert np.array_equal(p, pts) and c.shape == (0, 3)
    # two cells, everything on the boundary
    pts = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])
    cells = np.array([[0, 1, 2], [0, 2, 3]])
    for m in METHODS + ["cpt-linear-solve"]:
        p, c = ob.optimize_points_cells(pts, cells, m, 1e-3, 3)
        assert np.array_equal(p, pts)
        assert np.array_equal(canonical_cells(c), canonical_cells(cells))


def test_high_valence_vertices_use_the_walk_path(ob):
    """A fan of 12 cells around one vertex: no ring row fits (OM_RING_W = 8), so the
    list-driven walk kernel must produce the oracle's numbers."""
    k = 12
    t = np.arange(k) * 2 * np.pi / k
    rs = np.random.RandomState(0)
    ring = np.stack([np.cos(t), np.sin(t)], axis=1) * (1 + 0.2 * rs.rand(k))[:, None]
    outer = 2.5 * np.stack([np.cos(t + 0.1), np.sin(t + 0.1)], axis=1)
    pts = np.concatenate([[[0.05, -0.02]], ring, outer])
    cells = [[0, 1 + i, 1 + (i + 1) % k] for i in range(k)]
    cells += [[1 + i, 1 + k + i, 1 + (i + 1) % k] for i in range(k)]
    cells += [[1 + (i + 1) % k, 1 + k + i, 1 + k + (i + 1) % k] for i in range(k)]
    cells = np.array(cells)
    for m in METHODS:
        om = OMesh(pts, cells)
        oracle.driver.step(om, m)
        with ob.DeviceMesh(pts, cells) as dm:
            dm.set_method(m)
            st = dm.update_points(0.0)
            got = dm.points
        assert rel_err(got, om.points) <= STEP_TOL, m


def test_random_meshes_property(ob, G):
    """Property check over random valid meshes (seeded): one full step incl. flips equals
    the oracle, for every method, boundary-heavy and tiny meshes included."""
    rs = np.random.RandomState(123)
    for trial in range(12):
        nb = int(rs.randint(8, 60))
        pts, cells = G.disk(nb, int(rs.randint(0, 1000)))
        method = METHODS[trial % len(METHODS)]
        omega = float(rs.choice([1.0, 1.5, 2.0]))
        om = OMesh(pts, cells)
        om.flip_until_delaunay()
        oracle.driver.step(om, method, omega=omega)
        of = om.flip_until_delaunay()
        with ob.DeviceMesh(pts, cells) as dm:
            dm.set_method(method, omega)
            dm.flip_until_delaunay()
            st = dm.step(0.0)
            p, c = dm.points, dm.cells()
        assert (st["n_flips"], st["n_flip_rounds"]) == of, (trial, method)
        assert np.array_equal(c, om.cells("points")), (trial, method)
        assert rel_err(p, om.points) <= STEP_TOL, (trial, method)


def test_cli_end_to_end(ob, G, tmp_path):
    from optimesh_b200 import cli, io

    pts, cells = G.disk(40, 6)
    src, dst = str(tmp_path / "in.vtk"), str(tmp_path / "out.vtk")
    io.write(src, pts, cells)
    assert cli.main([src, dst, "-m", "lloyd", "--omega", "2.0", "-n", "5", "-t", "0", "-q"]) == 0
    p, c = io.read(dst)
    rp, rc = oracle.optimize_points_cells(pts, cells, "lloyd", 0.0, 5, omega=2.0)
    assert np.array_equal(c, rc) and rel_err(p, rp) <= 1e-8
    # verbose path prints the two histograms; step dumps are written
    fmt = str(tmp_path / "s{:02d}.npz")
    assert cli.main([src, dst, "-m", "cpt-fixed-point", "-n", "2", "-t", "0", "-f", fmt]) == 0
    assert os.path.exists(fmt.format(1)) and os.path.exists(fmt.format(2))


def test_partitioned_simulation(ob, G):
    """The partitioned-coordinates path of dist.py (band exchange, validated reads, deferred
    commit, round-wise flip pass with gathered record slots) emulated in one process with P
    handles; what NCCL does between GPUs is done here with device copies.  Must be
    bit-identical to the single-handle run."""
    import torch

    from optimesh_b200.dist import _DevPtr, owned_range

    pts, cells = G.disk_mapped_grid(70, 0.25, 3, shuffle=True)
    P, steps, cap = 3, 6, 1 << 14
    method, omega = "cvt-block-diagonal", 1.0
    log = []
    ref_p, ref_c = ob.optimize_points_cells(pts, cells, method, 0.0, steps, omega=omega, log=log)
    hs = [ob.DeviceMesh(pts, cells) for _ in range(P)]

    def flip_pass():
        for h in hs:
            h.flip_pass_begin()
        first = True
        while True:
            slots = []
            for r, h in enumerate(hs):
                h.flip_round_check_nofetch(first, *cranges[r])
                slot = torch.zeros(cap + 1, 2, dtype=torch.float64, device="cuda")
                h.flip_round_pack(cap, slot.data_ptr())
                slots.append(slot)
            torch.cuda.synchronize()
            gathered = torch.cat(slots).contiguous()
            res = [h.flip_round_apply_gathered(gathered.data_ptr(), P, cap) for h in hs]
            assert all(x == res[0] for x in res) and res[0][2] == 0  # same everywhere, no abort
            first = False
            if res[0][0] == 0:
                break
        ends = [h.flip_pass_end() for h in hs]
        assert all(e == ends[0] for e in ends)
        return ends[0]

    try:
        ranges = [owned_range(hs[0].n, r, P) for r in range(P)]
        cranges = [hs[r].cell_range_of_vertices(ranges[r][0], ranges[r][1] if r < P - 1
                                                else hs[r].n + 1) for r in range(P)]
        assert cranges[0][0] == 0 and cranges[-1][1] == hs[0].c
        for r, h in enumerate(hs):
            h.set_method(method, omega)
            h.set_owned_range(*ranges[r])
            h.coords_all_valid()
        flip_pass()
        for h in hs:
            h.set_deferred_commit(True)
        for k in range(steps):
            sts = [h.update_points(0.0) for h in hs]
            assert all(st["stale"] == 0 for st in sts)
            assert sum(st["n_limited"] for st in sts) == log[k]["n_limited"]
            for h in hs:
                h.commit_points()
                h.coords_invalidate()
            bands = [h.band_build(3) for h in hs]
            stride = hs[0].points_device()[2]
            for s_, (ptr, n) in enumerate(bands):
                assert 0 < n < hs[0].n // 2  # a band, not the whole range
                buf = torch.zeros(n, stride, dtype=torch.float64, device="cuda")
                hs[s_].band_pack(ptr, n, buf.data_ptr())
                torch.cuda.synchronize()
                for r_, h in enumerate(hs):
                    if r_ != s_:
                        h.band_unpack(ptr, n, buf.data_ptr())
                torch.cuda.synchronize()
            nf, nr = flip_pass()
            assert (nf, nr) == (log[k]["n_flips"], log[k]["n_flip_rounds"])
        # own ranges are current everywhere they are owned: assemble and compare
        xs = []
        for r, h in enumerate(hs):
            ptr, n_alloc, st = h.points_device()
            xs.append(torch.as_tensor(_DevPtr(ptr, (n_alloc, st), "<f8"), device="cuda"))
        for r in range(P):
            lo, hi = ranges[r]
            for s_ in range(P):
                if s_ != r:
                    xs[s_][lo:hi] = xs[r][lo:hi]
        torch.cuda.synchronize()
        for h in hs:
            h.set_deferred_commit(False)
            h.set_owned_range(0, -1)
            assert np.array_equal(h.points, ref_p) and np.array_equal(h.cells(), ref_c)
    finally:
        for h in hs:
            h.close()


# ------------------------------------------------------------------ pipelined loop (loop.cu)
@pytest.mark.parametrize("method,omega", [("lloyd", 2.0), ("cvt-block-diagonal", 1.0),
                                          ("cpt-fixed-point", 1.0), ("odt-fixed-point", 1.0),
                                          ("odt-dp-fp", 1.5)])
def test_pipelined_loop_equals_stepwise(ob, G, method, omega):
    """om_run (update with the fused Delaunay check, flips, recomputation of the touched
    vertices, all in one CUDA graph) against the reference order step by step (om_step):
    same bytes, same statistics of the last step."""
    meshes = [G.disk(80, 2), G.disk_mapped_grid(40, 0.28, 1, shuffle=True), G.disk(12, 5)]
    for pts, cells in meshes:
        for nsteps in (1, 2, 7):
            with ob.DeviceMesh(pts, cells) as a:
                a.set_method(method, omega)
                steps, last = a.run(0.0, nsteps)
                pa, ca = a.points, a.cells()
            with ob.DeviceMesh(pts, cells) as b:
                b.set_method(method, omega)
                b.flip_until_delaunay()
                for _ in range(nsteps):
                    st = b.step(0.0)
                pb, cb = b.points, b.cells()
            assert steps == nsteps
            assert np.array_equal(ca, cb), (method, nsteps)
            assert np.array_equal(pa, pb), (method, nsteps)
            for key in ("max_diff2", "n_limited", "n_flips", "n_flip_rounds"):
                assert last[key] == st[key], (method, nsteps, key)


def test_pipelined_loop_stops_at_tolerance(ob, G):
    pts, cells = G.disk(60, 7)
    for method, tol in (("lloyd", 2e-3), ("cvt-block-diagonal", 1e-3)):
        log = []
        p, c = ob.optimize_points_cells(pts, cells, method, tol, 200, log=log)  # step by step
        with ob.DeviceMesh(pts, cells) as dm:
            dm.set_method(method, 1.0)
            steps, last = dm.run(tol, 200)
            assert steps == len(log) and 1 < steps < 200
            assert np.array_equal(dm.points, p) and np.array_equal(dm.cells(), c)
            # a second run on the same handle continues from there (graph cache, buffer parity)
            steps2, _ = dm.run(0.0, 3)
            p3 = dm.points
        p4, _ = ob.optimize_points_cells(p, c, method, 0.0, 3)
        assert steps2 == 3 and rel_err(p3, p4) <= 1e-12


def test_pipelined_loop_without_graph(ob, G, tmp_path):
    """OM_NO_GRAPH=1 runs the same kernels from the stream: same bytes."""
    import subprocess
    import sys

    pts, cells = G.disk(70, 3)
    with ob.DeviceMesh(pts, cells) as dm:
        dm.set_method("cvt-block-diagonal", 1.0)
        dm.run(0.0, 6)
        p, c = dm.points, dm.cells()
    out = tmp_path / "nograph.npz"
    code = ("import numpy as np, optimesh_b200 as ob\n"
            "from optimesh_b200 import generators as G\n"
            "pts, cells = G.disk(70, 3)\n"
            "with ob.DeviceMesh(pts, cells) as dm:\n"
            "    dm.set_method('cvt-block-diagonal', 1.0)\n"
            "    dm.run(0.0, 6)\n"
            f"    np.savez(r'{out}', p=dm.points, c=dm.cells())\n")
    env = dict(os.environ, OM_NO_GRAPH="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, "-c", code], check=True, env=env, cwd=root)
    got = np.load(out)
    assert np.array_equal(got["p"], p) and np.array_equal(got["c"], c)


def test_random_walk_gives_a_random_delaunay_mesh(ob, G):
    """om_random_walk (synthetic 'random disk mesh' workloads): the result is a valid Delaunay
    triangulation of the same domain, boundary untouched, reproducible, independent of the
    internal numbering, with the vertex-degree spread of a random mesh."""
    pts, cells = G.disk_mapped_grid(50, 0.25, 0)
    res = []
    for renumber in (True, False, True):
        with ob.DeviceMesh(pts, cells, renumber=renumber) as dm:
            nf = dm.random_walk(80, seed=3)
            res.append((dm.points, dm.cells(), nf))
            bnd = dm.is_boundary_point
    p, c, nf = res[0]
    assert nf > 0
    assert np.array_equal(res[2][0], p) and np.array_equal(res[2][1], c)
    assert rel_err(res[1][0], p) <= 1e-13  # same random numbers whatever the numbering
    assert np.array_equal(p[bnd], pts[bnd])
    qh = scipy.spatial.Delaunay(p).simplices
    assert np.array_equal(canonical_cells(c), canonical_cells(qh))
    val = np.bincount(c.reshape(-1), minlength=len(p))[~bnd]
    assert (val > 8).mean() > 0.01 and val.max() >= 9


# ------------------------------------------------------------------ oracle parity at BASELINE sizes
def _avail_gb():
    try:
        import psutil

        return psutil.virtual_memory().available / 2**30
    except Exception:
        return 0.0


@pytest.mark.parametrize("method", METHODS)
def test_single_step_1m_random_mesh(ob, G, method):
    """SURVEY.md section 4 tier 2 at 1M vertices: one update (pin, omega, exact limiter) from
    an identical state on a RANDOM disk mesh (device generator, valence up to 12) against
    the oracle, then the flip pass: same flips, same rounds, same cell rows."""
    dm0 = G.disk_gpu(1000, rounds=40, seed=method.__hash__() % 7)
    pts, cells = dm0.points, dm0.cells(np.int64)
    dm0.close()
    om = OMesh(pts, cells)
    md2, nlim = oracle.driver.step(om, method, omega=1.0)
    with ob.DeviceMesh(pts, cells) as dm:
        dm.set_method(method, 1.0)
        st = dm.update_points(0.0)
        got = dm.points
        assert rel_err(got, om.points) <= STEP_TOL
        assert st["n_limited"] == nlim
        assert abs(st["max_diff2"] - md2) <= 1e-9 * md2
        of = om.flip_until_delaunay()
        gf = dm.flip_until_delaunay()
        assert gf == of and of[0] > 1000
        assert np.array_equal(dm.cells(np.int64), om.cells("points"))


def test_config2_single_step_10m(ob, G):
    """configs[1] at full size: one cvt-block-diagonal update of the 9.95M-vertex random disk
    mesh the bench runs on, from an identical state, against the oracle (the numpy step needs
    about 45 GB and a minute or two)."""
    if _avail_gb() < 100:
        pytest.skip("needs ~100 GB of host memory for the numpy oracle at 10M vertices")
    dm0 = G.disk_gpu(3154, rounds=120, seed=0)
    pts, cells = dm0.points, dm0.cells(np.int64)
    dm0.set_method("cvt-block-diagonal", 1.0)
    st = dm0.update_points(0.0)
    got = dm0.points
    dm0.close()
    assert pts.shape[0] == 9947716
    om = OMesh(pts, cells)
    md2, nlim = oracle.driver.step(om, "cvt-block-diagonal", omega=1.0)
    assert rel_err(got, om.points) <= STEP_TOL
    assert st["n_limited"] == nlim
    assert abs(st["max_diff2"] - md2) <= 1e-9 * md2


def test_config4_single_step_2m_sphere(ob, G):
    """configs[3] at full size: one odt-fixed-point update + sphere projection of the
    2M-vertex tetra-sphere (vertices perturbed on the sphere so that the update is not
    trivial) against the oracle."""
    pts, cells = G.tetra_sphere(1000)
    rs = np.random.RandomState(0)
    pts = pts + rs.normal(scale=1.0e-4, size=pts.shape)
    pts /= np.linalg.norm(pts, axis=1)[:, None]
    om = OMesh(pts, cells)
    c0 = cells
    sphere = ob.Sphere()
    md2, nlim = oracle.driver.step(om, "odt-fixed-point", implicit_surface=sphere)
    with ob.DeviceMesh(pts, c0) as dm:
        dm.set_method("odt-fixed-point")
        dm.set_sphere()
        st = dm.update_points(0.0)
        dm.project()
        got = dm.points
    assert rel_err(got, om.points) <= STEP_TOL
    assert st["n_limited"] == nlim


def test_cli_subdomains_on_gpu(ob, G, tmp_path):
    """`optimesh in out -s NAME` (README.md:17 "preserves submeshes") through the device:
    every subdomain is optimized on its own, interface vertices stay put, cell data survive."""
    from optimesh_b200 import cli, io

    pts, cells = G.disk(40, 1)
    bary = pts[cells].mean(axis=1)
    tag = (bary[:, 0] > 0.05).astype(np.int64) + 2 * (bary[:, 1] > 0.1).astype(np.int64)
    src, dst = tmp_path / "in.vtk", tmp_path / "out.vtk"
    io.write(str(src), pts, cells, cell_data={"region": tag})
    assert cli.main([str(src), str(dst), "-m", "cvt-block-diagonal", "-n", "8", "-q", "-s", "region"]) in (0, None)
    p2, c2, cd = io.read(str(dst), with_cell_data=True)
    assert p2.shape == pts.shape and c2.shape == cells.shape
    # the submeshes are preserved: same multiset of tags, and every interface vertex (one
    # that touches cells of two regions) has not moved
    assert np.array_equal(np.sort(cd["region"]), np.sort(tag))
    touch = [set() for _ in range(len(pts))]
    for row, t in zip(cells, tag):
        for v in row:
            touch[v].add(int(t))
    interface = np.array([len(t) > 1 for t in touch])
    assert interface.sum() > 10
    assert np.array_equal(p2[interface], pts[interface])
    moved = np.abs(p2 - pts).max(axis=1) > 0
    assert moved.sum() > 0.5 * (~interface).sum()


def test_shared_address_space_loop_world_of_one(ob, G):
    """csrc/shared.cu with a single rank: the mesh arrays are CUDA-VMM chunks mapped into one
    virtual range, the loop is the graph with the device-side meetings (trivial at world 1).
    Must equal the ordinary pipelined loop bit for bit.  The N-GPU check of the same path is
    tests/run_shared_gpu.py (bit-identical to one GPU on 2 and 4 B200s)."""
    import torch
    import torch.distributed as dist

    from optimesh_b200.dist import SharedMesh

    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", str(29600 + os.getpid() % 300))
        dist.init_process_group("gloo", rank=0, world_size=1)
        created = True
    try:
        for method, omega in (("cvt-block-diagonal", 1.0), ("lloyd", 2.0)):
            pts, cells = G.disk(90, 4)
            with ob.DeviceMesh(pts, cells) as ref:
                ref.set_method(method, omega)
                k_ref, last_ref = ref.run(0.0, 6)
                p_ref, c_ref = ref.points, ref.cells()
            full = ob.DeviceMesh(pts, cells)
            full.set_method(method, omega)
            sm = SharedMesh.from_complete(full)
            full.close()
            try:
                k, last = sm.run(0.0, 6)
                torch.cuda.synchronize()
                info = sm.info()
                assert k == k_ref
                assert np.array_equal(sm.points, p_ref) and np.array_equal(sm.cells(), c_ref)
                for key in ("max_diff2", "n_limited", "n_flips", "n_flip_rounds"):
                    assert last[key] == last_ref[key], key
                assert (info["vertex_lo"], info["vertex_hi"]) == (0, len(pts))
                assert sm.time_update(2) > 0.0
                k2, _ = sm.run(1.0e-3, 500)  # stops on the tolerance, like the plain loop
                assert 1 <= k2 < 500
            finally:
                sm.close()
    finally:
        if created:
            dist.destroy_process_group()
