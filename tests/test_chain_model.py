"""The spoke regrouping used by the step kernel (tests/chain_model.py) against the oracle,
on the CPU: same new points, same max |diff|^2, same limited count, for every fixed-point
method; the lazy limiter bound never hides a limited vertex; the Delaunay pre-flag of the
fused check never misses a non-Delaunay edge."""
import numpy as np
import pytest

import oracle
from oracle.meshtri import MeshTri as OMesh
from optimesh_b200 import generators as G

from chain_model import step_model

METHODS = ["lloyd", "cvt-block-diagonal", "cpt-fixed-point", "odt-fixed-point", "odt-dp-fp"]


def _meshes():
    out = {"disk40": G.disk(40, 3), "square": G.square(14, 0.28, 1)}
    sp, sc = G.tetra_sphere(5)
    rs = np.random.RandomState(5)
    sp = sp + rs.normal(scale=0.02, size=sp.shape)
    out["sphere"] = (sp, sc)
    return out


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("name", ["disk40", "square", "sphere"])
def test_chain_model_matches_oracle(name, method):
    pts, cells = _meshes()[name]
    for omega in (1.0, 2.0):
        m = OMesh(pts, cells)
        bnd = m.is_boundary_point
        bc = m.is_boundary_cell if method.startswith("odt") else None
        new, md2, nlim, lazy_ok, _ = step_model(pts, cells, method, omega, bnd, bc)
        ref_md2, ref_nlim = oracle.driver.step(m, method, omega=omega)
        scale = np.abs(m.points).max()
        assert np.abs(new - m.points).max() <= 1e-12 * scale
        assert abs(md2 - ref_md2) <= 1e-10 * max(ref_md2, 1e-300)
        assert nlim == ref_nlim
        assert lazy_ok


def test_preflag_covers_every_non_delaunay_edge():
    pts, cells = G.disk(40, 3)
    rs = np.random.RandomState(1)
    m = OMesh(pts, cells)
    bnd = m.is_boundary_point
    moved = pts.copy()
    moved[~bnd] += rs.normal(scale=0.02, size=moved[~bnd].shape)
    m = OMesh(moved, cells)
    # non-Delaunay interior edges according to the oracle (A.7): s = ce + ce' < 0
    ce = m.ce_ratios
    c = m.cells("points")
    s = {}
    for ci in range(c.shape[0]):
        for k in range(3):
            e = tuple(sorted((c[ci, (k + 1) % 3], c[ci, (k + 2) % 3])))
            s.setdefault(e, []).append(ce[k, ci])
    bad = {e for e, v in s.items() if len(v) == 2 and v[0] + v[1] < 0.0}
    assert len(bad) > 10
    _, _, _, _, flagged = step_model(moved, cells, "lloyd", 1.0, bnd)
    # edges with at least one free endpoint with a closed fan are seen by the kernel's chain
    seen = {e for e in bad if not (bnd[e[0]] and bnd[e[1]])}
    assert seen <= flagged
    # and the flag is sharp: nothing far from the criterion is flagged
    ok = {e for e, v in s.items() if len(v) == 2 and v[0] + v[1] > 1e-6 * (abs(v[0]) + abs(v[1]))}
    masked_cells = np.any(ce < -0.5, axis=0)
    touched = set()
    for ci in np.nonzero(masked_cells)[0]:
        for k in range(3):
            touched.add(tuple(sorted((c[ci, (k + 1) % 3], c[ci, (k + 2) % 3]))))
    assert not ((flagged & ok) - touched)
