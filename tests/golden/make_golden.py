"""Regenerates the fixtures in tests/golden/ from the CPU oracle.

The reference (optimesh/meshplex) cannot be imported in this container (source absent,
licence-gated: SURVEY.md section 0), so these vectors come from ``oracle/`` -- which is
itself pinned against the recollected upstream literals in tests/test_oracle.py.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from oracle.meshtri import MeshTri  # noqa: E402
from optimesh_b200 import generators as G  # noqa: E402


def norms(p):
    return [float(np.abs(p).sum()), float(np.sqrt((p * p).sum())), float(np.abs(p).max())]


def config1():
    pts, cells = G.disk(120, 0)
    log = []
    p, c = oracle.optimize_points_cells(pts, cells, "lloyd", 1.0e-5, 50, log=log)
    ah, qh, s = oracle.stats(MeshTri(p, c))
    out = dict(
        steps=len(log),
        n_flips=[l["n_flips"] for l in log],
        n_rounds=[l["n_rounds"] for l in log],
        n_limited=[l["n_limited"] for l in log],
        max_diff2=[l["max_diff2"] for l in log],
        norms=norms(p),
        angle_hist=ah.tolist(),
        q_hist=qh.tolist(),
        summary=s,
    )
    with open(os.path.join(HERE, "config1_lloyd.json"), "w") as f:
        json.dump(out, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "config1_lloyd_final.npz"), points=p,
                        cells=c.astype(np.int32))


def single_steps():
    """One step of every method from the same state, 2D and surface (small meshes)."""
    out = {}
    pts, cells = G.disk(40, 3)
    for m in oracle.METHODS:
        out[f"disk40_{m}"] = oracle.get_new_points(MeshTri(pts, cells), m)
    sp, sc = G.tetra_sphere(6)
    rs = np.random.RandomState(5)
    sp = sp + rs.normal(scale=0.02, size=sp.shape)
    sp /= np.linalg.norm(sp, axis=1)[:, None]
    mesh = MeshTri(sp, sc)
    mesh.flip_until_delaunay()
    out["sphere6_points"] = sp
    out["sphere6_cells"] = mesh.cells("points").astype(np.int32)
    for m in oracle.METHODS:
        out[f"sphere6_{m}"] = oracle.get_new_points(MeshTri(sp, mesh.cells("points")), m)
    np.savez_compressed(os.path.join(HERE, "single_steps.npz"), **out)


if __name__ == "__main__":
    config1()
    single_steps()
    print("golden fixtures written to", HERE)
