"""CPU oracle for the optimesh smoothing step (TEST INFRASTRUCTURE, not product code).

This package is a plain numpy/scipy restatement of the arithmetic of the reference's
hot path: ``optimesh.optimize_points_cells`` -> per-step ``get_new_points`` (relaxed
Lloyd, CVT block-diagonal, CPT fixed-point / quasi-Newton / linear-solve, ODT fixed-point
in its uniform and density-preserving forms), the
driver loop (pin boundary, omega relaxation, step limiter, surface projection) and
meshplex's ``MeshTri`` geometry + ``flip_until_delaunay``.

Reference citations: the mounted reference tree holds only ``README.md``; the API
names and call sites followed here are /root/reference/README.md:124-126
(``optimize_points_cells``), :131-133 (``meshplex.MeshTri`` + ``optimize``), :141
(``get_new_points``), :157-162 (implicit-surface protocol), :80/:90/:104 (method
names), :55-60 (statistics).  The arithmetic itself lives in optimesh and in the
third-party, un-vendored ``meshplex`` (no pinned version in the tree); it is restated
from the published algorithms (README.md:244-251) per SURVEY.md Appendix A.

PARITY UNPINNED: the reference source and its tests are absent from
/root/reference, and the package cannot be installed offline (licence-gated).  What
pins this oracle instead (tests/test_oracle.py): the recollected upstream
known-answer literals for the 5-point "simple1" mesh, Qhull
(``scipy.spatial.Delaunay``/``ConvexHull``) as an independent topology oracle,
Qhull's Voronoi cells (``scipy.spatial.Voronoi``) for control volumes and their centroids,
central differences of the published ODT energy for the ODT target,
``scipy.sparse.linalg.spsolve`` for the CPT solves, and analytic invariants
(equivariance under similarity transforms and relabelling, conservation of area,
hexagonal-patch fixed points).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product (``optimesh_b200``)
never does.
"""
from .meshtri import MeshTri, DegenerateCellsError  # noqa: F401
from .methods import get_new_points, normalize_method_name, METHODS  # noqa: F401
from .driver import optimize, optimize_points_cells  # noqa: F401
from .stats import stats  # noqa: F401
