"""Legacy per-method entry points, e.g. ``optimesh.odt.fixed_point(X, cells, 1e-2, 100)``
(/root/reference/README.md:234-240)."""
from .main import optimize_points_cells


def fixed_point(points, cells, tol, max_num_steps, **kwargs):
    return optimize_points_cells(points, cells, "odt-fixed-point", tol, max_num_steps, **kwargs)
