"""Legacy per-method entry points for CPT (/root/reference/README.md:90, :234-240)."""
from .main import optimize_points_cells


def fixed_point(points, cells, tol, max_num_steps, **kwargs):
    return optimize_points_cells(points, cells, "cpt-fixed-point", tol, max_num_steps, **kwargs)


def linear_solve(points, cells, tol, max_num_steps, **kwargs):
    return optimize_points_cells(points, cells, "cpt-linear-solve", tol, max_num_steps, **kwargs)


def quasi_newton(points, cells, tol, max_num_steps, **kwargs):
    return optimize_points_cells(points, cells, "cpt-quasi-newton", tol, max_num_steps, **kwargs)
