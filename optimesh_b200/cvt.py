"""Legacy per-method entry points for CVT (/root/reference/README.md:80, :234-240)."""
from .main import optimize_points_cells


def lloyd(points, cells, tol, max_num_steps, omega=1.0, **kwargs):
    return optimize_points_cells(points, cells, "lloyd", tol, max_num_steps, omega=omega, **kwargs)


def block_diagonal(points, cells, tol, max_num_steps, **kwargs):
    return optimize_points_cells(points, cells, "cvt-block-diagonal", tol, max_num_steps,
                                 **kwargs)


quasi_newton_uniform_lloyd = lloyd
quasi_newton_uniform_blocks = block_diagonal
