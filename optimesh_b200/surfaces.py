"""Implicit surfaces with a device fast path.

Any object with ``f(x)`` and ``grad(x)`` on ``(3, n)`` arrays is accepted by
``optimize_points_cells(..., implicit_surface=...)`` (/root/reference/README.md:157-162);
``Sphere`` additionally tells the library to project on the GPU.
"""
from __future__ import annotations

import numpy as np


class Sphere:
    """f(x) = R^2 - |x - c|^2; ``Sphere()`` is the README's unit sphere."""

    def __init__(self, center=(0.0, 0.0, 0.0), radius: float = 1.0):
        self.center = tuple(float(c) for c in center)
        self.radius = float(radius)

    def f(self, x):
        c = np.asarray(self.center)[:, None]
        d = x - c
        return self.radius ** 2 - (d[0] ** 2 + d[1] ** 2 + d[2] ** 2)

    def grad(self, x):
        c = np.asarray(self.center)[:, None]
        return -2 * (x - c)
