"""Angle and quality histograms (oracle, test infrastructure).

Restates ``optimesh.helpers.print_stats`` -- the two histograms shown by the CLI
(/root/reference/README.md:55-60; quality = 2 r_in / r_circ, max 1: README.md:203-204)
per SURVEY.md Appendix A.11.  Parity unpinned -- see oracle/__init__.py.
"""
from __future__ import annotations

import numpy as np

from .meshtri import MeshTri


def stats(mesh: MeshTri):
    """Returns (angle_hist[72], q_hist[40], summary dict)."""
    angles = mesh.angles / np.pi * 180.0
    angle_hist, _ = np.histogram(angles, bins=np.linspace(0.0, 180.0, num=73, endpoint=True))
    q = mesh.q_radius_ratio
    q_hist, _ = np.histogram(q, bins=np.linspace(0.0, 1.0, num=41, endpoint=True))
    summary = dict(
        angle_min=float(angles.min()),
        angle_max=float(angles.max()),
        angle_avg=float(angles.mean()),
        angle_std=float(angles.std()),
        q_min=float(q.min()),
        q_avg=float(q.mean()),
        q_max=float(q.max()),
        q_std=float(q.std()),
    )
    return angle_hist.astype(np.int64), q_hist.astype(np.int64), summary
