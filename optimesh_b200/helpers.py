"""Terminal statistics (the two histograms of /root/reference/README.md:55-60)."""
from __future__ import annotations

import numpy as np


def _bars(hist, width=30):
    m = max(int(np.max(hist)), 1)
    return ["#" * int(round(width * int(h) / m)) for h in hist]


def print_stats(angle_hist, q_hist, summary, file=None):
    """Angle distribution (72 bins of 2.5 deg, printed as 18 rows of 10 deg) and quality
    distribution (2 r_in / r_circ, 40 bins printed as 20 rows)."""
    a = np.asarray(angle_hist).reshape(18, 4).sum(axis=1)
    q = np.asarray(q_hist).reshape(20, 2).sum(axis=1)
    ab, qb = _bars(a), _bars(q)
    lines = [f"{'angles (deg)':<46}{'quality (2 r_in / r_circ)'}"]
    for i in range(20):
        left = f"{10 * i:>4}-{10 * (i + 1):<4} {a[i]:>9d} {ab[i]:<30}" if i < 18 else " " * 50
        right = f"{0.05 * i:>4.2f}-{0.05 * (i + 1):<4.2f} {q[i]:>9d} {qb[i]}"
        lines.append(f"{left[:50]:<50}{right}")
    lines.append(
        "angle min/avg/max/std: {angle_min:.3f} / {angle_avg:.3f} / {angle_max:.3f} / "
        "{angle_std:.3f}   quality min/avg/max: {q_min:.4f} / {q_avg:.4f} / {q_max:.4f}".format(
            **summary)
    )
    print("\n".join(lines), file=file)
