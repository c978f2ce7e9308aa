"""``optimesh in out`` (/root/reference/README.md:49-66, :80, :194)."""
from __future__ import annotations

import argparse
import sys

import numpy as np

from . import io
from .__about__ import __version__
from .main import optimize_points_cells
from .mesh import METHOD_IDS, NOT_IMPLEMENTED


def _parser():
    p = argparse.ArgumentParser(prog="optimesh", description="Triangular mesh optimization (B200).")
    p.add_argument("input_file", metavar="INPUT_FILE", help="input mesh file")
    p.add_argument("output_file", metavar="OUTPUT_FILE", help="output mesh file")
    p.add_argument("--method", "-m", default="cvt-block-diagonal",
                   choices=sorted(METHOD_IDS) + sorted(NOT_IMPLEMENTED),
                   help="smoothing method (the reference's default cvt-full is not part of this "
                        "build; default here: cvt-block-diagonal)")
    p.add_argument("--omega", type=float, default=1.0, help="relaxation parameter (default 1.0)")
    p.add_argument("--max-num-steps", "-n", type=int, default=100, help="maximum number of steps")
    p.add_argument("--tolerance", "-t", type=float, default=1.0e-5, help="convergence tolerance")
    p.add_argument("--quiet", "-q", action="store_true", help="no statistics output")
    p.add_argument("--step-filename-format", "-f", default=None,
                   help="dump the mesh after every step, e.g. 'step{:03d}.vtk'")
    p.add_argument("--device", type=int, default=0, help="CUDA device")
    p.add_argument("--version", "-v", action="version", version=f"optimesh_b200 {__version__}")
    return p


def main(argv=None):
    args = _parser().parse_args(argv)
    points, cells = io.read(args.input_file)
    # drop points that no triangle uses (the reference removes orphans before smoothing)
    used = np.zeros(points.shape[0], dtype=bool)
    used[cells.reshape(-1)] = True
    if not used.all():
        remap = np.cumsum(used) - 1
        points, cells = points[used], remap[cells]
    # flat meshes often arrive with a zero z column
    if points.shape[1] == 3 and np.all(points[:, 2] == 0.0):
        points = np.ascontiguousarray(points[:, :2])
    points, cells = optimize_points_cells(
        points, cells, args.method, args.tolerance, args.max_num_steps, omega=args.omega,
        verbose=not args.quiet, step_filename_format=args.step_filename_format,
        device=args.device)
    io.write(args.output_file, points, cells)
    return 0


if __name__ == "__main__":
    sys.exit(main())
