"""``optimesh in out`` (/root/reference/README.md:49-66, :80, :194)."""
from __future__ import annotations

import argparse
import os
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
    p.add_argument("--subdomain-field-name", "-s", default=None, metavar="NAME",
                   help="name of a per-cell field in the input file; every subdomain (cells with "
                        "the same field value) is optimized on its own, so the interfaces "
                        "between submeshes are preserved")
    p.add_argument("--device", type=int, default=0, help="CUDA device")
    p.add_argument("--version", "-v", action="version", version=f"optimesh_b200 {__version__}")
    return p


def _step_format(fmt, k, n_sets):
    """Per-subdomain name of the step dumps: 'step{:03d}.vtk' -> 'step{:03d}_sub1.vtk'."""
    if fmt is None or n_sets == 1:
        return fmt
    root, ext = os.path.splitext(fmt)
    return f"{root}_sub{k}{ext}"


def main(argv=None):
    args = _parser().parse_args(argv)
    points, cells, cell_data = io.read(args.input_file, with_cell_data=True)
    # drop points that no triangle uses (the reference removes orphans before smoothing)
    used = np.zeros(points.shape[0], dtype=bool)
    used[cells.reshape(-1)] = True
    if not used.all():
        remap = np.cumsum(used) - 1
        points, cells = points[used], remap[cells]
    # flat meshes often arrive with a zero z column
    if points.shape[1] == 3 and np.all(points[:, 2] == 0.0):
        points = np.ascontiguousarray(points[:, :2])
    # submeshes (README.md:17): one optimization per subdomain.  The vertices on an interface
    # are boundary vertices of both neighbouring submeshes and therefore stay where they are,
    # and no edge across or along an interface is ever flipped.
    if args.subdomain_field_name is not None:
        if args.subdomain_field_name not in cell_data:
            raise SystemExit(
                f"{args.input_file}: no cell field {args.subdomain_field_name!r} "
                f"(available: {sorted(cell_data) or 'none'})")
        field = np.asarray(cell_data[args.subdomain_field_name])
        if field.ndim != 1:
            raise SystemExit("the subdomain field must hold one value per cell")
        cell_sets = [field == v for v in np.unique(field)]
    else:
        cell_sets = [np.ones(cells.shape[0], dtype=bool)]
    points = np.array(points, dtype=np.float64)
    cells = np.array(cells)
    kwargs = dict(omega=args.omega, verbose=not args.quiet, device=args.device)
    if len(cell_sets) == 1:
        points, cells = optimize_points_cells(
            points, cells, args.method, args.tolerance, args.max_num_steps,
            step_filename_format=args.step_filename_format, **kwargs)
    else:
        for k, in_set in enumerate(cell_sets):
            # compact submesh: its own points, local numbering
            idx, local = np.unique(cells[in_set].reshape(-1), return_inverse=True)
            sub_points, sub_cells = optimize_points_cells(
                points[idx], local.reshape(-1, 3), args.method, args.tolerance,
                args.max_num_steps,
                step_filename_format=_step_format(args.step_filename_format, k, len(cell_sets)),
                **kwargs)
            points[idx] = sub_points
            cells[in_set] = idx[np.asarray(sub_cells)]
    io.write(args.output_file, points, cells, cell_data=cell_data)
    return 0


if __name__ == "__main__":
    sys.exit(main())
