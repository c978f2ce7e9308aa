"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares,
and the host logic (method names, I/O, CLI parsing, mesh mirror) behaves.  No compute."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__

    __graft_entry__.build()
    from optimesh_b200 import _lib

    return _lib.load()


def header_symbols():
    with open(os.path.join(ROOT, "include", "optimesh_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(om_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    from optimesh_b200 import _lib

    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/optimesh_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms  # the ctypes table covers the header exactly


def test_struct_layout_matches_header():
    from optimesh_b200 import _lib

    # double + 2*int64 + 6*int32 = 48 bytes
    assert ctypes.sizeof(_lib.StepStats) == 48


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (and C++), and a C program
    that takes the address of every declared entry point must link against the library."""
    import re
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = os.path.join(root, "include", "optimesh_b200.h")
    names = sorted(set(re.findall(r"\b(om_[a-z0-9_]+)\s*\(", open(header).read())))
    assert len(names) > 40
    src = tmp_path / "abi.c"
    src.write_text(
        '#include "optimesh_b200.h"\n#include <stdio.h>\n'
        "int main(void) {\n  void* fns[] = {" + ", ".join(f"(void*){n}" for n in names) + "};\n"
        "  unsigned i, ok = 1;\n  for (i = 0; i < sizeof(fns) / sizeof(fns[0]); i++) ok &= fns[i] != 0;\n"
        '  printf("%u %u\\n", (unsigned)sizeof(om_step_stats), ok);\n  return 0;\n}\n')
    exe = tmp_path / "abi"
    libdir = os.path.join(root, "optimesh_b200")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.dirname(header), str(src),
                    "-o", str(exe), "-L", libdir, "-loptimesh_b200", f"-Wl,-rpath,{libdir}"],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out == ["48", "1"]
    gxx = shutil.which("g++")
    if gxx:
        subprocess.run([gxx, "-std=c++11", "-fsyntax-only", "-x", "c++", "-I",
                        os.path.dirname(header), str(src)], check=True, capture_output=True)


def test_no_gpu_fails_loudly(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import optimesh_b200
    from optimesh_b200 import generators as G

    with pytest.raises(RuntimeError, match="no CUDA device"):
        optimesh_b200.optimize_points_cells(*G.SIMPLE1, "lloyd", 1e-2, 10)


def test_method_names():
    from optimesh_b200.mesh import method_id, normalize_method_name

    assert normalize_method_name("CVT (block-diagonal)") == "cvt-block-diagonal"
    assert method_id("Lloyd") == 0 and method_id("CVT (block-diagonal)") == 1
    assert method_id("cpt-fixed-point") == 2 and method_id("ODT (fixed-point)") == 3
    assert method_id("cpt-linear-solve") == 4 and method_id("odt-dp-fp") == 5
    assert method_id("CPT (quasi-newton)") == 6
    for name in ("cvt-full", "cvt-uniform-qnf", "odt-bfgs"):
        with pytest.raises(NotImplementedError):
            method_id(name)
    with pytest.raises(KeyError):
        method_id("laplace")


def test_drop_in_alias_exposes_reference_api():
    import optimesh

    for name in ("optimize_points_cells", "optimize", "get_new_points"):
        assert callable(getattr(optimesh, name))
    assert callable(optimesh.odt.fixed_point) and callable(optimesh.cpt.linear_solve)
    assert callable(optimesh.cvt.quasi_newton_uniform_lloyd)


def test_meshtri_mirror():
    from optimesh_b200 import MeshTri
    from optimesh_b200.main import _mesh_cells

    m = MeshTri([[0, 0], [1, 0], [0, 1]], [[0, 1, 2]])
    assert m.points.dtype == np.float64
    assert np.array_equal(m.cells, [[0, 1, 2]])
    assert np.array_equal(m.cells("points"), [[0, 1, 2]])
    assert np.array_equal(_mesh_cells(m), [[0, 1, 2]])


def test_io_roundtrip(tmp_path):
    from optimesh_b200 import generators as G, io

    pts, cells = G.disk(20, 0)
    for ext in (".vtk", ".npz"):
        path = str(tmp_path / f"m{ext}")
        io.write(path, pts, cells)
        p, c = io.read(path)
        assert np.array_equal(p, pts) and np.array_equal(c, cells)
    sp, sc = G.tetra_sphere(3)
    io.write(str(tmp_path / "s.vtk"), sp, sc)
    p, c = io.read(str(tmp_path / "s.vtk"))
    assert np.array_equal(p, sp) and np.array_equal(c, sc)
    with pytest.raises(ValueError):
        io.read(str(tmp_path / "m.xyz"))


def test_cli_arguments():
    from optimesh_b200.cli import _parser

    a = _parser().parse_args(["in.vtk", "out.vtk", "-m", "lloyd", "--omega", "2.0", "-n", "7",
                              "-t", "1e-3", "-q"])
    assert (a.method, a.omega, a.max_num_steps, a.tolerance, a.quiet) == ("lloyd", 2.0, 7, 1e-3,
                                                                         True)
    assert _parser().parse_args(["a", "b"]).method == "cvt-block-diagonal"


def test_io_cell_data_roundtrip(tmp_path):
    from optimesh_b200 import generators as G, io

    pts, cells = G.square(6, 0.2, 0)
    field = (pts[cells].mean(axis=1)[:, 0] > 0.5).astype(np.int64)
    quality = np.linspace(0.0, 1.0, len(cells))
    for ext in (".vtk", ".npz"):
        path = str(tmp_path / ("m" + ext))
        io.write(path, pts, cells, cell_data={"subdomain": field, "quality": quality})
        p, c, data = io.read(path, with_cell_data=True)
        assert np.allclose(p, pts) and np.array_equal(c, cells)
        assert np.array_equal(data["subdomain"], field) and np.allclose(data["quality"], quality)
        assert len(io.read(path)) == 2  # the two-value form is unchanged
    with pytest.raises(ValueError):
        io.write(str(tmp_path / "bad.npz"), pts, cells, cell_data={"f": field[:-1]})


@pytest.mark.parametrize("ext", [".vtk", ".msh", ".off", ".obj", ".npz"])
def test_builtin_formats_roundtrip(tmp_path, ext):
    from optimesh_b200 import generators as G, io

    for name, (pts, cells) in {"flat": G.square(6, 0.2, 0), "surface": G.tetra_sphere(3)}.items():
        field = (pts[cells].mean(axis=1)[:, 0] > 0.5).astype(np.int64) + 1
        path = str(tmp_path / (name + ext))
        io.write(path, pts, cells, cell_data={"gmsh:physical": field})
        p, c, data = io.read(path, with_cell_data=True)
        assert np.array_equal(p, pts) and np.array_equal(c, cells)  # %.17g round-trips doubles
        if ext in (".vtk", ".msh", ".npz"):
            assert np.array_equal(data["gmsh:physical"], field)


def test_msh_reader_skips_other_elements_and_sparse_node_ids(tmp_path):
    from optimesh_b200 import io

    path = tmp_path / "m.msh"
    path.write_text(
        "$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n4\n"
        "10 0 0 0\n20 1 0 0\n30 1 1 0\n40 0 1 0\n$EndNodes\n$Elements\n4\n"
        "1 15 2 0 1 10\n2 1 2 7 1 10 20\n3 2 2 5 1 10 20 30\n4 2 2 6 2 10 30 40\n"
        "$EndElements\n")
    p, c, data = io.read(str(path), with_cell_data=True)
    assert p.shape == (4, 2) and c.tolist() == [[0, 1, 2], [0, 2, 3]]
    assert data["gmsh:physical"].tolist() == [5, 6] and data["gmsh:geometrical"].tolist() == [1, 2]


def test_cli_subdomains_preserve_interfaces(tmp_path, monkeypatch):
    """`optimesh in out -s NAME` (README.md:17 "preserves submeshes"): every subdomain is
    optimized on its own.  The device call is replaced by the CPU oracle here, so this checks
    the splitting, the write-back and the cell data -- not the kernels."""
    import oracle
    from optimesh_b200 import cli, generators as G, io

    pts, cells = G.square(12, 0.25, 3)
    centroid_x = pts[cells].mean(axis=1)[:, 0]
    field = np.where(centroid_x < 0.35, 3, np.where(centroid_x < 0.7, 5, 9)).astype(np.int64)
    src, dst = str(tmp_path / "in.vtk"), str(tmp_path / "out.vtk")
    io.write(src, pts, cells, cell_data={"gmsh:physical": field})

    calls = []

    def fake(points, cells_, method, tol, max_num_steps, **kwargs):
        calls.append((points.shape[0], cells_.shape[0], method, kwargs.get("omega")))
        assert cells_.min() == 0 and cells_.max() == points.shape[0] - 1  # compact submesh
        kwargs.pop("device", None)
        kwargs.pop("verbose", None)
        return oracle.optimize_points_cells(points, cells_, method, tol, max_num_steps,
                                            omega=kwargs.get("omega", 1.0))

    monkeypatch.setattr(cli, "optimize_points_cells", fake)
    assert cli.main([src, dst, "-m", "cpt-fixed-point", "-n", "5", "-t", "0", "-q",
                     "-s", "gmsh:physical"]) == 0
    assert len(calls) == 3 and sum(c[1] for c in calls) == len(cells)
    p, c, data = io.read(dst, with_cell_data=True)
    assert np.array_equal(data["gmsh:physical"], field)  # every cell kept its subdomain
    # vertices used by more than one subdomain (the interfaces) and the outer boundary stay
    owners = np.zeros((len(pts), 3), dtype=bool)
    for k, v in enumerate((3, 5, 9)):
        owners[c[field == v].reshape(-1), k] = True
        # a subdomain never gains or loses vertices: flips stay inside it
        assert set(c[field == v].reshape(-1)) == set(cells[field == v].reshape(-1))
    shared = owners.sum(axis=1) > 1
    assert shared.sum() > 10 and np.array_equal(p[shared], pts[shared])
    bnd = oracle.MeshTri(pts, cells).is_boundary_point
    assert np.array_equal(p[bnd], pts[bnd])
    assert not np.allclose(p[~shared & ~bnd], pts[~shared & ~bnd])  # the interiors did move
    # without -s the whole mesh is one set and interface vertices move too
    calls.clear()
    assert cli.main([src, dst, "-m", "cpt-fixed-point", "-n", "5", "-t", "0", "-q"]) == 0
    assert len(calls) == 1 and calls[0][:2] == (len(pts), len(cells))
    p2, _ = io.read(dst)
    assert not np.array_equal(p2[shared & ~bnd], pts[shared & ~bnd])
    with pytest.raises(SystemExit):
        cli.main([src, dst, "-q", "-s", "no-such-field"])


class _OracleMesh:
    """Stands in for DeviceMesh in host-logic tests: same methods, CPU oracle inside."""

    def __init__(self, points, cells):
        import oracle

        self._o = oracle
        self.mesh = oracle.MeshTri(np.array(points, dtype=float), np.array(cells))
        self.n, self.dim = self.mesh.points.shape
        self.method, self.omega = None, 1.0

    def set_method(self, method, omega=1.0):
        self.method, self.omega = method, omega

    def set_odt_boundary_barycenters(self, on):
        pass

    def clear_surface(self):
        pass

    def flip_until_delaunay(self):
        return self.mesh.flip_until_delaunay(), 0

    def new_points(self):
        return self._o.get_new_points(self.mesh, self.method)

    def project(self):
        return 0

    @property
    def points(self):
        return self.mesh.points.copy()

    @points.setter
    def points(self, new):
        self.mesh.points = new

    def cells(self, dtype=None):
        return self.mesh.cells("points").astype(dtype or np.int64)

    @property
    def is_boundary_point(self):
        return self.mesh.is_boundary_point


def test_boundary_step_host_path():
    """`boundary_step` (hook of the reference's loop): targets from the mesh object, the
    callback moves the boundary targets, relaxation + limiter on every vertex."""
    import oracle
    from optimesh_b200 import generators as G
    from optimesh_b200.main import _half_min_inradius, _run_loop

    pts, cells = G.disk(40, 3)
    om = oracle.MeshTri(pts, cells)
    want = np.full(len(pts), np.inf)
    np.minimum.at(want, cells.reshape(-1), np.repeat(om.cell_inradius, 3))
    assert np.allclose(_half_min_inradius(pts, cells), 0.5 * want, rtol=1e-13)

    # 1. a callback that returns what it gets changes nothing for methods that pin the
    #    boundary themselves: same trajectory as the plain loop
    for method in ("cpt-fixed-point", "cvt-block-diagonal"):
        log = []
        dm = _OracleMesh(pts, cells)
        steps = _run_loop(dm, method, 0.0, 6, omega=1.0, boundary_step=lambda x: x, log=log)
        rlog = []
        rp, rc = oracle.optimize_points_cells(pts, cells, method, 0.0, 6, log=rlog)
        assert steps == 6 and np.array_equal(dm.cells(), rc)
        assert np.allclose(dm.points, rp, rtol=0, atol=1e-13)
        assert [l["n_limited"] for l in log] == [l["n_limited"] for l in rlog]

    # 2. Lloyd proposes the control-volume centroid for boundary vertices too; projecting it
    #    back onto the circle lets them slide along the boundary
    def to_circle(x):
        return x / np.sqrt(np.einsum("ij,ij->j", x, x))

    dm = _OracleMesh(pts, cells)
    _run_loop(dm, "lloyd", 0.0, 8, omega=1.0, boundary_step=to_circle)
    p, c = dm.points, dm.cells()
    bnd = oracle.MeshTri(pts, cells).is_boundary_point
    assert not np.allclose(p[bnd], pts[bnd])             # they moved ...
    assert np.abs(np.linalg.norm(p[bnd], axis=1) - 1.0).max() < 0.02  # ... near the circle
    a = p[c[:, 1]] - p[c[:, 0]]
    b = p[c[:, 2]] - p[c[:, 0]]
    assert oracle.MeshTri(p, c).num_delaunay_violations() == 0
    assert np.all(np.abs(a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]) > 0)  # no degenerate cell
    with pytest.raises(ValueError):
        _run_loop(_OracleMesh(pts, cells), "lloyd", 0.0, 1, boundary_step=lambda x: x[:, :-1])
    with pytest.raises(TypeError):
        _run_loop(_OracleMesh(pts, cells), "lloyd", 0.0, 1, boundary_step=1.0)


def test_print_stats_runs(capsys):
    from optimesh_b200.helpers import print_stats
    import oracle
    from optimesh_b200 import generators as G

    print_stats(*oracle.stats(oracle.MeshTri(*G.disk(30, 0))))
    out = capsys.readouterr().out
    assert "angles" in out and "quality" in out
