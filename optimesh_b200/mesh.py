"""Host-side mesh objects.

``DeviceMesh`` owns one ``om_handle`` (the mesh resident in HBM).  ``MeshTri`` mirrors
the small part of ``meshplex.MeshTri`` that the reference API exposes
(/root/reference/README.md:128-133: constructor, ``.points``, ``.cells``) so that
``optimesh.optimize(mesh, ...)`` keeps working without meshplex.
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _lib
from ._lib import check

METHOD_IDS = {
    "lloyd": _lib.OM_LLOYD,
    "cvt-block-diagonal": _lib.OM_CVT_BLOCK_DIAGONAL,
    "cpt-fixed-point": _lib.OM_CPT_FIXED_POINT,
    "odt-fixed-point": _lib.OM_ODT_FIXED_POINT,
    "cpt-linear-solve": _lib.OM_CPT_LINEAR_SOLVE,
    "odt-dp-fp": _lib.OM_ODT_DP_FP,
    "cpt-quasi-newton": _lib.OM_CPT_QUASI_NEWTON,
}
# names the reference knows (README.md:80, :90, :104, :194) that are outside this build
NOT_IMPLEMENTED = ("cvt-full", "cvt-uniform-qnf", "odt-bfgs")


def normalize_method_name(name: str) -> str:
    """'CVT (block-diagonal)' -> 'cvt-block-diagonal' (README.md:80 vs :125)."""
    return "-".join(name.lower().replace("(", "").replace(")", "").split())


def method_id(name: str) -> int:
    key = normalize_method_name(name)
    if key in NOT_IMPLEMENTED:
        raise NotImplementedError(
            f"method {key!r} is known to optimesh but not part of this build; "
            f"available: {sorted(METHOD_IDS)}"
        )
    if key not in METHOD_IDS:
        raise KeyError(f"unknown method {name!r}; available: {sorted(METHOD_IDS)}")
    return METHOD_IDS[key]


class _PinnedBlock:
    """A block of the library's pinned result cache, exposed through the array interface; it
    goes back to the cache when the numpy array built on it is garbage collected."""

    def __init__(self, lib, ptr, nbytes, shape, typestr):
        self._lib, self._ptr, self._nbytes = lib, ptr, nbytes
        self.__array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                    "data": (ptr, False), "version": 3}

    def __del__(self):
        try:
            self._lib.om_result_free(self._ptr, self._nbytes)
        except Exception:
            pass


def result_array(shape, dtype):
    """An uninitialised array for a result: backed by the pinned result cache when the result
    is large and the cache has room (om_result_alloc), an ordinary numpy array otherwise."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    if nbytes >= (8 << 20):
        lib = _lib.load()
        p = C.c_void_p()
        if lib.om_result_alloc(nbytes, C.byref(p)) == 0 and p.value:
            return np.asarray(_PinnedBlock(lib, p.value, nbytes, shape, dtype.str))
    return np.empty(shape, dtype=dtype)


class DeviceMesh:
    """A triangular mesh resident on one GPU."""

    def __init__(self, points, cells, device: int = 0, renumber: bool = True, stream=None):
        lib = _lib.load()
        points = np.ascontiguousarray(points, dtype=np.float64)
        cells = np.asarray(cells)
        if points.ndim != 2 or points.shape[1] not in (2, 3):
            raise ValueError("points must have shape (N, 2) or (N, 3)")
        if cells.ndim != 2 or cells.shape[1] != 3:
            raise ValueError("cells must have shape (C, 3)")
        if not np.issubdtype(cells.dtype, np.integer):
            raise ValueError("cells must be an integer array")
        self.cells_dtype = cells.dtype
        if cells.dtype.itemsize not in (4, 8) or cells.dtype.kind == "u" and cells.dtype.itemsize == 8:
            cells = cells.astype(np.int64)
        cells = np.ascontiguousarray(cells)
        self.n, self.dim = points.shape
        self.c = cells.shape[0]
        self._h = C.c_void_p()
        self._lib = lib
        check(lib.om_create(C.byref(self._h), device, stream, self.n, self.dim, self.c,
                            points.ctypes.data, cells.ctypes.data, cells.dtype.itemsize,
                            _lib.OM_RENUMBER if renumber else 0))

    @classmethod
    def from_torch(cls, points, cells, renumber: bool = True, stream=None):
        """Builds the mesh from CUDA tensors already resident in HBM (float64 (N,d) and
        int32/int64 (C,3), contiguous) through om_create_device: no host round trip."""
        lib = _lib.load()
        if not (points.is_cuda and cells.is_cuda and points.is_contiguous()
                and cells.is_contiguous()):
            raise ValueError("from_torch needs contiguous CUDA tensors")
        self = cls.__new__(cls)
        self.n, self.dim = int(points.shape[0]), int(points.shape[1])
        self.c = int(cells.shape[0])
        itemsize = cells.element_size()
        self.cells_dtype = np.dtype(np.int32 if itemsize == 4 else np.int64)
        self._h = C.c_void_p()
        self._lib = lib
        check(lib.om_create_device(C.byref(self._h), points.device.index or 0, stream, self.n,
                                   self.dim, self.c, points.data_ptr(), cells.data_ptr(),
                                   itemsize, _lib.OM_RENUMBER if renumber else 0))
        return self

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.om_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- settings
    def set_method(self, method: str, omega: float = 1.0):
        check(self._lib.om_set_method(self._h, method_id(method), float(omega)))

    def set_limiter(self, on: bool):
        check(self._lib.om_set_limiter(self._h, int(bool(on))))

    def set_odt_boundary_barycenters(self, on: bool):
        """ODT methods: cells with a boundary edge contribute their barycenter (default) or,
        if off, their circumcenter like every other cell."""
        check(self._lib.om_set_odt_boundary_barycenters(self._h, int(bool(on))))

    def set_sphere(self, center=(0.0, 0.0, 0.0), radius=1.0, tol=1.0e-10, max_sweeps=100):
        params = (C.c_double * 4)(center[0], center[1], center[2], radius)
        check(self._lib.om_set_surface(self._h, 1, float(tol), params, int(max_sweeps)))

    def clear_surface(self):
        check(self._lib.om_set_surface(self._h, 0, 0.0, None, 0))

    def set_solver(self, rtol=1.0e-13, max_iter=100000):
        check(self._lib.om_set_solver(self._h, float(rtol), int(max_iter)))

    # -- the hot path
    def flip_until_delaunay(self, tol: float = 0.0, max_steps: int = 100):
        nf, nr, cap = C.c_int64(), C.c_int32(), C.c_int32()
        check(self._lib.om_flip_until_delaunay(self._h, float(tol), int(max_steps), C.byref(nf),
                                               C.byref(nr), C.byref(cap)))
        if cap.value:
            warnings.warn("Maximum number of edge flips reached.")
        return nf.value, nr.value

    def step(self, tol: float = 0.0) -> dict:
        st = _lib.StepStats()
        check(self._lib.om_step(self._h, float(tol), C.byref(st)))
        if st.flip_cap_hit:
            warnings.warn("Maximum number of edge flips reached.")
        return st.as_dict()

    def update_points(self, tol: float = 0.0) -> dict:
        st = _lib.StepStats()
        check(self._lib.om_update_points(self._h, float(tol), C.byref(st)))
        return st.as_dict()

    def project(self) -> int:
        n = C.c_int32()
        check(self._lib.om_project(self._h, C.byref(n)))
        return n.value

    def run(self, tol: float, max_num_steps: int):
        steps = C.c_int64()
        st = _lib.StepStats()
        check(self._lib.om_run(self._h, float(tol), int(max_num_steps), C.byref(steps),
                               C.byref(st)))
        if st.flip_cap_hit:
            warnings.warn("Maximum number of edge flips reached.")
        return steps.value, st.as_dict()

    def run_prepare(self):
        """Builds the CUDA graph `run` would launch (keeps the build out of a timed region)."""
        check(self._lib.om_run_prepare(self._h))

    def run_totals(self) -> dict:
        """Flips, flip rounds and limited vertices summed over the steps of the last `run`."""
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        check(self._lib.om_get_run_totals(self._h, C.byref(a), C.byref(b), C.byref(c),
                                          C.byref(d)))
        return dict(n_flips=a.value, n_flip_rounds=b.value, n_limited=c.value,
                    n_deferred=d.value)

    def random_walk(self, rounds: int, seed: int = 0, amplitude: float = 1.0) -> int:
        """Synthetic workloads: `rounds` random moves bounded by half the smallest incident
        inradius, each followed by flip-until-Delaunay (om_random_walk).  Returns the flips."""
        nf = C.c_int64()
        check(self._lib.om_random_walk(self._h, int(rounds), int(seed), float(amplitude),
                                       C.byref(nf)))
        return nf.value

    def new_points(self) -> np.ndarray:
        out = np.empty((self.n, self.dim), dtype=np.float64)
        check(self._lib.om_new_points(self._h, out.ctypes.data))
        return out

    def targets_device(self) -> int:
        """Device pointer of the un-relaxed targets of the current method (what `new_points`
        returns), internal numbering and layout like `points_device`; owned by the handle,
        valid until the next call.  The caller may overwrite entries."""
        p = C.c_void_p()
        check(self._lib.om_targets_device(self._h, C.byref(p)))
        return p.value

    def update_from_targets(self, targets_ptr: int, tol: float = 0.0) -> dict:
        """x <- x + limiter(omega (target - x)) for every vertex, boundary vertices included,
        from a device buffer laid out like `targets_device`."""
        st = _lib.StepStats()
        check(self._lib.om_update_from_targets(self._h, C.c_void_p(targets_ptr), float(tol),
                                               C.byref(st)))
        return st.as_dict()

    def device_ptrs(self):
        """(points, cells4, perm, point_stride): raw device pointers, internal numbering; perm
        (internal -> caller vertex id) is None for the identity."""
        a, b, c, s = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int32()
        check(self._lib.om_device_ptrs(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(s)))
        return a.value, b.value, c.value, s.value

    def solve_graph_laplacian(self, rtol=1.0e-13, max_iter=100000):
        it, res = C.c_int32(), C.c_double()
        check(self._lib.om_solve_graph_laplacian(self._h, float(rtol), int(max_iter),
                                                 C.byref(it), C.byref(res)))
        return it.value, res.value

    def stats(self):
        ah = np.zeros(72, dtype=np.int64)
        qh = np.zeros(40, dtype=np.int64)
        s = np.zeros(8, dtype=np.float64)
        check(self._lib.om_stats(self._h, ah.ctypes.data, qh.ctypes.data, s.ctypes.data))
        keys = ("angle_min", "angle_max", "angle_avg", "angle_std", "q_min", "q_avg", "q_max",
                "q_std")
        return ah, qh, dict(zip(keys, s.tolist()))

    # -- data movement
    @property
    def points(self) -> np.ndarray:
        return self.get_points()

    def get_points(self, out=None) -> np.ndarray:
        if out is None:
            out = result_array((self.n, self.dim), np.float64)
        assert out.shape == (self.n, self.dim) and out.dtype == np.float64 and out.flags.c_contiguous
        check(self._lib.om_get_points(self._h, out.ctypes.data))
        return out

    @points.setter
    def points(self, new):
        new = np.ascontiguousarray(new, dtype=np.float64)
        if new.shape != (self.n, self.dim):
            raise ValueError(f"points must have shape {(self.n, self.dim)}")
        check(self._lib.om_set_points(self._h, new.ctypes.data))

    @staticmethod
    def cells_wire_dtype(dtype):
        """int32 or int64: what om_get_cells writes for a result of `dtype`."""
        dtype = np.dtype(dtype)
        return np.dtype(np.int32 if dtype.itemsize <= 4 and dtype != np.uint32 else np.int64)

    def cells(self, dtype=None, out=None) -> np.ndarray:
        dtype = np.dtype(dtype or self.cells_dtype)
        wire = self.cells_wire_dtype(dtype)
        if out is None:
            out = result_array((self.c, 3), wire)
        assert out.shape == (self.c, 3) and out.dtype == wire and out.flags.c_contiguous
        check(self._lib.om_get_cells(self._h, out.ctypes.data, out.dtype.itemsize))
        return out.astype(dtype, copy=False)

    @property
    def is_boundary_point(self) -> np.ndarray:
        out = np.zeros(self.n, dtype=np.uint8)
        check(self._lib.om_get_boundary_flags(self._h, out.ctypes.data))
        return out.astype(bool)

    def pin_vertices(self, idx):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        check(self._lib.om_pin_vertices(self._h, idx.ctypes.data, idx.size))

    def flip_check_range(self, cell_lo: int, cell_hi: int, tol: float = 0.0):
        """Sharded first flip round: returns (device pointer of the records, count)."""
        n, p = C.c_int64(), C.c_void_p()
        check(self._lib.om_flip_check_range(self._h, float(tol), int(cell_lo), int(cell_hi),
                                            C.byref(n), C.byref(p)))
        return p.value, n.value

    def flip_add_records(self, ptr: int, n: int):
        check(self._lib.om_flip_add_records(self._h, C.c_void_p(ptr), int(n)))

    def flip_finish(self, tol: float = 0.0, max_steps: int = 100):
        nf, nr, cap = C.c_int64(), C.c_int32(), C.c_int32()
        check(self._lib.om_flip_finish(self._h, float(tol), int(max_steps), C.byref(nf),
                                       C.byref(nr), C.byref(cap)))
        if cap.value:
            warnings.warn("Maximum number of edge flips reached.")
        return nf.value, nr.value

    # -- partitioned coordinates (dist.py)
    def cell_range_of_vertices(self, vlo: int, vhi: int):
        a, b = C.c_int64(), C.c_int64()
        check(self._lib.om_cell_range_of_vertices(self._h, int(vlo), int(vhi), C.byref(a),
                                                  C.byref(b)))
        return a.value, b.value

    def set_deferred_commit(self, on: bool):
        check(self._lib.om_set_deferred_commit(self._h, int(bool(on))))

    def commit_points(self):
        check(self._lib.om_commit_points(self._h))

    def coords_invalidate(self):
        check(self._lib.om_coords_invalidate(self._h))

    def coords_all_valid(self):
        check(self._lib.om_coords_all_valid(self._h))

    def band_build(self, depth: int = 3):
        n, p = C.c_int64(), C.c_void_p()
        check(self._lib.om_band_build(self._h, int(depth), C.byref(n), C.byref(p)))
        return p.value, n.value

    def band_pack(self, idx_ptr: int, n: int, buf_ptr: int):
        check(self._lib.om_band_pack(self._h, C.c_void_p(idx_ptr), int(n), C.c_void_p(buf_ptr)))

    def band_unpack(self, idx_ptr: int, n: int, buf_ptr: int):
        check(self._lib.om_band_unpack(self._h, C.c_void_p(idx_ptr), int(n), C.c_void_p(buf_ptr)))

    def flip_pass_begin(self):
        check(self._lib.om_flip_pass_begin(self._h))

    def flip_round_check(self, first: bool, cell_lo: int, cell_hi: int, tol: float = 0.0):
        n, p, st = C.c_int64(), C.c_void_p(), C.c_int32()
        check(self._lib.om_flip_round_check(self._h, float(tol), int(bool(first)), int(cell_lo),
                                            int(cell_hi), C.byref(n), C.byref(p), C.byref(st)))
        return p.value, n.value, bool(st.value)

    def flip_round_check_nofetch(self, first: bool, cell_lo: int, cell_hi: int, tol: float = 0.0):
        check(self._lib.om_flip_round_check(self._h, float(tol), int(bool(first)), int(cell_lo),
                                            int(cell_hi), None, None, None))

    def flip_round_pack(self, capacity: int, slot_ptr: int):
        check(self._lib.om_flip_round_pack(self._h, int(capacity), C.c_void_p(slot_ptr)))

    def flip_round_apply_gathered(self, gathered_ptr: int, n_ranks: int, capacity: int):
        nc, nf, ab, own = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int64()
        check(self._lib.om_flip_round_apply_gathered(self._h, C.c_void_p(gathered_ptr),
                                                     int(n_ranks), int(capacity), C.byref(nc),
                                                     C.byref(nf), C.byref(ab), C.byref(own)))
        return nc.value, nf.value, ab.value, own.value

    def flip_round_apply(self, total_records: int):
        nc, nf = C.c_int64(), C.c_int64()
        check(self._lib.om_flip_round_apply(self._h, int(total_records), C.byref(nc),
                                            C.byref(nf)))
        return nc.value, nf.value

    def flip_pass_end(self):
        nf, nr = C.c_int64(), C.c_int32()
        check(self._lib.om_flip_pass_end(self._h, C.byref(nf), C.byref(nr)))
        return nf.value, nr.value

    def set_owned_range(self, lo: int, hi: int):
        check(self._lib.om_set_owned_range(self._h, int(lo), int(hi)))

    def points_device(self):
        """(device pointer, allocated vertices, stride) of the internal point array."""
        p, n, s = C.c_void_p(), C.c_int64(), C.c_int32()
        check(self._lib.om_points_device(self._h, C.byref(p), C.byref(n), C.byref(s)))
        return p.value, n.value, s.value

    def set_timing(self, on: bool = True):
        check(self._lib.om_set_timing(self._h, int(bool(on))))

    def phase_timing(self) -> dict:
        """ms per loop iteration of the phases of the timed (stream-driven) loop."""
        ph = (C.c_double * 5)()
        n = C.c_int64()
        check(self._lib.om_get_phase_timing(self._h, ph, C.byref(n)))
        k = max(n.value, 1)
        names = ("reset_and_ring_kernel", "k_post", "flag_check", "flip_rounds",
                 "rings_recompute_stats")
        return {"iterations": n.value, **{m: ph[i] / k for i, m in enumerate(names)}}

    def timing(self) -> dict:
        a, b = C.c_double(), C.c_double()
        na, nb = C.c_int64(), C.c_int64()
        check(self._lib.om_get_timing(self._h, C.byref(a), C.byref(na), C.byref(b), C.byref(nb)))
        return dict(step_kernel_ms=a.value, step_kernel_launches=na.value,
                    flip_pass_ms=b.value, flip_passes=nb.value)

    @property
    def launch_count(self) -> int:
        n = C.c_int64()
        check(self._lib.om_launch_count(self._h, C.byref(n)))
        return n.value

    def synchronize(self):
        check(self._lib.om_synchronize(self._h))

    @property
    def stream(self) -> int:
        p = C.c_void_p()
        check(self._lib.om_stream(self._h, C.byref(p)))
        return p.value or 0

    @property
    def handle(self):
        return self._h


class _Cells(np.ndarray):
    """ndarray that can also be called like meshplex's ``mesh.cells("points")``."""

    def __call__(self, which="points"):
        if which != "points":
            raise KeyError(which)
        return np.asarray(self)


class MeshTri:
    """Minimal stand-in for ``meshplex.MeshTri(points, cells)`` (README.md:131)."""

    def __init__(self, points, cells):
        self.points = np.array(points, dtype=np.float64)
        self._set_cells(cells)

    def _set_cells(self, cells):
        self._cells = np.array(cells).view(_Cells)

    @property
    def cells(self):
        return self._cells
