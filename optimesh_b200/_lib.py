"""ctypes binding of liboptimesh_b200.so (the C-ABI in include/optimesh_b200.h).

There is no CPU fallback: if the shared library is missing, loading fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboptimesh_b200.so")

OM_RENUMBER = 1
(OM_LLOYD, OM_CVT_BLOCK_DIAGONAL, OM_CPT_FIXED_POINT, OM_ODT_FIXED_POINT,
 OM_CPT_LINEAR_SOLVE, OM_ODT_DP_FP, OM_CPT_QUASI_NEWTON) = range(7)
(OM_OK, OM_ERR_CUDA, OM_ERR_ARG, OM_ERR_DEGENERATE, OM_ERR_NONMANIFOLD, OM_ERR_INDEX,
 OM_ERR_NOT_CONVERGED) = range(7)


class StepStats(C.Structure):
    _fields_ = [
        ("max_diff2", C.c_double),
        ("n_limited", C.c_int64),
        ("n_flips", C.c_int64),
        ("n_flip_rounds", C.c_int32),
        ("flip_cap_hit", C.c_int32),
        ("is_final", C.c_int32),
        ("solver_iters", C.c_int32),
        ("surface_sweeps", C.c_int32),
        ("reserved", C.c_int32),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}
        d["stale"] = self.reserved  # partitioned runs only (om_set_deferred_commit)
        return d


_H = C.c_void_p
_P = C.POINTER
# name -> (restype, argtypes); every symbol declared in include/optimesh_b200.h
SIGNATURES = {
    "om_last_error": (C.c_char_p, []),
    "om_device_count": (C.c_int, [_P(C.c_int)]),
    "om_create": (C.c_int, [_P(_H), C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int64,
                            C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "om_create_device": (C.c_int, [_P(_H), C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int64,
                                   C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "om_destroy": (C.c_int, [_H]),
    "om_set_method": (C.c_int, [_H, C.c_int, C.c_double]),
    "om_set_limiter": (C.c_int, [_H, C.c_int]),
    "om_set_odt_boundary_barycenters": (C.c_int, [_H, C.c_int]),
    "om_set_surface": (C.c_int, [_H, C.c_int, C.c_double, _P(C.c_double), C.c_int]),
    "om_set_solver": (C.c_int, [_H, C.c_double, C.c_int]),
    "om_flip_until_delaunay": (C.c_int, [_H, C.c_double, C.c_int, _P(C.c_int64), _P(C.c_int32),
                                         _P(C.c_int32)]),
    "om_step": (C.c_int, [_H, C.c_double, _P(StepStats)]),
    "om_update_points": (C.c_int, [_H, C.c_double, _P(StepStats)]),
    "om_project": (C.c_int, [_H, _P(C.c_int32)]),
    "om_run": (C.c_int, [_H, C.c_double, C.c_int64, _P(C.c_int64), _P(StepStats)]),
    "om_run_prepare": (C.c_int, [_H]),
    "om_get_run_totals": (C.c_int, [_H, _P(C.c_int64), _P(C.c_int64), _P(C.c_int64),
                                    _P(C.c_int64)]),
    "om_random_walk": (C.c_int, [_H, C.c_int, C.c_uint64, C.c_double, _P(C.c_int64)]),
    "om_new_points": (C.c_int, [_H, C.c_void_p]),
    "om_targets_device": (C.c_int, [_H, _P(C.c_void_p)]),
    "om_update_from_targets": (C.c_int, [_H, C.c_void_p, C.c_double, _P(StepStats)]),
    "om_solve_graph_laplacian": (C.c_int, [_H, C.c_double, C.c_int, _P(C.c_int32),
                                           _P(C.c_double)]),
    "om_stats": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p]),
    "om_get_points": (C.c_int, [_H, C.c_void_p]),
    "om_set_points": (C.c_int, [_H, C.c_void_p]),
    "om_get_cells": (C.c_int, [_H, C.c_void_p, C.c_int]),
    "om_get_boundary_flags": (C.c_int, [_H, C.c_void_p]),
    "om_device_ptrs": (C.c_int, [_H, _P(C.c_void_p), _P(C.c_void_p), _P(C.c_void_p),
                                 _P(C.c_int32)]),
    "om_pack_points": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_void_p]),
    "om_unpack_points": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_void_p]),
    "om_pin_vertices": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "om_set_owned_range": (C.c_int, [_H, C.c_int64, C.c_int64]),
    "om_flip_check_range": (C.c_int, [_H, C.c_double, C.c_int64, C.c_int64, _P(C.c_int64),
                                      _P(C.c_void_p)]),
    "om_flip_add_records": (C.c_int, [_H, C.c_void_p, C.c_int64]),
    "om_flip_finish": (C.c_int, [_H, C.c_double, C.c_int, _P(C.c_int64), _P(C.c_int32),
                                 _P(C.c_int32)]),
    "om_points_device": (C.c_int, [_H, _P(C.c_void_p), _P(C.c_int64), _P(C.c_int32)]),
    "om_cell_range_of_vertices": (C.c_int, [_H, C.c_int64, C.c_int64, _P(C.c_int64),
                                            _P(C.c_int64)]),
    "om_set_deferred_commit": (C.c_int, [_H, C.c_int]),
    "om_commit_points": (C.c_int, [_H]),
    "om_coords_invalidate": (C.c_int, [_H]),
    "om_coords_all_valid": (C.c_int, [_H]),
    "om_band_build": (C.c_int, [_H, C.c_int, _P(C.c_int64), _P(C.c_void_p)]),
    "om_band_pack": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_void_p]),
    "om_band_unpack": (C.c_int, [_H, C.c_void_p, C.c_int64, C.c_void_p]),
    "om_flip_pass_begin": (C.c_int, [_H]),
    "om_flip_round_check": (C.c_int, [_H, C.c_double, C.c_int, C.c_int64, C.c_int64,
                                      _P(C.c_int64), _P(C.c_void_p), _P(C.c_int32)]),
    "om_flip_round_apply": (C.c_int, [_H, C.c_int64, _P(C.c_int64), _P(C.c_int64)]),
    "om_flip_pass_end": (C.c_int, [_H, _P(C.c_int64), _P(C.c_int32)]),
    "om_flip_round_pack": (C.c_int, [_H, C.c_int32, C.c_void_p]),
    "om_flip_round_apply_gathered": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32, _P(C.c_int64),
                                               _P(C.c_int64), _P(C.c_int32), _P(C.c_int64)]),
    "om_shared_begin": (C.c_int, [_H, C.c_int, C.c_int, _P(_H), _P(C.c_int32), _P(C.c_int32)]),
    "om_shared_map": (C.c_int, [_H, _P(C.c_int32), C.c_int32]),
    "om_shared_run": (C.c_int, [_H, C.c_double, C.c_int64, _P(C.c_int64), _P(StepStats)]),
    "om_shared_prepare": (C.c_int, [_H]),
    "om_shared_time_update": (C.c_int, [_H, C.c_int, _P(C.c_double)]),
    "om_shared_info": (C.c_int, [_H, _P(C.c_int64), _P(C.c_int64), _P(C.c_int64),
                                 _P(C.c_int64)]),
    "om_set_timing": (C.c_int, [_H, C.c_int]),
    "om_get_phase_timing": (C.c_int, [_H, C.c_void_p, _P(C.c_int64)]),
    "om_get_timing": (C.c_int, [_H, _P(C.c_double), _P(C.c_int64), _P(C.c_double),
                                _P(C.c_int64)]),
    "om_release_cached_memory": (C.c_int, [C.c_int]),
    "om_result_alloc": (C.c_int, [C.c_int64, _P(C.c_void_p)]),
    "om_result_free": (C.c_int, [C.c_void_p, C.c_int64]),
    "om_prefault_host": (C.c_int, [C.c_void_p, C.c_int64]),
    "om_launch_count": (C.c_int, [_H, _P(C.c_int64)]),
    "om_synchronize": (C.c_int, [_H]),
    "om_stream": (C.c_int, [_H, _P(C.c_void_p)]),
}

_lib = None


def load():
    """Loads the shared library (once) and sets the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m optimesh_b200.build` "
            "(optimesh_b200 has no CPU fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class OptimeshError(RuntimeError):
    pass


class DegenerateCellsError(ValueError):
    """Zero-area cell (upstream: MeshplexError("Degenerate cells."))."""


class MeshTopologyError(ValueError):
    pass


def check(rc: int):
    if rc == OM_OK:
        return
    msg = load().om_last_error().decode(errors="replace")
    if rc == OM_ERR_DEGENERATE:
        raise DegenerateCellsError(msg)
    if rc in (OM_ERR_NONMANIFOLD, OM_ERR_INDEX):
        raise MeshTopologyError(msg)
    if rc == OM_ERR_ARG:
        raise ValueError(msg)
    raise OptimeshError(msg)
