"""optimesh_b200 -- B200-native smoothing step of meshpro/optimesh.

Same API as the reference for this path (/root/reference/README.md:119-142, :239):

    points, cells = optimesh_b200.optimize_points_cells(points, cells, "lloyd", 1e-5, 100,
                                                        omega=2.0)

The compute runs in hand-written sm_100a CUDA behind a C-ABI (include/optimesh_b200.h,
liboptimesh_b200.so); there is no CPU fallback.
"""
from .__about__ import __version__
from . import cpt, cvt, odt
from .main import get_new_points, optimize, optimize_points_cells
from .mesh import DeviceMesh, MeshTri, normalize_method_name
from .surfaces import Sphere

__all__ = [
    "__version__", "optimize_points_cells", "optimize", "get_new_points", "DeviceMesh",
    "MeshTri", "Sphere", "normalize_method_name", "cpt", "cvt", "odt",
]
