"""gpview_b200 -- B200-native replacement of GPView's hybrid two-level voxelizer hot path.

The product is the C-ABI shared library ``libgpview_b200.so`` (include/gpview_b200.h; sources in gpview_b200/csrc/,
sm_100a only).  This Python package is only the thin ctypes harness the tests and bench.py drive it through, plus the
``torch.distributed`` plumbing of the z-slab sharded path (gpview_b200/sharded.py).  There is no CPU fallback: importing
works without a GPU (so that the symbol table can be checked), every compute call fails loudly without one.
"""
from .binding import (  # noqa: F401
    GPV_GATHER, GPV_KEEP_LISTS, GPV_NO_LEVEL2, GPV_NORMALS, GPV_PROFILE, GPV_PROFILE_L2, GPV_PACKED_L2, GPV_COLLISION, GPV_SAVE_COMPUTED_ONLY, LIB_PATH, Context, GpvError, Mesh, Params, Result, build, grid_for, lib,
    load_mesh, mesh_from_triangles,
)

__all__ = ["Context", "Mesh", "Params", "Result", "GpvError", "build", "lib", "load_mesh", "mesh_from_triangles", "grid_for",
           "GPV_NORMALS", "GPV_NO_LEVEL2", "GPV_KEEP_LISTS", "GPV_PROFILE", "GPV_PROFILE_L2", "GPV_PACKED_L2", "GPV_COLLISION", "GPV_GATHER", "GPV_SAVE_COMPUTED_ONLY", "LIB_PATH"]
