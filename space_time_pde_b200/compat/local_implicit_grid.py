"""Flat-import shim with the reference's module name (reference src/local_implicit_grid.py)."""
import _bootstrap  # noqa: F401
from space_time_pde_b200.local_implicit_grid import query_local_implicit_grid  # noqa: F401
