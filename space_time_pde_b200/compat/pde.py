"""Flat-import shim with the reference's module name (reference src/pde.py)."""
import _bootstrap  # noqa: F401
from space_time_pde_b200.pde import PDELayer, torch_diff  # noqa: F401
