"""Flat-import shim with the reference's module name (reference src/regular_nd_grid_interpolation.py)."""
import _bootstrap  # noqa: F401
from space_time_pde_b200.regular_nd_grid_interpolation import (clip_tensor, regular_nd_grid_interpolation,  # noqa: F401
                                                                regular_nd_grid_interpolation_coefficients)
