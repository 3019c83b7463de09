"""Flat-import shim with the reference's module name (reference src/nonlinearities.py)."""
import _bootstrap  # noqa: F401
from space_time_pde_b200.nonlinearities import NONLINEARITIES, Swish  # noqa: F401
