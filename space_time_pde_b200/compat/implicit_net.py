"""Flat-import shim with the reference's module name (reference src/implicit_net.py)."""
import _bootstrap  # noqa: F401
from space_time_pde_b200.implicit_net import ImNet  # noqa: F401
