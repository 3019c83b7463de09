"""Flat-import shim with the reference's module name (reference experiments/rb2d/physics.py)."""
import _bootstrap  # noqa: F401
from space_time_pde_b200.physics import get_rb2_pde_layer  # noqa: F401
