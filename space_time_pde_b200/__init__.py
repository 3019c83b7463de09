"""space_time_pde_b200: B200-native decode + PDE-residual hot path of MeshfreeFlowNet.

Public surface = the reference's (maxjiang93/space_time_pde) Python call surface for this path:
``query_local_implicit_grid``, ``regular_nd_grid_interpolation[_coefficients]``, ``ImNet``,
``NONLINEARITIES``, ``PDELayer``, ``torch_diff``, ``get_rb2_pde_layer``.
"""
from .implicit_net import ImNet
from .local_implicit_grid import query_local_implicit_grid
from .nonlinearities import NONLINEARITIES, Swish
from .pde import PDELayer, torch_diff
from .physics import get_rb2_pde_layer
from .regular_nd_grid_interpolation import (clip_tensor, regular_nd_grid_interpolation,
                                            regular_nd_grid_interpolation_coefficients)
from .jets import deferred_checks, fused_query, set_default_precision
from .equations import JetSpec

__all__ = ["ImNet", "query_local_implicit_grid", "NONLINEARITIES", "Swish", "PDELayer", "torch_diff",
           "get_rb2_pde_layer", "clip_tensor", "regular_nd_grid_interpolation",
           "regular_nd_grid_interpolation_coefficients", "fused_query", "set_default_precision", "deferred_checks", "JetSpec"]
