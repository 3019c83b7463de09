"""Local implicit grid query.  Drop-in for reference src/local_implicit_grid.py:10-61.

When ``model`` is an ImNet-structured decoder (this package's ``ImNet``, the reference's own
class, or a DataParallel/DDP wrapper of either) the whole chain

    corner gather -> [x_rel, latent] rows -> MLP at the 2^d corners -> multilinear blend

runs in the fused CUDA path (``stpde_jet_forward``), and if a ``PDELayer`` is collecting
residuals the partial derivatives it needs are produced by the same launch (jets.JetRequest).
Any other ``nn.Module`` is evaluated on the gathered rows like the reference does.
"""
import torch

from . import regular_nd_grid_interpolation as rgi
from .implicit_net import decoder_signature
from .jets import JetSpec, active_request, fused_query


def query_local_implicit_grid(model, latent_grid, query_pts, xmin, xmax):
    """Query the latent grid at ``query_pts`` [b,p,d]; returns [b,p,o] (see reference docstring)."""
    sig = decoder_signature(model)
    if sig is not None and latent_grid.dtype == torch.float32 and query_pts.dtype == torch.float32:
        layers, act, act_param = sig
        request = active_request()
        if request is not None and query_pts.shape[-1] == latent_grid.dim() - 2:
            # derivatives w.r.t. the coordinates come from the jets, not from autograd through q
            y, jets = fused_query(latent_grid, query_pts.detach(), xmin, xmax, layers, act, act_param,
                                  spec=request.spec)
            request.records.append((y, jets, query_pts))
            return y
        y, _ = fused_query(latent_grid, query_pts, xmin, xmax, layers, act, act_param, spec=JetSpec())
        return y

    # generic decoder: same algorithm as the reference on top of the lookup kernels / torch ops
    corner_values, weights, x_relative = rgi.regular_nd_grid_interpolation_coefficients(
        latent_grid, query_pts, xmin, xmax)
    rows = torch.cat([x_relative.to(corner_values.dtype), corner_values], dim=-1)
    b, p, j, _ = rows.shape
    decoded = model(rows.reshape(b * p * j, -1)).reshape(b, p, j, -1)
    return torch.sum(decoded * weights.unsqueeze(-1).to(decoded.dtype), dim=-2)
