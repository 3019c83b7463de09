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
from .implicit_net import ddp_wrapper, decoder_signature
from .jets import JetSpec, active_request, bounds_tensors, fused_query


class _DdpGradAverage(torch.autograd.Function):
    """Identity on the decoder parameters whose backward averages their gradients over the DDP process group."""

    @staticmethod
    def forward(ctx, ddp, *params):
        ctx.ddp = ddp
        return tuple(p.view_as(p) for p in params)

    @staticmethod
    def backward(ctx, *grads):
        from .jets import _ddp_average

        live = [g.contiguous().clone() for g in grads if g is not None]
        _ddp_average(ctx.ddp, live)
        it = iter(live)
        return (None,) + tuple(next(it) if g is not None else None for g in grads)


def query_local_implicit_grid(model, latent_grid, query_pts, xmin, xmax):
    """Query the latent grid at ``query_pts`` [b,p,d]; returns [b,p,o] (see reference docstring)."""
    sig = decoder_signature(model)
    if sig is not None and latent_grid.dtype == torch.float32 and query_pts.dtype == torch.float32:
        layers, act, act_param = sig
        ddp = ddp_wrapper(model)
        request = active_request()
        if request is not None and query_pts.shape[-1] == latent_grid.dim() - 2:
            # derivatives w.r.t. the coordinates come from the jets, not from autograd through q
            y, jets = fused_query(latent_grid, query_pts.detach(), xmin, xmax, layers, act, act_param,
                                  spec=request.spec, ddp=ddp)
            request.records.append((y, jets, query_pts))
            return y
        if torch.is_grad_enabled() and query_pts.requires_grad:
            # The caller differentiates through the query points itself (a PDELayer forward method that post-processes
            # this output, third derivatives, ...): like the reference, the result must carry a graph that can be
            # differentiated AGAIN w.r.t. the points, which a custom backward cannot offer - so this case runs the
            # same formulation as plain differentiable torch ops (not the hot path; DDP is bypassed here as well,
            # hence the explicit gradient averaging hook).
            from ._torch_jets import query_jets

            lo, hi = bounds_tensors(xmin, xmax, query_pts.shape[-1], query_pts.device)
            dev = query_pts.device
            params = [l.weight for l in layers] + [l.bias for l in layers] + ([act_param] if act_param is not None else [])
            if ddp is not None:
                params = _DdpGradAverage.apply(ddp, *params)
            n = len(layers)
            y, _ = query_jets(latent_grid, query_pts, lo.to(dev), hi.to(dev), params[:n], params[n:2 * n], act,
                              params[2 * n] if act_param is not None else None, JetSpec())
            return y
        y, _ = fused_query(latent_grid, query_pts, xmin, xmax, layers, act, act_param, spec=JetSpec(), ddp=ddp)
        return y

    # generic decoder: same algorithm as the reference on top of the lookup kernels / torch ops
    corner_values, weights, x_relative = rgi.regular_nd_grid_interpolation_coefficients(
        latent_grid, query_pts, xmin, xmax)
    rows = torch.cat([x_relative.to(corner_values.dtype), corner_values], dim=-1)
    b, p, j, _ = rows.shape
    decoded = model(rows.reshape(b * p * j, -1)).reshape(b, p, j, -1)
    return torch.sum(decoded * weights.unsqueeze(-1).to(decoded.dtype), dim=-2)
