"""Regular n-d grid lookup / multilinear interpolation on the GPU.

Drop-in for reference src/regular_nd_grid_interpolation.py:14-104 (same names, argument meaning,
return shapes and quirks).  CUDA float32 tensors that do not need gradients go through the
``stpde_interp*`` kernels (include/stpde.h); if a gradient is required the same arithmetic is
expressed with differentiable torch ops (the fused decode path never takes this route - it has
its own kernels, see local_implicit_grid.py).
"""
import ctypes

import torch

from . import _lib
from ._torch_jets import cell_geometry, corner_bits
from .jets import _i64, bounds_tensors


def clip_tensor(input_tensor, xmin, xmax):
    """Per-column clip; ties propagate half of the gradient to each side (torch.max/min semantics)."""
    return torch.max(torch.min(input_tensor, xmax), xmin)


def _needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t.requires_grad for t in tensors)


def _check_status(status):
    if int(status.item()) & 1:
        raise IndexError("index out of range for the latent grid (reference rgi.py:52 ignores xmin)")


def _kernel_args(grid, query_pts, xmin, xmax):
    if not (grid.is_cuda and query_pts.is_cuda):
        raise RuntimeError("stpde grid interpolation needs CUDA tensors (no CPU fallback)")
    dim = grid.dim() - 2
    lo, hi = bounds_tensors(xmin, xmax, dim, grid.device)
    size = (ctypes.c_int32 * dim)(*[int(s) for s in grid.shape[1:-1]])
    f32 = lambda t: (ctypes.c_float * dim)(*[float(v) for v in t])
    grid32 = grid.detach().float()
    q32 = query_pts.detach().float()
    return dim, size, f32(lo), f32(hi), grid32, q32


def _coefficients_torch(grid, query_pts, xmin, xmax):
    dim = grid.dim() - 2
    lo, hi = bounds_tensors(xmin, xmax, dim, grid.device)
    lo, hi = lo.to(grid.device), hi.to(grid.device)
    b, p, _ = query_pts.shape
    qd = query_pts.detach()
    qc, clipgrad, ind0, xyz0, xyz1, cs = cell_geometry(grid.shape[1:-1], qd, lo, hi)
    qc = clip_tensor(query_pts, (lo + 1e-6 * (hi - lo)).to(query_pts.dtype), (hi - 1e-6 * (hi - lo)).to(query_pts.dtype))
    bits = corner_bits(dim, grid.device)
    idx = ind0[:, :, None, :] + bits[None, None]
    ib = torch.arange(b, device=grid.device)[:, None, None].expand(b, p, bits.shape[0])
    corner_values = grid[(ib,) + tuple(idx[..., k] for k in range(dim))]
    sel = bits[None, None].bool()
    pos = torch.where(sel, xyz1[:, :, None, :], xyz0[:, :, None, :])
    opp = torch.where(sel, xyz0[:, :, None, :], xyz1[:, :, None, :])
    weights = torch.prod(torch.abs(qc.unsqueeze(-2) - opp) / cs, dim=-1)
    x_relative = (qc.unsqueeze(-2) - pos) / cs
    return corner_values, weights, x_relative


def regular_nd_grid_interpolation_coefficients(grid, query_pts, xmin=0., xmax=1.):
    """Corner values [b,p,2^d,c], weights [b,p,2^d], relative coordinates [b,p,2^d,d].

    Reference: src/regular_nd_grid_interpolation.py:14-78.
    """
    if _needs_grad(grid, query_pts) or grid.dtype != torch.float32 or query_pts.dtype != torch.float32:
        if not grid.is_cuda:
            raise RuntimeError("stpde grid interpolation needs CUDA tensors (no CPU fallback)")
        return _coefficients_torch(grid, query_pts, xmin, xmax)
    dim, size, lo, hi, g32, q32 = _kernel_args(grid, query_pts, xmin, xmax)
    b, p, _ = q32.shape
    c = g32.shape[-1]
    J = 1 << dim
    cv = torch.empty(b, p, J, c, dtype=torch.float32, device=grid.device)
    w = torch.empty(b, p, J, dtype=torch.float32, device=grid.device)
    xr = torch.empty(b, p, J, dim, dtype=torch.float32, device=grid.device)
    status = torch.zeros(1, dtype=torch.int32, device=grid.device)
    lib = _lib.load()
    with torch.cuda.device(grid.device):
        rc = lib.stpde_interp_coefficients(b, p, dim, size, c, g32.data_ptr(), _i64(g32.stride()), q32.data_ptr(),
                                           _i64(q32.stride()), lo, hi, cv.data_ptr(), w.data_ptr(), xr.data_ptr(),
                                           status.data_ptr(), torch.cuda.current_stream(grid.device).cuda_stream)
    _lib.check(rc)
    _check_status(status)
    return cv, w, xr


def regular_nd_grid_interpolation(grid, query_pts, xmin=0., xmax=1.):
    """Multilinear interpolation, [b,p,c].  Reference: src/regular_nd_grid_interpolation.py:81-104."""
    if _needs_grad(grid, query_pts) or grid.dtype != torch.float32 or query_pts.dtype != torch.float32:
        cv, w, _ = regular_nd_grid_interpolation_coefficients(grid, query_pts, xmin, xmax)
        return torch.sum(cv * w.unsqueeze(-1), dim=-2)
    dim, size, lo, hi, g32, q32 = _kernel_args(grid, query_pts, xmin, xmax)
    b, p, _ = q32.shape
    c = g32.shape[-1]
    out = torch.empty(b, p, c, dtype=torch.float32, device=grid.device)
    status = torch.zeros(1, dtype=torch.int32, device=grid.device)
    lib = _lib.load()
    with torch.cuda.device(grid.device):
        rc = lib.stpde_interp(b, p, dim, size, c, g32.data_ptr(), _i64(g32.stride()), q32.data_ptr(),
                              _i64(q32.stride()), lo, hi, out.data_ptr(), status.data_ptr(),
                              torch.cuda.current_stream(grid.device).cuda_stream)
    _lib.check(rc)
    _check_status(status)
    return out
