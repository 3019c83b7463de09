"""Forward-mode jets written with differentiable torch ops.

NOT the hot path.  Two uses only:
  * ``FusedJetQuery.backward`` re-evaluates the jets with these ops under autograd to obtain the
    reverse-mode gradients w.r.t. the latent grid / decoder weights (a fused CUDA backward is
    SURVEY.md 8(f) rank 1, scheduled after the forward path);
  * CPU tests of the host logic (equation compiler, residual programs) use it as a stand-in jet
    provider through ``space_time_pde_b200.jets.set_test_backend``.
It mirrors the kernel's formulation (csrc/simt_kernels.cu): derivatives inside the MLP are taken
w.r.t. the cell-local coordinate and rescaled by clipgrad/cubesize in the blend.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from .equations import JetSpec


def act_jet(kind: str, z: torch.Tensor, beta=None):
    if kind == "tanh":
        t = torch.tanh(z)
        u = 1 - t * t
        return t, u, -2 * t * u
    if kind == "relu":
        m = (z > 0).to(z.dtype)
        return z * m, m, torch.zeros_like(z)
    if kind == "leakyrelu":
        m = torch.where(z > 0, torch.ones_like(z), torch.full_like(z, 0.01))
        return z * m, m, torch.zeros_like(z)
    if kind == "softplus":
        lin = z > 20
        s = torch.sigmoid(z)
        return F.softplus(z), torch.where(lin, torch.ones_like(z), s), torch.where(lin, torch.zeros_like(z), s * (1 - s))
    if kind == "elu":
        neg = z <= 0
        e = torch.exp(torch.where(neg, z, torch.zeros_like(z)))
        return torch.where(neg, e - 1, z), torch.where(neg, e, torch.ones_like(z)), torch.where(neg, e, torch.zeros_like(z))
    if kind == "swish":
        bz = beta * z
        s = torch.sigmoid(bz)
        ds = s * (1 - s)
        return z * s, s + bz * ds, beta * ds * (2 + bz * (1 - 2 * s))
    raise ValueError(kind)


def cell_geometry(grid_shape: Sequence[int], q: torch.Tensor, xmin: torch.Tensor, xmax: torch.Tensor):
    """Clip / cell lookup in float32 like the kernel (reference rgi.py:47-52,69-70)."""
    dim = q.shape[-1]
    size = torch.tensor(list(grid_shape), dtype=torch.float32, device=q.device)
    xmin = xmin.to(torch.float32)
    xmax = xmax.to(torch.float32)
    eps = 1e-6 * (xmax - xmin)
    lo, hi = xmin + eps, xmax - eps
    cs = (xmax - xmin) / (size - 1)
    qmin = torch.minimum(q, hi)
    qc = torch.maximum(qmin, lo)
    one, half, zero = torch.ones_like(q), torch.full_like(q, 0.5), torch.zeros_like(q)
    gmin = torch.where(q < hi, one, torch.where(q == hi, half, zero))
    gmax = torch.where(qmin > lo, one, torch.where(qmin == lo, half, zero))
    ind0 = torch.floor(qc / cs).long()
    xyz0 = ind0.float() * cs
    xyz1 = (ind0.float() + 1) * cs
    return qc, gmin * gmax, ind0, xyz0.to(q.dtype), xyz1.to(q.dtype), cs.to(q.dtype)


def corner_bits(dim: int, device) -> torch.Tensor:
    j = torch.arange(1 << dim, device=device)
    return torch.stack([(j >> (dim - 1 - k)) & 1 for k in range(dim)], dim=-1)   # [J, d], dim 0 = MSB


def query_jets(grid: torch.Tensor, q: torch.Tensor, xmin: torch.Tensor, xmax: torch.Tensor,
               Ws: Sequence[torch.Tensor], bs: Sequence[torch.Tensor], act: str, act_param,
               spec: JetSpec) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """y [b,p,o] and jets [n_jet,b,p,o]; differentiable w.r.t. grid, Ws, bs, act_param (and q)."""
    b, p, dim = q.shape
    qc, clipgrad, ind0, xyz0, xyz1, cs = cell_geometry(grid.shape[1:-1], q.detach(), xmin, xmax)
    if q.requires_grad:   # keep the (piecewise linear) dependence on q for autograd users
        qc = qc + (q - q.detach()) * clipgrad
    bits = corner_bits(dim, q.device)                                       # [J,d]
    J = bits.shape[0]
    idx = ind0[:, :, None, :] + bits[None, None]                            # [b,p,J,d]
    ib = torch.arange(b, device=q.device)[:, None, None].expand(b, p, J)
    latent = grid[(ib,) + tuple(idx[..., k] for k in range(dim))]           # [b,p,J,c]
    pos = torch.where(bits[None, None].bool(), xyz1[:, :, None, :], xyz0[:, :, None, :])
    opp = torch.where(bits[None, None].bool(), xyz0[:, :, None, :], xyz1[:, :, None, :])
    xrel = (qc[:, :, None, :] - pos) / cs                                   # [b,p,J,d]
    diff = qc[:, :, None, :] - opp
    fac = diff.abs() / cs                                                   # [b,p,J,d]
    dfac = torch.sign(diff.detach()) * (clipgrad / cs)[:, :, None, :]
    dxr = (clipgrad / cs)                                                   # [b,p,d]

    x = torch.cat([xrel, latent], dim=-1)
    n = len(Ws)
    first, second = spec.first, spec.second
    comp_of_dir = {k: 1 + i for i, k in enumerate(first)}
    # layer 0
    z = F.linear(x, Ws[0], bs[0])
    s0, s1, s2 = act_jet(act, z, act_param)
    h: List[torch.Tensor] = [s0]
    for k in first:
        h.append(s1 * Ws[0][:, k])
    for (a, c) in second:
        h.append(s2 * Ws[0][:, a] * Ws[0][:, c])
    for l in range(1, n - 1):
        kh = Ws[l - 1].shape[0]
        Wh = Ws[l][:, :kh]
        z = F.linear(torch.cat([h[0], x], dim=-1), Ws[l], bs[l])
        zt = [z]
        for i, k in enumerate(first):
            zt.append(F.linear(h[1 + i], Wh) + Ws[l][:, kh + k])
        for i in range(len(second)):
            zt.append(F.linear(h[1 + len(first) + i], Wh))
        s0, s1, s2 = act_jet(act, z, act_param)
        h = [s0] + [s1 * zt[1 + i] for i in range(len(first))]
        for i, (a, c) in enumerate(second):
            h.append(s2 * zt[comp_of_dir[a]] * zt[comp_of_dir[c]] + s1 * zt[1 + len(first) + i])
    out = [F.linear(h[0], Ws[n - 1], bs[n - 1])] + [F.linear(hc, Ws[n - 1]) for hc in h[1:]]   # [b,p,J,o]

    def wprod(repl: dict):
        w = None
        for k in range(dim):
            f = repl.get(k, fac)[..., k]
            w = f if w is None else w * f
        return w[..., None]                                                  # [b,p,J,1]

    w = wprod({})
    y = (out[0] * w).sum(dim=2)
    planes = []
    for i, a in enumerate(first):
        planes.append((wprod({a: dfac}) * out[0] + w * dxr[:, :, None, a:a + 1] * out[1 + i]).sum(dim=2))
    for i, (a, c) in enumerate(second):
        oa, oc_ = out[comp_of_dir[a]], out[comp_of_dir[c]]
        da, dc = dxr[:, :, None, a:a + 1], dxr[:, :, None, c:c + 1]
        term = wprod({a: dfac}) * dc * oc_ + wprod({c: dfac}) * da * oa + w * da * dc * out[1 + len(first) + i]
        if a != c:
            term = term + wprod({a: dfac, c: dfac}) * out[0]
        planes.append(term.sum(dim=2))
    jets = torch.stack(planes, dim=0) if planes else None
    return y, jets
