"""Torch/CPU port of the reference's hot path, autograd and all (TEST / BASELINE INFRASTRUCTURE ONLY).

Purpose: the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` need the reference's
*algorithm and cost structure* on the host cores of the GPU box, where /root/reference does not
exist.  This module restates that algorithm with the same torch operations:

  lookup + gather + weights          reference src/regular_nd_grid_interpolation.py:14-78
  rows -> MLP -> weighted corner sum reference src/local_implicit_grid.py:47-61
  skip-MLP                           reference src/implicit_net.py:40-54
  residuals: ONE torch.autograd.grad(create_graph=True) per textual dif()   reference src/pde.py:8-9,139-142

It is validated against golden vectors of the real reference (tests/test_oracle_golden.py) and its
wall-clock was checked against the real reference in the build container ("CPU baseline").
The product never imports it.
"""
from __future__ import annotations

import itertools
from typing import Dict, Sequence

import sympy
import torch
from sympy.core.function import AppliedUndef
from sympy.parsing.sympy_parser import parse_expr

ACTS = {"tanh": torch.nn.Tanh, "relu": torch.nn.ReLU, "softplus": torch.nn.Softplus, "elu": torch.nn.ELU,
        "leakyrelu": torch.nn.LeakyReLU}


class _Swish(torch.nn.Module):
    def __init__(self, beta=1.0):
        super().__init__()
        self.beta = torch.nn.Parameter(torch.tensor(float(beta)))

    def forward(self, x):
        return x * torch.sigmoid(self.beta * x)


class SkipMLP(torch.nn.Module):
    """ImNet-shaped decoder built from explicit weight arrays."""

    def __init__(self, Ws: Sequence, bs: Sequence, act: str, act_param: float = 1.0, dtype=torch.float32):
        super().__init__()
        self.act = _Swish(act_param) if act == "swish" else ACTS[act]()
        self.layers = torch.nn.ModuleList()
        for W, b in zip(Ws, bs):
            W = torch.as_tensor(W, dtype=dtype)
            lin = torch.nn.Linear(W.shape[1], W.shape[0])
            lin.weight.data = W.clone()
            lin.bias.data = torch.as_tensor(b, dtype=dtype).clone()
            self.layers.append(lin)
        self.to(dtype)

    def forward(self, rows):
        h = rows
        n = len(self.layers)
        for i in range(n - 2):
            h = torch.cat([self.act(self.layers[i](h)), rows], dim=-1)
        h = self.act(self.layers[n - 2](h))
        return self.layers[n - 1](h)


def lookup(grid, pts, xmin, xmax):
    """corner values, weights, relative coordinates (differentiable w.r.t. pts)."""
    dim = grid.dim() - 2
    dev = grid.device
    n = torch.tensor(grid.shape[1:-1]).float().to(dev)
    if isinstance(xmin, (int, float)) or isinstance(xmax, (int, float)):
        xmin = float(xmin) * torch.ones(dim, device=dev)
        xmax = float(xmax) * torch.ones(dim, device=dev)
    else:
        xmin, xmax = torch.as_tensor(xmin).to(dev), torch.as_tensor(xmax).to(dev)
    margin = 1e-6 * (xmax - xmin)
    pts = torch.max(torch.min(pts, xmax - margin), xmin + margin)
    h = (xmax - xmin) / (n - 1)
    lower = torch.floor(pts / h).long()
    corners = torch.tensor(list(itertools.product((0, 1), repeat=dim)), device=dev)   # [2^d, d]
    idx = lower.unsqueeze(2) + corners                                            # [b,p,2^d,d]
    batch_idx = torch.arange(grid.shape[0], device=dev).view(-1, 1, 1).expand(idx.shape[:-1])
    values = grid[(batch_idx,) + tuple(idx[..., k] for k in range(dim))]
    node_lo = lower.float() * h
    node_hi = (lower.float() + 1) * h
    sel = corners.bool()
    here = torch.where(sel, node_hi.unsqueeze(2), node_lo.unsqueeze(2))
    across = torch.where(sel, node_lo.unsqueeze(2), node_hi.unsqueeze(2))
    weights = torch.prod(torch.abs(pts.unsqueeze(2) - across) / h, dim=-1)
    rel = (pts.unsqueeze(2) - here) / h
    return values, weights, rel


def decode(model, grid, pts, xmin, xmax):
    values, weights, rel = lookup(grid, pts, xmin, xmax)
    rows = torch.cat([rel, values], dim=-1)
    shape = rows.shape
    out = model(rows.reshape(-1, shape[-1])).reshape(shape[0], shape[1], shape[2], -1)
    return torch.sum(out * weights.unsqueeze(-1), dim=-2)


def _dif(y, x):
    return torch.autograd.grad(y, x, grad_outputs=torch.ones_like(y), create_graph=True, allow_unused=True)[0]


def _evaluate(expr, env: Dict[str, torch.Tensor]):
    """Walk the expression; every dif node costs one autograd sweep (no memoisation, as lambdify)."""
    if expr.is_Symbol:
        return env[expr.name]
    if expr.is_Number:
        return float(expr)
    if expr.is_Add:
        acc = _evaluate(expr.args[0], env)
        for a in expr.args[1:]:
            acc = acc + _evaluate(a, env)
        return acc
    if expr.is_Mul:
        acc = _evaluate(expr.args[0], env)
        for a in expr.args[1:]:
            acc = acc * _evaluate(a, env)
        return acc
    if expr.is_Pow:
        return _evaluate(expr.args[0], env) ** float(expr.args[1])
    if isinstance(expr, AppliedUndef) and expr.func.__name__ == "dif":
        return _dif(_evaluate(expr.args[0], env), _evaluate(expr.args[1], env))
    raise NotImplementedError(str(expr.func))


def compile_equations(equations):
    """{name: (string, subs)} -> {name: sympy expression} (setup time, as PDELayer.add_equation)."""
    out = {}
    for name, (eqn, subs) in equations.items():
        expr = parse_expr(eqn)
        if subs:
            for k, v in subs.items():
                expr = expr.subs(k, v)
        out[name] = expr
    return out


def values_and_residuals(model, grid, pts, xmin, xmax, in_vars, out_vars, exprs):
    """(y, {name: residual}) exactly as PDELayer.__call__(x, return_residue=True)."""
    cols = [pts[..., k:k + 1].detach().clone().requires_grad_(True) for k in range(pts.shape[-1])]
    y = decode(model, grid, torch.cat(cols, dim=-1), xmin, xmax)
    env = {n: c for n, c in zip(in_vars, cols)}
    env.update({n: y[..., i:i + 1] for i, n in enumerate(out_vars)})
    return y, {name: _evaluate(expr, env) for name, expr in exprs.items()}
