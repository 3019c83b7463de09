"""CPU oracle for the decode + PDE-residual hot path (TEST INFRASTRUCTURE ONLY).

This file restates, in plain numpy, the algorithm of the reference's hot path:

  * cell lookup / corner gather / weights   -> reference src/regular_nd_grid_interpolation.py:14-78
  * plain multilinear interpolation         -> reference src/regular_nd_grid_interpolation.py:81-104
  * local implicit grid query (MLP x 2^d)   -> reference src/local_implicit_grid.py:47-61
  * ImNet skip-MLP                          -> reference src/implicit_net.py:40-54
  * activations                             -> reference src/nonlinearities.py:5-22
  * PDE residuals from equation strings     -> reference src/pde.py:115-143

The reference obtains derivatives with one ``torch.autograd.grad`` per ``dif(...)``
occurrence (src/pde.py:8-9).  The oracle computes the *same mathematical quantities*
with second-order forward-mode jets (value, gradient, full Hessian w.r.t. the query
coordinates) in float64, including the reference's quirks:

  Q1  ind0 = floor(q / cubesize) ignores xmin                     (rgi.py:52)
  Q2  clip = max(min(q, hi), lo): gradient 1 inside, 0 outside, 0.5 on exact ties (rgi.py:11)
      |x| has gradient sign(x) with sign(0) = 0                   (rgi.py:75, torch.abs)
  Q3  cubesize / clip bounds are formed in float32 when xmin/xmax are python scalars,
      lists or float32 tensors                                    (rgi.py:40-51)

Parity pinning: ``tests/test_oracle_golden.py`` checks this oracle against
  (i)  the reference's own known-answer tests (identity grids, rgi_test.py:12-40;
       heat equation residual -7 at (1,2,3), pde_test.py:12-53), and
  (ii) golden vectors produced by importing the *real* reference in the build container
       (``tests/golden/make_golden.py``; fp32 and fp64 runs, all six activations).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may
import this module.  The product path (``space_time_pde_b200``) never does.
"""
from __future__ import annotations

import itertools
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

ACTIVATIONS = ("tanh", "relu", "softplus", "elu", "swish", "leakyrelu")


# ----------------------------------------------------------------------------------------------
# activations: sigma, sigma', sigma''  (torch semantics, reference src/nonlinearities.py:15-22)
# ----------------------------------------------------------------------------------------------
def _sigmoid(x):
    out = np.empty_like(x)
    pos = x >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-x[pos]))
    ex = np.exp(x[~pos])
    out[~pos] = ex / (1.0 + ex)
    return out


def activation_jet(kind: str, z: np.ndarray, param: float = 1.0):
    """Return (sigma(z), sigma'(z), sigma''(z)) with torch's autograd conventions."""
    if kind == "tanh":
        t = np.tanh(z)
        return t, 1.0 - t * t, -2.0 * t * (1.0 - t * t)
    if kind == "relu":
        m = (z > 0).astype(z.dtype)
        return z * m, m, np.zeros_like(z)
    if kind == "leakyrelu":
        m = np.where(z > 0, 1.0, 0.01).astype(z.dtype)
        return z * m, m, np.zeros_like(z)
    if kind == "softplus":
        # torch.nn.Softplus(beta=1, threshold=20): linear branch above the threshold
        lin = z > 20.0
        zc = np.where(lin, 0.0, z)
        s = _sigmoid(zc)
        v = np.where(lin, z, np.log1p(np.exp(zc)))
        return v, np.where(lin, 1.0, s), np.where(lin, 0.0, s * (1.0 - s))
    if kind == "elu":
        neg = z <= 0
        e = np.exp(np.where(neg, z, 0.0))
        return np.where(neg, e - 1.0, z), np.where(neg, e, 1.0), np.where(neg, e, 0.0)
    if kind == "swish":
        b = param
        s = _sigmoid(b * z)
        ds = s * (1.0 - s)
        return z * s, s + b * z * ds, b * ds * (2.0 + b * z * (1.0 - 2.0 * s))
    raise ValueError(f"unknown activation {kind!r}")


# ----------------------------------------------------------------------------------------------
# second-order jets over the d query coordinates
# ----------------------------------------------------------------------------------------------
class Jet:
    """value v[...], gradient g[k][...], Hessian h[k][l][...] (symmetric), k,l < d."""

    __slots__ = ("v", "g", "h", "d")

    def __init__(self, v, g=None, h=None, d=None):
        self.v = v
        self.d = d if d is not None else (len(g) if g is not None else 0)
        zero = lambda: np.zeros_like(v)
        self.g = g if g is not None else [zero() for _ in range(self.d)]
        self.h = h if h is not None else [[zero() for _ in range(self.d)] for _ in range(self.d)]

    @staticmethod
    def const(v, d):
        return Jet(np.asarray(v), d=d)

    def _lift(self, o):
        return o if isinstance(o, Jet) else Jet(np.broadcast_to(np.asarray(o, dtype=self.v.dtype), self.v.shape).copy(), d=self.d)

    def __add__(self, o):
        o = self._lift(o)
        d = self.d
        return Jet(self.v + o.v, [self.g[k] + o.g[k] for k in range(d)],
                   [[self.h[k][l] + o.h[k][l] for l in range(d)] for k in range(d)])

    __radd__ = __add__

    def __neg__(self):
        d = self.d
        return Jet(-self.v, [-self.g[k] for k in range(d)], [[-self.h[k][l] for l in range(d)] for k in range(d)])

    def __sub__(self, o):
        return self + (-self._lift(o))

    def __rsub__(self, o):
        return self._lift(o) - self

    def __mul__(self, o):
        o = self._lift(o)
        d = self.d
        g = [self.g[k] * o.v + self.v * o.g[k] for k in range(d)]
        h = [[self.h[k][l] * o.v + self.g[k] * o.g[l] + self.g[l] * o.g[k] + self.v * o.h[k][l]
              for l in range(d)] for k in range(d)]
        return Jet(self.v * o.v, g, h)

    __rmul__ = __mul__

    def apply(self, f0, f1, f2):
        """Elementwise y = f(self) given f, f', f'' evaluated at self.v."""
        d = self.d
        g = [f1 * self.g[k] for k in range(d)]
        h = [[f2 * self.g[k] * self.g[l] + f1 * self.h[k][l] for l in range(d)] for k in range(d)]
        return Jet(f0, g, h)

    def powi(self, n: int):
        if n == 0:
            return Jet(np.ones_like(self.v), d=self.d)
        if n < 0:
            f0 = self.v ** n
            return self.apply(f0, n * self.v ** (n - 1), n * (n - 1) * self.v ** (n - 2))
        f2 = n * (n - 1) * self.v ** (n - 2) if n >= 2 else np.zeros_like(self.v)
        return self.apply(self.v ** n, n * self.v ** (n - 1), f2)

    def linear(self, W, b=None):
        """Last-axis affine map with a constant matrix W[out, in]."""
        d = self.d
        f = lambda a: a @ W.T
        v = f(self.v) + (b if b is not None else 0.0)
        return Jet(v, [f(self.g[k]) for k in range(d)], [[f(self.h[k][l]) for l in range(d)] for k in range(d)])

    def dif(self, k: int):
        """Jet of d(self)/dq_k: value = g[k], gradient = h[k][:]; third order is unavailable."""
        nan = np.full_like(self.v, np.nan)
        return Jet(self.g[k], [self.h[k][l] for l in range(self.d)],
                   [[nan for _ in range(self.d)] for _ in range(self.d)])


def jet_cat(jets: Sequence[Jet]) -> Jet:
    d = jets[0].d
    cat = lambda xs: np.concatenate(xs, axis=-1)
    return Jet(cat([j.v for j in jets]), [cat([j.g[k] for j in jets]) for k in range(d)],
               [[cat([j.h[k][l] for j in jets]) for l in range(d)] for k in range(d)])


# ----------------------------------------------------------------------------------------------
# grid lookup (reference rgi.py:14-78)
# ----------------------------------------------------------------------------------------------
def _bounds(xmin, xmax, dim, bounds_dtype):
    """xmin/xmax conversion of rgi.py:39-45 (python scalars / sequences -> float32)."""
    if isinstance(xmin, (int, float)) or isinstance(xmax, (int, float)):
        lo = np.full([dim], float(xmin), dtype=np.float32)
        hi = np.full([dim], float(xmax), dtype=np.float32)
    else:
        lo = np.asarray(xmin)
        hi = np.asarray(xmax)
        if lo.dtype != np.float64 or bounds_dtype == np.float32:
            lo = lo.astype(np.float32)
            hi = hi.astype(np.float32)
    return lo, hi


def corner_bits(dim: int) -> np.ndarray:
    """{0,1}^dim in meshgrid('ij') order: dimension 0 is the most significant bit (rgi.py:56-57)."""
    return np.array(list(itertools.product((0, 1), repeat=dim)), dtype=np.int64)


def cell_lookup(grid_shape: Sequence[int], q: np.ndarray, xmin, xmax, dtype=np.float64):
    """Clip, cell index, local coordinates.  Returns a dict of per-point quantities.

    ``dtype`` is the arithmetic type of the query points (float64 = exact oracle,
    float32 = bit-faithful restatement of the reference's fp32 path).
    """
    dim = q.shape[-1]
    size = np.asarray(grid_shape, dtype=np.float32)          # rgi.py:37 (.float())
    lo32, hi32 = _bounds(xmin, xmax, dim, np.float32)
    eps = (np.float32(1e-6) * (hi32 - lo32)).astype(np.float32)   # rgi.py:48
    lo = (lo32 + eps).astype(np.float32)
    hi = (hi32 - eps).astype(np.float32)
    cubesize = ((hi32 - lo32) / (size - np.float32(1.0))).astype(np.float32)  # rgi.py:51
    q = q.astype(dtype)
    lo_t, hi_t, cs = lo.astype(dtype), hi.astype(dtype), cubesize.astype(dtype)
    qmin = np.minimum(q, hi_t)
    qc = np.maximum(qmin, lo_t)                               # rgi.py:11
    # torch.max/min backward: 1 to the selected operand, 0.5 each on exact ties
    gmin = np.where(q < hi_t, 1.0, np.where(q == hi_t, 0.5, 0.0))
    gmax = np.where(qmin > lo_t, 1.0, np.where(qmin == lo_t, 0.5, 0.0))
    clipgrad = (gmin * gmax).astype(dtype)
    ind0 = np.floor(qc / cs).astype(np.int64)                 # rgi.py:52 (no "- xmin": quirk Q1)
    # rgi.py:69-70: ind0.float() * cubesize is a float32 product whatever the dtype of the points
    xyz0 = (ind0.astype(np.float32) * cubesize).astype(dtype)
    xyz1 = ((ind0.astype(np.float32) + np.float32(1.0)) * cubesize).astype(dtype)
    return dict(qc=qc, clipgrad=clipgrad, ind0=ind0, xyz0=xyz0, xyz1=xyz1, cubesize=cs)


def _gather_corners(grid: np.ndarray, ind0: np.ndarray, bits: np.ndarray) -> np.ndarray:
    """grid[b, ind...] with python negative-index wrap-around / IndexError (rgi.py:59-66)."""
    b, p, dim = ind0.shape
    idx = ind0[:, :, None, :] + bits[None, None, :, :]                     # [b,p,2^d,d]
    ib = np.arange(b)[:, None, None]
    return grid[(np.broadcast_to(ib, idx.shape[:-1]),) + tuple(idx[..., k] for k in range(dim))]


def interp_coefficients(grid: np.ndarray, q: np.ndarray, xmin=0.0, xmax=1.0, dtype=np.float64):
    """reference regular_nd_grid_interpolation_coefficients (rgi.py:14-78)."""
    dim = q.shape[-1]
    cl = cell_lookup(grid.shape[1:-1], q, xmin, xmax, dtype)
    bits = corner_bits(dim)
    corner_values = _gather_corners(grid, cl["ind0"], bits)
    xyz01 = np.stack([cl["xyz0"], cl["xyz1"]], axis=0)                      # [2,b,p,d]
    k = np.arange(dim)
    pos = np.stack([xyz01[bits[j], :, :, k] for j in range(bits.shape[0])], axis=0)      # [2^d,d,b,p]
    pos_ = np.stack([xyz01[1 - bits[j], :, :, k] for j in range(bits.shape[0])], axis=0)
    pos = pos.transpose(2, 3, 0, 1)
    pos_ = pos_.transpose(2, 3, 0, 1)
    qc = cl["qc"][:, :, None, :]
    cs = cl["cubesize"]
    dxyz = np.abs(qc - pos_) / cs
    weights = np.prod(dxyz, axis=-1)
    x_rel = (qc - pos) / cs
    return corner_values, weights, x_rel


def interp(grid: np.ndarray, q: np.ndarray, xmin=0.0, xmax=1.0, dtype=np.float64):
    """reference regular_nd_grid_interpolation (rgi.py:81-104)."""
    cv, w, _ = interp_coefficients(grid, q, xmin, xmax, dtype)
    return np.sum(cv.astype(dtype) * w[..., None], axis=-2)


# ----------------------------------------------------------------------------------------------
# ImNet (reference implicit_net.py:40-54)
# ----------------------------------------------------------------------------------------------
def imnet_forward(x: np.ndarray, Ws: Sequence[np.ndarray], bs: Sequence[np.ndarray], act: str,
                  act_param: float = 1.0) -> np.ndarray:
    h = x
    n = len(Ws)
    for i in range(n - 2):
        h = activation_jet(act, h @ Ws[i].T + bs[i], act_param)[0]
        h = np.concatenate([h, x], axis=-1)
    h = activation_jet(act, h @ Ws[n - 2].T + bs[n - 2], act_param)[0]
    return h @ Ws[n - 1].T + bs[n - 1]


def imnet_forward_jet(x: Jet, Ws, bs, act: str, act_param: float = 1.0) -> Jet:
    h = x
    n = len(Ws)
    for i in range(n - 1):
        z = h.linear(Ws[i], bs[i])
        h = z.apply(*activation_jet(act, z.v, act_param))
        if i < n - 2:
            h = jet_cat([h, x])
    return h.linear(Ws[n - 1], bs[n - 1])


# ----------------------------------------------------------------------------------------------
# local implicit grid query with jets (reference local_implicit_grid.py:47-61 + pde.py:131-136)
# ----------------------------------------------------------------------------------------------
def query_jet(grid: np.ndarray, q: np.ndarray, xmin, xmax, Ws, bs, act: str, act_param: float = 1.0,
              dtype=np.float64) -> Jet:
    """y[b,p,o] with gradient/Hessian w.r.t. the (unclipped) query coordinates q[b,p,:]."""
    b, p, dim = q.shape
    cl = cell_lookup(grid.shape[1:-1], q, xmin, xmax, dtype)
    bits = corner_bits(dim)
    corner_values = _gather_corners(grid, cl["ind0"], bits).astype(dtype)   # [b,p,2^d,c]
    cs = cl["cubesize"]
    Ws = [np.asarray(W, dtype=dtype) for W in Ws]
    bs = [np.asarray(v, dtype=dtype) for v in bs]

    # clipped coordinate as a jet in q: value qc, d qc_k / d q_k = clipgrad_k, no curvature
    def coord_jet(k):
        g = [np.zeros((b, p), dtype=dtype) for _ in range(dim)]
        g[k] = cl["clipgrad"][..., k].astype(dtype)
        return Jet(cl["qc"][..., k], g)

    qj = [coord_jet(k) for k in range(dim)]
    total = None
    for j in range(bits.shape[0]):
        xr, w = [], None
        for k in range(dim):
            pos = cl["xyz1"][..., k] if bits[j, k] else cl["xyz0"][..., k]
            opp = cl["xyz0"][..., k] if bits[j, k] else cl["xyz1"][..., k]
            xr.append((qj[k] - pos) * (1.0 / cs[k]))
            diff = qj[k] - opp
            sgn = np.sign(diff.v)                                             # torch.abs backward
            fac = diff.apply(np.abs(diff.v), sgn, np.zeros_like(sgn)) * (1.0 / cs[k])
            w = fac if w is None else w * fac                                 # torch.prod order
        xj = jet_cat([Jet(a.v[..., None], [g[..., None] for g in a.g],
                          [[hh[..., None] for hh in row] for row in a.h]) for a in xr]
                     + [Jet.const(corner_values[:, :, j, :], dim)])
        out = imnet_forward_jet(xj, Ws, bs, act, act_param)                   # [b,p,o]
        wj = Jet(w.v[..., None], [g[..., None] for g in w.g], [[hh[..., None] for hh in row] for row in w.h])
        term = out * wj
        total = term if total is None else total + term
    return total


def query(grid, q, xmin, xmax, Ws, bs, act, act_param=1.0, dtype=np.float64) -> np.ndarray:
    """Values only (reference query_local_implicit_grid)."""
    cv, w, xr = interp_coefficients(grid, q, xmin, xmax, dtype)
    x = np.concatenate([xr, cv.astype(dtype)], axis=-1)
    Ws = [np.asarray(W, dtype=dtype) for W in Ws]
    bs = [np.asarray(v, dtype=dtype) for v in bs]
    out = imnet_forward(x.reshape(-1, x.shape[-1]), Ws, bs, act, act_param)
    out = out.reshape(x.shape[0], x.shape[1], x.shape[2], -1)
    return np.sum(out * w[..., None], axis=-2)


# ----------------------------------------------------------------------------------------------
# equation strings -> residuals on jets (reference pde.py:36-86,115-143)
# ----------------------------------------------------------------------------------------------
def parse_equation(eqn_str: str, in_vars: Sequence[str], out_vars: Sequence[str], subs_dict=None):
    import sympy
    from sympy.parsing.sympy_parser import parse_expr

    expr = parse_expr(eqn_str)
    if subs_dict:
        for key, val in subs_dict.items():
            expr = expr.subs(key, val)
    allowed = {sympy.Symbol(s) for s in list(in_vars) + list(out_vars)}
    if not expr.free_symbols <= allowed:
        raise ValueError(f"variables {expr.free_symbols} do not match {allowed}")
    return expr


def eval_expr_jet(expr, env: Dict[str, Jet], in_vars: Sequence[str], d: int) -> Jet:
    """Evaluate a sympy expression on jets; dif(Y, a) = jet derivative along input a."""
    import sympy

    def ev(e) -> Jet:
        if e.is_Symbol:
            return env[e.name]
        if e.is_Number:
            any_j = next(iter(env.values()))
            return Jet(np.full_like(any_j.v, float(e)), d=d)
        if e.is_Add:
            acc = ev(e.args[0])
            for a in e.args[1:]:
                acc = acc + ev(a)
            return acc
        if e.is_Mul:
            acc = ev(e.args[0])
            for a in e.args[1:]:
                acc = acc * ev(a)
            return acc
        if e.is_Pow:
            base, ex = e.args
            if ex.is_Integer:
                return ev(base).powi(int(ex))
            raise NotImplementedError(f"non-integer power {e}")
        if isinstance(e, sympy.core.function.AppliedUndef) and e.func.__name__ == "dif":
            y, x = e.args
            return ev(y).dif(list(in_vars).index(x.name))
        raise NotImplementedError(f"unsupported expression node {e.func}")

    return ev(expr)


def pde_residuals(yjet: Jet, q: np.ndarray, in_vars: Sequence[str], out_vars: Sequence[str],
                  equations: Dict[str, Tuple[str, Optional[dict]]]) -> Dict[str, np.ndarray]:
    """Residuals of ``{name: (eqn_str, subs_dict)}`` evaluated on the output jet yjet[b,p,o].

    Variables bind to tensor columns by *position* (quirk Q4; pde.py:131,137).
    """
    d = q.shape[-1]
    env: Dict[str, Jet] = {}
    for k, name in enumerate(in_vars):
        g = [np.zeros(q.shape[:-1] + (1,)) for _ in range(d)]
        g[k] = np.ones(q.shape[:-1] + (1,))
        env[name] = Jet(q[..., k:k + 1].astype(np.float64), g)
    for i, name in enumerate(out_vars):
        sl = lambda a: a[..., i:i + 1]
        env[name] = Jet(sl(yjet.v), [sl(g) for g in yjet.g], [[sl(h) for h in row] for row in yjet.h])
    out = {}
    for name, (eqn, subs) in equations.items():
        expr = parse_equation(eqn, in_vars, out_vars, subs)
        out[name] = eval_expr_jet(expr, env, in_vars, d).v
    return out


def rb2_equations(mean=None, std=None, t_crop=2., z_crop=1., x_crop=2., prandtl=1., rayleigh=1e6,
                  use_continuity=False):
    """Equation strings of the Rayleigh-Benard layer (reference experiments/rb2d/physics.py:18-54)."""
    P = (rayleigh * prandtl) ** (-1 / 2)
    R = (rayleigh / prandtl) ** (-1 / 2)
    nt, nz, nx = 1. / t_crop, 1. / z_crop, 1. / x_crop
    diffusion = lambda v: f"(({nx})**2*dif(dif({v},x),x)+({nz})**2*dif(dif({v},z),z))"
    advect = lambda v: f"(u*{nx}*dif({v},x)+w*{nz}*dif({v},z))"
    eqs = {
        "transport_eqn_b": f"{nt}*dif(b,t)-{P}*{diffusion('b')}+{advect('b')}",
        "transport_eqn_u": f"{nt}*dif(u,t)-{R}*{diffusion('u')}+dif(p,x)+{advect('u')}",
        "transport_eqn_w": f"{nt}*dif(w,t)-{R}*{diffusion('w')}+dif(p,z)-b+{advect('w')}",
    }
    if use_continuity:
        eqs["continuity"] = f"{nx} * dif(u, x) + {nz} * dif(w, z)"
    subs = None
    if (mean is not None) or (std is not None):
        if mean is None or std is None or len(mean) != 4 or len(std) != 4:
            raise ValueError("mean and std must both be arrays of len 4")
        subs = {v: f"{v}*{std[i]}+{mean[i]}" for i, v in enumerate(("p", "b", "u", "w"))}
    return ("t", "x", "z"), ("p", "b", "u", "w"), {k: (v, subs) for k, v in eqs.items()}
