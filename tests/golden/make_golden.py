"""Generate golden input/output vectors by running the REAL reference (read-only import).

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so the vectors are committed as small ``.npz``
fixtures next to this script.  Every case stores the inputs, the ImNet weights and the
outputs of the unmodified reference path

    PDELayer.__call__ -> query_local_implicit_grid -> regular_nd_grid_interpolation_coefficients
                      -> ImNet.forward -> torch.autograd.grad per dif()

run in float32 (``*_f32``) and float64 (``*_f64``).  Individual partial derivatives are
obtained with the reference's own ``pde.torch_diff``.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "src"))
sys.path.insert(0, os.path.join(REF, "experiments", "rb2d"))

import warnings

warnings.filterwarnings("ignore")

import regular_nd_grid_interpolation as rgi  # noqa: E402
from implicit_net import ImNet  # noqa: E402
from local_implicit_grid import query_local_implicit_grid  # noqa: E402
from nonlinearities import NONLINEARITIES  # noqa: E402
from pde import PDELayer, torch_diff  # noqa: E402

_cwd = os.getcwd()
os.chdir(os.path.join(REF, "experiments", "rb2d"))
from physics import get_rb2_pde_layer  # noqa: E402

os.chdir(_cwd)

HERE = os.path.dirname(os.path.abspath(__file__))


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def weights_of(model):
    out = {}
    for i in range(6):
        out[f"W{i}"] = getattr(model, f"fc{i}").weight.detach().numpy().astype(np.float32)
        out[f"b{i}"] = getattr(model, f"fc{i}").bias.detach().numpy().astype(np.float32)
    if isinstance(getattr(model.activ, "beta", None), torch.nn.Parameter):  # Swish only
        out["act_param"] = np.float32(model.activ.beta.item())
    return out


def run_reference(model, grid, q, xmin, xmax, pde_layer, dtype, hessian=True):
    """Values, residuals and (optionally) all first/second partials from the reference path."""
    model = model.double() if dtype == torch.float64 else model.float()
    grid_t = grid.to(dtype)
    q_t = q.to(dtype).clone()
    if torch.is_tensor(xmax):
        xmin_t, xmax_t = xmin.to(torch.float32), xmax.to(torch.float32)
    else:
        xmin_t, xmax_t = xmin, xmax
    fwd = lambda pts: query_local_implicit_grid(model, grid_t, pts, xmin_t, xmax_t)
    pde_layer.update_forward_method(fwd)
    y, res = pde_layer(q_t, return_residue=True)
    out = {"y": y.detach().numpy()}
    for k, v in res.items():
        out["res_" + k] = v.detach().numpy()
    if hessian:
        d, o = q.shape[-1], y.shape[-1]
        inputs = [q_t[..., i:i + 1].clone().requires_grad_(True) for i in range(d)]
        yy = fwd(torch.cat(inputs, dim=-1))
        g1 = np.zeros(tuple(y.shape) + (d,))
        g2 = np.zeros(tuple(y.shape) + (d, d))
        for i in range(o):
            for a in range(d):
                ga = torch_diff(yy[..., i:i + 1], inputs[a])
                g1[..., i, a] = ga.detach().numpy()[..., 0]
                for b in range(d):
                    gab = torch_diff(ga, inputs[b])
                    g2[..., i, a, b] = 0.0 if gab is None else gab.detach().numpy()[..., 0]
        out["g1"], out["g2"] = g1, g2
    model.float()
    return out


def case(name, *, dim, grid_shape, c, o, nf, act, p, batch, xmax, layer_fn, seed, q_fn=None,
         hessian=True, noncontig=False):
    torch.manual_seed(seed)
    model = ImNet(dim=dim, in_features=c, out_features=o, nf=nf, activation=NONLINEARITIES[act])
    if act == "swish":
        with torch.no_grad():
            model.activ.beta.fill_(1.3)
    grid = torch.randn(batch, *grid_shape, c) * 0.7
    xmax_arr = np.asarray(xmax, dtype=np.float32) if not np.isscalar(xmax) else None
    if q_fn is None:
        scale = torch.tensor(xmax_arr) if xmax_arr is not None else float(xmax)
        q = torch.rand(batch, p, dim) * scale
    else:
        q = q_fn(batch, p, dim)
    if xmax_arr is None:
        xmin_t, xmax_t = 0.0, float(xmax)
    else:
        xmin_t, xmax_t = torch.zeros(dim), torch.tensor(xmax_arr)
    arrays = dict(grid=grid.numpy(), q=q.numpy(), act=np.array(act), dim=dim, nf=nf,
                  xmax=(xmax_arr if xmax_arr is not None else np.float32(xmax)), **weights_of(model))
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        out = run_reference(model, grid, q, xmin_t, xmax_t, layer_fn(), dtype, hessian)
        for k, v in out.items():
            arrays[f"{k}_{tag}"] = v
    save(name, **arrays)


def rb2_layer(**kw):
    return lambda: get_rb2_pde_layer(**kw)


def diffusion_layer():
    layer = PDELayer(in_vars="x, y, t", out_vars="u, v")
    layer.add_equation("dif(u, t) - (dif(dif(u, x), x) + dif(dif(u, y), y))", "diffusion_u")
    layer.add_equation("dif(v, t) - (dif(dif(v, x), x) + dif(dif(v, y), y))", "diffusion_v")
    return layer


def ns3d_layer():
    """Config 4 style: steady 3-D incompressible Navier-Stokes with Laplacians (custom strings)."""
    layer = PDELayer(in_vars="x, y, z", out_vars="u, v, w, p")
    lap = lambda f: f"(dif(dif({f},x),x)+dif(dif({f},y),y)+dif(dif({f},z),z))"
    adv = lambda f: f"(u*dif({f},x)+v*dif({f},y)+w*dif({f},z))"
    layer.add_equation(f"{adv('u')}+dif(p,x)-0.01*{lap('u')}", "mom_u")
    layer.add_equation(f"{adv('v')}+dif(p,y)-0.01*{lap('v')}", "mom_v")
    layer.add_equation(f"{adv('w')}+dif(p,z)-0.01*{lap('w')}", "mom_w")
    layer.add_equation("dif(u,x)+dif(v,y)+dif(w,z)", "continuity")
    return layer


def generic_layer(dim, o):
    names_in = ["x", "y", "z", "s"][:dim]
    names_out = ["u", "v", "w", "r"][:o]

    def make():
        layer = PDELayer(in_vars=", ".join(names_in), out_vars=", ".join(names_out))
        # mixed second derivative + product inside dif + explicit coordinate dependence
        a, b = names_in[0], names_in[-1]
        layer.add_equation(f"dif(dif({names_out[0]},{a}),{b}) + {a}*dif({names_out[-1]}*{names_out[0]},{b})", "mixed")
        return layer

    return make


def eval_grid_points(n):
    """Tie points of train.py:136-139 / evaluation.py:229-232: linspace(eps, 1-eps) hits the clip bounds."""
    def fn(batch, p, dim):
        eps = 1e-6
        seqs = [torch.linspace(eps, 1 - eps, n) for _ in range(dim)]
        pts = torch.stack(torch.meshgrid(*seqs, indexing="ij"), dim=-1).reshape(-1, dim)
        return pts[None].expand(batch, -1, -1).contiguous()
    return fn


def main():
    # --- interpolation known-answer grids (reference rgi_test.py:12-40), seeded ---
    torch.manual_seed(0)
    arrays = {}
    for d in (1, 2, 3):
        axes = torch.meshgrid(*([torch.arange(11)] * d), indexing="ij")
        grid = torch.stack(axes, dim=-1).unsqueeze(0).float()
        pts = torch.rand(1, 100, d)
        out = rgi.regular_nd_grid_interpolation(grid, pts, 0., 1.)
        cv, w, xr = rgi.regular_nd_grid_interpolation_coefficients(grid, pts, 0., 1.)
        arrays.update({f"grid{d}": grid.numpy(), f"pts{d}": pts.numpy(), f"out{d}": out.numpy(),
                       f"cv{d}": cv.numpy(), f"w{d}": w.numpy(), f"xr{d}": xr.numpy()})
    save("interp_identity", **arrays)

    # --- decode + RB2 residuals, every activation ---
    for i, act in enumerate(("tanh", "relu", "softplus", "elu", "swish", "leakyrelu")):
        case(f"rb2_{act}", dim=3, grid_shape=(3, 5, 4), c=8, o=4, nf=4, act=act, p=48, batch=2, xmax=1.0,
             layer_fn=rb2_layer(t_crop=2., z_crop=1., x_crop=1., use_continuity=True), seed=100 + i)
    # paper-like shape (latent 4x16x16x32), normalised equations, softplus
    case("rb2_paper_softplus", dim=3, grid_shape=(4, 16, 16), c=32, o=4, nf=16, act="softplus", p=96, batch=1,
         xmax=1.0, layer_fn=rb2_layer(mean=[0.1, -0.2, 0.05, 0.3], std=[1.1, 0.9, 1.3, 0.7], t_crop=2.,
                                      z_crop=1., x_crop=2., use_continuity=True), seed=7, hessian=False)
    # non-unit domain (evaluation.py:224-235) with tensor bounds
    case("rb2_nonunit_tanh", dim=3, grid_shape=(4, 6, 5), c=8, o=4, nf=4, act="tanh", p=64, batch=1,
         xmax=[0.75, 1.0, 4.0], layer_fn=rb2_layer(use_continuity=False), seed=11)
    # eval-grid tie points (quirk Q2)
    case("rb2_ties_softplus", dim=3, grid_shape=(3, 4, 4), c=8, o=4, nf=4, act="softplus", p=0, batch=2,
         xmax=1.0, layer_fn=rb2_layer(use_continuity=True), seed=13, q_fn=eval_grid_points(4))
    # diffusion equations of the integration test (x, y, t -> u, v)
    case("diffusion_leakyrelu", dim=3, grid_shape=(4, 4, 4), c=8, o=2, nf=4, act="leakyrelu", p=64, batch=2,
         xmax=1.0, layer_fn=diffusion_layer, seed=17)
    case("ns3d_swish", dim=3, grid_shape=(4, 4, 4), c=8, o=4, nf=4, act="swish", p=64, batch=1,
         xmax=1.0, layer_fn=ns3d_layer, seed=19)
    # other dimensionalities (local_implicit_grid_test.py:16-20 covers d=3 and d=4)
    for d, o in ((1, 2), (2, 3), (4, 3)):
        case(f"generic_d{d}_softplus", dim=d, grid_shape=(4,) * d, c=6, o=o, nf=4, act="softplus", p=40,
             batch=2, xmax=1.0, layer_fn=generic_layer(d, o), seed=23 + d)


if __name__ == "__main__":
    main()
