"""Golden GRADIENTS from the real reference (read-only import), for the fused reverse sweep.

    python tests/golden/make_golden_grads.py

For a few of the committed forward cases (same inputs and weights, read back from their ``.npz``) the unmodified
reference path is run in float64 with a training-style loss

    loss = |y|.mean() + 0.0125 * |stack(residuals)|.mean()            (experiments/rb2d/train.py:70-75 with zero targets)

and ``loss.backward()`` - i.e. autograd THROUGH every ``torch.autograd.grad(create_graph=True)`` of src/pde.py:8 -
gives the gradients w.r.t. the latent grid and every ImNet weight / bias.  Stored as ``grads_<case>.npz``.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import make_golden as mg  # noqa: E402  (sets up the reference imports)
from tests.helpers import RB2_CASES  # noqa: E402

CASES = {
    "rb2_tanh": lambda: mg.get_rb2_pde_layer(**RB2_CASES["rb2_tanh"]),
    "rb2_softplus": lambda: mg.get_rb2_pde_layer(**RB2_CASES["rb2_softplus"]),
    "rb2_elu": lambda: mg.get_rb2_pde_layer(**RB2_CASES["rb2_elu"]),
    "rb2_swish": lambda: mg.get_rb2_pde_layer(**RB2_CASES["rb2_swish"]),        # learnable beta: g_beta is stored too
    "rb2_paper_softplus": lambda: mg.get_rb2_pde_layer(**RB2_CASES["rb2_paper_softplus"]),
    "rb2_nonunit_tanh": lambda: mg.get_rb2_pde_layer(**RB2_CASES["rb2_nonunit_tanh"]),
    "generic_d1_softplus": mg.generic_layer(1, 2),
    "generic_d2_softplus": mg.generic_layer(2, 3),
    "generic_d4_softplus": mg.generic_layer(4, 3),
}


def run(z, layer_fn, dtype):
    act, dim, nf = str(z["act"]), int(z["dim"]), int(z["nf"])
    grid = torch.tensor(z["grid"], dtype=dtype, requires_grad=True)
    q = torch.tensor(z["q"], dtype=dtype)
    c, o = grid.shape[-1], z["W5"].shape[0]
    model = mg.ImNet(dim=dim, in_features=c, out_features=o, nf=nf, activation=mg.NONLINEARITIES[act]).to(dtype)
    with torch.no_grad():
        for i in range(6):
            getattr(model, f"fc{i}").weight.copy_(torch.tensor(z[f"W{i}"], dtype=dtype))
            getattr(model, f"fc{i}").bias.copy_(torch.tensor(z[f"b{i}"], dtype=dtype))
        if act == "swish":
            model.activ.beta.fill_(float(z["act_param"]))
    xmax = z["xmax"]
    if xmax.ndim == 0:
        xmin_t, xmax_t = 0.0, float(xmax)
    else:
        xmin_t, xmax_t = torch.zeros(dim), torch.tensor(xmax.astype(np.float32))
    layer = layer_fn()
    layer.update_forward_method(lambda pts: mg.query_local_implicit_grid(model, grid, pts, xmin_t, xmax_t))
    y, res = layer(q, return_residue=True)
    loss = y.abs().mean() + 0.0125 * torch.stack(list(res.values())).abs().mean()
    loss.backward()
    arrays = {"loss": np.float64(loss.item()), "g_grid": grid.grad.numpy()}
    for i in range(6):
        arrays[f"g_W{i}"] = getattr(model, f"fc{i}").weight.grad.numpy()
        arrays[f"g_b{i}"] = getattr(model, f"fc{i}").bias.grad.numpy()
    if act == "swish":
        arrays["g_beta"] = model.activ.beta.grad.numpy().reshape(1)
    return arrays


def main():
    """float64 gradients (``g_*``: the golden values) and, for the parity gates, the rel-L-infinity distance of the
    reference's own float32 run from them (``noise_*``: one number per tensor - the fp32 arrays are not stored)."""
    for name, layer_fn in CASES.items():
        z = np.load(os.path.join(HERE, name + ".npz"))
        arrays = run(z, layer_fn, torch.float64)
        f32 = run(z, layer_fn, torch.float32)
        noise = {}
        for k, v in list(arrays.items()):
            if k.startswith("g_"):
                den = np.max(np.abs(v))
                noise[k] = float(np.max(np.abs(f32[k].astype(np.float64) - v)) / (den if den > 0 else 1.0))
                arrays["noise_" + k[2:]] = np.float64(noise[k])
        mg.save("grads_" + name, **arrays)
        print("   reference fp32 vs fp64 gradient rel-Linf: max %.1e" % max(noise.values()),
              {k: f"{v:.0e}" for k, v in noise.items()})


if __name__ == "__main__":
    main()
