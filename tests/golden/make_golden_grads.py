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


def main():
    for name, layer_fn in CASES.items():
        z = np.load(os.path.join(HERE, name + ".npz"))
        act, dim, nf = str(z["act"]), int(z["dim"]), int(z["nf"])
        grid = torch.tensor(z["grid"], dtype=torch.float64, requires_grad=True)
        q = torch.tensor(z["q"], dtype=torch.float64)
        c, o = grid.shape[-1], z["W5"].shape[0]
        model = mg.ImNet(dim=dim, in_features=c, out_features=o, nf=nf, activation=mg.NONLINEARITIES[act]).double()
        with torch.no_grad():
            for i in range(6):
                getattr(model, f"fc{i}").weight.copy_(torch.tensor(z[f"W{i}"], dtype=torch.float64))
                getattr(model, f"fc{i}").bias.copy_(torch.tensor(z[f"b{i}"], dtype=torch.float64))
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
        mg.save("grads_" + name, **arrays)


if __name__ == "__main__":
    main()
