"""Run the REAL reference (float32 and float64) on the seeded cases of tests/golden/seeded.py and store its outputs.

    python tests/golden/make_golden_seeded.py        (build container only: needs /root/reference)

These fixtures pin the parity gates of the shapes whose inputs are too large to commit: every test tolerance above 1e-5
is max(1e-5, 2 * err(reference fp32, reference fp64)) measured here, as in the small RB2 goldens."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import seeded  # noqa: E402

REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "src"))
warnings.filterwarnings("ignore")
from implicit_net import ImNet  # noqa: E402
from local_implicit_grid import query_local_implicit_grid  # noqa: E402
from nonlinearities import NONLINEARITIES  # noqa: E402
from pde import PDELayer, torch_diff  # noqa: E402

_cwd = os.getcwd()
os.chdir(os.path.join(REF, "experiments", "rb2d"))
sys.path.insert(0, os.getcwd())
from physics import get_rb2_pde_layer  # noqa: E402

os.chdir(_cwd)


def make_layer(name):
    eq = seeded.equations(name)
    if eq[0] == "rb2":
        return get_rb2_pde_layer(**eq[1])
    layer = PDELayer(in_vars=eq[0], out_vars=eq[1])
    for eq_name, string in eq[2]:
        layer.add_equation(string, eq_name)
    return layer


def run(name, dtype):
    k = seeded.CASES[name]
    Ws, bs, grid, q = seeded.build(name)
    model = ImNet(dim=k["dim"], in_features=k["c"], out_features=k["o"], nf=k["nf"], activation=NONLINEARITIES[k["act"]])
    with torch.no_grad():
        for i in range(6):
            getattr(model, f"fc{i}").weight.copy_(Ws[i])
            getattr(model, f"fc{i}").bias.copy_(bs[i])
    model = model.to(dtype)
    grid_t, q_t = grid.to(dtype), q.to(dtype)
    fwd = lambda pts: query_local_implicit_grid(model, grid_t, pts, 0., 1.)
    layer = make_layer(name)
    layer.update_forward_method(fwd)
    y, res = layer(q_t, return_residue=True)
    out = {"y": y.detach().numpy()}
    for key, v in res.items():
        out["res_" + key] = v.detach().numpy()
    # first partials and the diagonal second partials of every output (what the equation sets use)
    d, o = k["dim"], k["o"]
    inputs = [q_t[..., i:i + 1].clone().requires_grad_(True) for i in range(d)]
    yy = fwd(torch.cat(inputs, dim=-1))
    g1 = np.zeros(tuple(y.shape) + (d,))
    g2 = np.zeros(tuple(y.shape) + (d,))
    for i in range(o):
        for a in range(d):
            ga = torch_diff(yy[..., i:i + 1], inputs[a])
            g1[..., i, a] = ga.detach().numpy()[..., 0]
            gaa = torch_diff(ga, inputs[a])
            g2[..., i, a] = 0.0 if gaa is None else gaa.detach().numpy()[..., 0]
    out["g1"], out["g2diag"] = g1, g2
    return out


def main():
    only = sys.argv[1:]
    for name in seeded.CASES:
        if only and name not in only:
            continue
        arrays = {"checksum": np.float64(seeded.checksum(*seeded.build(name)))}
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            for key, v in run(name, dtype).items():
                arrays[f"{key}_{tag}"] = v
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **arrays)
        noise = {key[:-4]: float(np.max(np.abs(arrays[key] - arrays[key[:-4] + "_f64"])) /
                                 max(np.max(np.abs(arrays[key[:-4] + "_f64"])), 1e-300))
                 for key in arrays if key.endswith("_f32")}
        print(name, f"{os.path.getsize(path) / 1024:.1f} KiB", "ref fp32 vs fp64 rel-Linf:",
              {k_: f"{v:.1e}" for k_, v in noise.items()}, flush=True)


if __name__ == "__main__":
    main()
