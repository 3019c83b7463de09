"""Quick RB2 parity check of the loaded library against the fp64 oracle (all three precisions, two decoder widths)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import space_time_pde_b200 as sp
from oracle import jet_oracle as jo
from space_time_pde_b200 import jets

dev = torch.device("cuda:0")
kw = dict(t_crop=2., z_crop=1., x_crop=1., use_continuity=True)
for nf, c, gshape, npts in ((8, 16, (4, 6, 5), 3000), (32, 32, (4, 16, 16), 5000), (128, 32, (4, 16, 16), 2048)):
    for act in ("softplus", "tanh"):
        torch.manual_seed(nf)
        model = sp.ImNet(dim=3, in_features=c, out_features=4, nf=nf, activation=sp.NONLINEARITIES[act]).to(dev)
        grid = (torch.randn(1, *gshape, c) * 0.5).to(dev)
        q = torch.rand(1, npts, 3, device=dev)
        layer = sp.get_rb2_pde_layer(**kw)
        layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
        Ws = [l.weight.detach().cpu().numpy() for l in model.fc]
        bs = [l.bias.detach().cpu().numpy() for l in model.fc]
        n = 256
        yj = jo.query_jet(grid.cpu().numpy(), q[:, :n].cpu().numpy(), 0., 1., Ws, bs, act)
        iv, ov, eqs = jo.rb2_equations(**kw)
        ref = jo.pde_residuals(yj, q[:, :n].cpu().numpy(), iv, ov, eqs)
        for prec in ("fp16x3", "fp16"):
            jets.set_default_precision(prec)
            with torch.no_grad():
                y, res = layer(q)
            torch.cuda.synchronize()
            rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
            errs = [rel(y[:, :n].cpu().numpy(), yj.v)] + [rel(res[k][:, :n].cpu().numpy(), ref[k]) for k in res]
            print(f"nf={nf} {act} {prec}: max rel-Linf {max(errs):.2e}", "OK" if max(errs) < (1e-5 if prec == "fp16x3" else 3e-2) else "FAIL", flush=True)
