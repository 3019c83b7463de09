"""Forward values + jets and a fused backward of narrow decoders, saved to an .npz (argv[1]); run once with STPDE_PACK=0
and once with STPDE_PACK=1 by tests/test_gpu_configs.py::test_row_group_packing_is_bitwise_neutral."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import space_time_pde_b200 as sp
from space_time_pde_b200 import jets

dev = torch.device("cuda:0")
out = {}
for nf, npts in ((8, 777), (16, 1500), (32, 4099), (64, 2048)):
    for prec in ("fp16x3", "fp16"):
        jets.set_default_precision(prec)
        torch.manual_seed(nf)
        model = sp.ImNet(dim=3, in_features=16, out_features=4, nf=nf, activation=sp.NONLINEARITIES["softplus"]).to(dev)
        grid = (torch.randn(2, 4, 6, 5, 16) * 0.5).to(dev).requires_grad_(True)
        q = torch.rand(2, npts, 3, device=dev) * (1 - 2e-6) + 1e-6
        layer = sp.get_rb2_pde_layer(t_crop=2., z_crop=1., x_crop=1., use_continuity=True)
        layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
        with torch.no_grad():
            y, res = layer(q)                                  # inference route (fused final layer where it applies)
        out[f"y_{nf}_{prec}"] = y.cpu().numpy()
        for k, v in res.items():
            out[f"{k}_{nf}_{prec}"] = v.cpu().numpy()
        y, res = layer(q)                                      # training route: forward-save + fused reverse sweep
        (y.square().mean() + sum(v.square().mean() for v in res.values())).backward()
        out[f"ytrain_{nf}_{prec}"] = y.detach().cpu().numpy()
        # wgrad / vertex adjoints are sums of atomics (order varies run to run): compared with a tolerance by the test
        out[f"ggrid_{nf}_{prec}"] = grid.grad.cpu().numpy()
        out[f"gw_{nf}_{prec}"] = model.fc[-2].weight.grad.cpu().numpy()
        model.zero_grad()
np.savez(sys.argv[1], **out)
print("saved", len(out))
