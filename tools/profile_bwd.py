"""One training step (forward + residuals + L1 losses + fused reverse sweep) at BASELINE config[1] shapes with fewer
points, for ncu captures of the reverse-mode kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import space_time_pde_b200 as sp
from space_time_pde_b200 import jets

precision = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
jets.set_default_precision(precision)
dev = torch.device("cuda:0")
nf = int(os.environ.get("NF", bench.NF))
torch.manual_seed(0)
model = sp.ImNet(dim=3, in_features=bench.CHANNELS, out_features=4, nf=nf, activation=sp.NONLINEARITIES[bench.ACT]).to(dev)
grid, q = bench.synthetic_inputs(1234, dev, npts)
grid.requires_grad_(True)
layer = sp.get_rb2_pde_layer(**bench.RB2)
layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
from space_time_pde_b200 import _lib
lib = _lib.load()
for it in range(steps):
    if it == steps - 1 and os.environ.get("STPDE_PRINT_PROFILE"):
        torch.cuda.synchronize()
        lib.stpde_profile_enable(1)
        _lib.profile_read()
    model.zero_grad()
    grid.grad = None
    y, res = layer(q)
    loss = y.abs().mean() + 0.0125 * torch.stack(list(res.values())).abs().mean()
    loss.backward()
torch.cuda.synchronize()
if os.environ.get("STPDE_PRINT_PROFILE"):
    print({k: round(v[0], 2) for k, v in _lib.profile_read().items() if v[1] > 0})
print("ok", float(loss), float(grid.grad.abs().max()))
