"""Per-kernel CUDA-event breakdown of one forward+residual pass (args: precision nf channels grid_n npts)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import space_time_pde_b200 as sp
from space_time_pde_b200 import _lib, jets

precision, nf, c, gn, npts = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
jets.set_default_precision(precision)
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = sp.ImNet(dim=3, in_features=c, out_features=4, nf=nf, activation=sp.NONLINEARITIES["softplus"]).to(dev)
grid = torch.randn(1, gn, gn, gn, c, device=dev) * 0.5
q = torch.rand(1, npts, 3, device=dev) * (1 - 2e-6) + 1e-6
layer = sp.get_rb2_pde_layer(**bench.RB2)
layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
for _ in range(3):
    with torch.no_grad():
        layer(q)
torch.cuda.synchronize()
lib = _lib.load()
lib.stpde_profile_enable(1)
_lib.profile_read()
with torch.no_grad():
    layer(q)
prof = _lib.profile_read()
lib.stpde_profile_enable(0)
tot = sum(v[0] for v in prof.values())
print(precision, "nf", nf, "pts", npts, "total_ms %.2f" % tot, {k: round(v[0], 2) for k, v in prof.items() if v[1] > 0})
