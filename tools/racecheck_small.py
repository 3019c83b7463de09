"""Tiny forward + training step for compute-sanitizer (racecheck / memcheck): every kernel family once."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import space_time_pde_b200 as sp
from space_time_pde_b200 import jets

dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
jets.set_default_precision(prec)
torch.manual_seed(0)
for nf in (8, 32):
    model = sp.ImNet(dim=3, in_features=8, out_features=4, nf=nf, activation=sp.NONLINEARITIES["softplus"]).to(dev)
    grid = (torch.randn(1, 3, 4, 5, 8) * 0.5).to(dev).requires_grad_(True)
    q = torch.rand(1, 300, 3, device=dev)
    layer = sp.get_rb2_pde_layer(t_crop=2., z_crop=1., x_crop=1., use_continuity=True)
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
    with torch.no_grad():
        y, res = layer(q)                                   # inference: fused final layer
    y, sums, counts = layer.loss_sums(q, torch.zeros_like(y), "l1")   # training forward + fused loss
    (sums[0] / counts[0] + 0.0125 * sums[1] / counts[1]).backward()
    torch.cuda.synchronize()
    print("nf", nf, prec, "ok", float(sums[0]), float(grid.grad.abs().max()))
