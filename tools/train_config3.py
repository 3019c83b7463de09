"""BASELINE config 3 (paper-like training shape) on ONE GPU's share: 8 crops x 131072 query points = 2^20 points,
ImNet nf=32 Softplus, normalised RB2 equations + continuity, L1 losses (alpha_pde = 0.0125), single-pass fp16 MLP
operands with fp32 accumulation and fp32 jets.  Prints points/s of forward+residuals and of the full training step
(chunks of the batch with the forward planes kept for the backward)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import space_time_pde_b200 as sp
from space_time_pde_b200 import jets

precision = sys.argv[1] if len(sys.argv) > 1 else "fp16"
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 16384          # points per crop and chunk
os.environ["STPDE_WORKSPACE_MB"] = os.environ.get("STPDE_WORKSPACE_MB", "40960")
jets.set_default_precision(precision)
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, P = 8, 131072
model = sp.ImNet(dim=3, in_features=32, out_features=4, nf=32, activation=sp.NONLINEARITIES["softplus"]).to(dev)
grid = (torch.randn(B, 4, 16, 16, 32) * 0.5).to(dev).requires_grad_(True)
q = torch.rand(B, P, 3, device=dev) * (1 - 2e-6) + 1e-6
target = torch.randn(B, P, 4, device=dev)
layer = sp.get_rb2_pde_layer(mean=[0.1, -0.2, 0.05, 0.3], std=[1.1, 0.9, 1.3, 0.7], t_crop=2., z_crop=1., x_crop=2.,
                             prandtl=1., rayleigh=1e6, use_continuity=True)
layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))


def forward_only():
    with torch.no_grad():
        return layer(q)


def train_step():
    model.zero_grad(set_to_none=True)
    grid.grad = None
    for s0 in range(0, P, chunk):
        y, res = layer(q[:, s0:s0 + chunk])
        reg = (y - target[:, s0:s0 + chunk]).abs().sum() / (B * P * 4)
        pde = torch.stack(list(res.values())).abs().sum() / (B * P * 4)
        (reg + 0.0125 * pde).backward()


def time_it(fn, n):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


fwd_ms = time_it(forward_only, 3)
trn_ms = time_it(train_step, 2)
from space_time_pde_b200 import _lib
lib = _lib.load()
lib.stpde_profile_enable(1)
_lib.profile_read()
train_step()
prof = {k: round(v[0], 1) for k, v in _lib.profile_read().items() if v[1] > 0}
lib.stpde_profile_enable(0)
print(json.dumps({"config": "BASELINE config 3 share of one GPU: 8 crops x 131072 points, ImNet nf=32, normalised RB2, "
                            f"precision {precision}, chunks of 8 x {chunk} points",
                  "forward_residuals_ms": fwd_ms, "forward_points_per_s": B * P / (fwd_ms * 1e-3),
                  "train_step_ms": trn_ms, "train_points_per_s": B * P / (trn_ms * 1e-3), "train_kernel_ms": prof}))
