"""One forward + backward of a config-3-shaped chunk (8 crops x P points, ImNet nf=32, fp16 both sweeps): per-slot kernel
times and whether the training forward kept its planes (no recompute).  usage: train_chunk_probe.py POINTS_PER_CROP [WS_MB]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
P = int(sys.argv[1])
if len(sys.argv) > 2:
    os.environ["STPDE_WORKSPACE_MB"] = sys.argv[2]
import torch
import space_time_pde_b200 as sp
from space_time_pde_b200 import _lib, jets

dev = torch.device("cuda:0")
jets.set_default_precision("fp16"); jets.set_backward_precision("fp16")
torch.manual_seed(3)
model = sp.ImNet(dim=3, in_features=32, out_features=4, nf=32, activation=sp.NONLINEARITIES["softplus"]).to(dev)
grid = (torch.randn(8, 4, 16, 16, 32) * 0.5).to(dev).requires_grad_(True)
q = torch.rand(8, P, 3, device=dev) * (1 - 2e-6) + 1e-6
layer = sp.get_rb2_pde_layer(mean=[0.1, -0.2, 0.05, 0.3], std=[1.1, 0.9, 1.3, 0.7], t_crop=2., z_crop=1., x_crop=2., use_continuity=True)
layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
lib = _lib.load()
def step():
    y, sums, _ = layer.loss_sums(q, None, "l1")
    (sums[0] * 1e-6 + sums[1] * 1e-8).backward()
for _ in range(2): step()
torch.cuda.synchronize()
lib.stpde_profile_enable(1); _lib.profile_read()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
prof = _lib.profile_read(); lib.stpde_profile_enable(0)
tot = 8 * P
print(f"points {tot} ms {e0.elapsed_time(e1):.2f} -> {tot / e0.elapsed_time(e1) * 1e3:.3g} points/s; launches {sum(v[1] for v in prof.values())}")
print({k: (round(v[0], 2), v[1]) for k, v in prof.items() if v[1] > 0})
