"""Host-side profile of the reference-size training step (10 crops x 1024 points, ImNet nf=32): where the
milliseconds go between Python, the C ABI and the GPU."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import space_time_pde_b200 as sp
from space_time_pde_b200 import _lib

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, P, nf = 10, 1024, int(os.environ.get("NF", "32"))
model = sp.ImNet(dim=3, in_features=32, out_features=4, nf=nf, activation=sp.NONLINEARITIES["softplus"]).to(dev)
grid = (torch.randn(B, 4, 16, 16, 32) * 0.5).to(dev).requires_grad_(True)
q = torch.rand(B, P, 3, device=dev)
target = torch.randn(B, P, 4, device=dev)
layer = sp.get_rb2_pde_layer(**bench.RB2)
layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))


def step():
    model.zero_grad(set_to_none=True)
    grid.grad = None
    y, res = layer(q, return_residue=True)
    loss = torch.nn.functional.l1_loss(y, target) + 0.0125 * torch.stack(list(res.values())).abs().mean()
    loss.backward()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
lib = _lib.load()
lib.stpde_profile_enable(1)
_lib.profile_read()
step()
prof = _lib.profile_read()
lib.stpde_profile_enable(0)
print("kernel ms:", {k: round(v[0], 3) for k, v in prof.items() if v[1] > 0}, "sum", round(sum(v[0] for v in prof.values()), 3),
      "launches", sum(v[1] for v in prof.values()))
n = 50
t0 = time.perf_counter()
for _ in range(n):
    step()
torch.cuda.synchronize()
print("ms/step:", (time.perf_counter() - t0) / n * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
