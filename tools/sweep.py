"""BASELINE config[4] (index 4 = 'throughput sweep'): latent 32x32x32x128, ImNet nf=32, RB2 residuals,
p in {1e4 .. 1.6e7}; prints one JSON line per point count (device-resident inputs, CUDA events)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import space_time_pde_b200 as sp
from space_time_pde_b200 import _lib, jets

precision = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 32
jets.set_default_precision(precision)
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = sp.ImNet(dim=3, in_features=128, out_features=4, nf=nf, activation=sp.NONLINEARITIES["softplus"]).to(dev)
grid = torch.randn(1, 32, 32, 32, 128, device=dev) * 0.5
layer = sp.get_rb2_pde_layer(**bench.RB2)
layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
fpt = bench.flops_per_point(nf, 3, 128, 4, 6)
peak = bench.measured_peaks()["tflops"]
for p in (10_000, 100_000, 1_000_000, 4_000_000, 16_000_000):
    q = torch.rand(1, p, 3, device=dev) * (1 - 2e-6) + 1e-6
    for _ in range(3):
        with torch.no_grad():
            layer(q)
    torch.cuda.synchronize()
    reps = 5 if p <= 1_000_000 else 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        with torch.no_grad():
            layer(q)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rate = p / (ms * 1e-3)
    print(json.dumps({"points": p, "nf": nf, "precision": precision, "ms": ms, "points_per_s": rate,
                      "roofline_frac_bf16_sustained": rate * fpt / 1e12 / peak}), flush=True)
    del q
