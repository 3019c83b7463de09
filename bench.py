#!/usr/bin/env python
"""Benchmark of the decode + PDE-residual hot path (BASELINE.json metric / config[1]).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision fp32|fp16x3|fp16]

One "step" = one pass of the hot path over one batch of synthetic query points: values + all
Rayleigh-Benard residuals (3 transport equations + continuity) for 2^20 points against a
4x16x16x32 latent grid with ImNet(nf=128, Softplus), float32.

  value : whole-job throughput with inputs resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through the public API with HOST (pinned) buffers; H2D of the query points
          and latent grid and D2H of values + residuals are inside the timed region
  roofline : dominant kernel (hidden layer 1 contraction) timed with CUDA events on its stream
  cpu_baseline : the reference's algorithm (oracle/ref_port.py: torch + one autograd.grad per dif)
                 on the host cores, bounded sample
Under torchrun each rank processes its own 2^20 points (weak scaling, no data-path collective).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "query-points/sec (fwd+PDE residuals)"
UNIT = "points/s"
NF, CHANNELS, GRID, NPTS, ACT = 128, 32, (4, 16, 16), 1 << 20, "softplus"
RB2 = dict(t_crop=2., z_crop=1., x_crop=1., prandtl=1., rayleigh=1e6, use_continuity=True)
KC = 6   # value + d/dt, d/dx, d/dz + d2/dx2, d2/dz2


def flops_per_point(nf, d, c, o, kc):
    """SURVEY.md 8(d) / BASELINE.md 4: F_pt = 2 * MAC_row * 2^d * K."""
    mac_row = 170 * nf * nf + 31 * nf * (d + c) + nf * o
    return 2 * mac_row * (1 << d) * kc


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(tflops=float(p["bf16_tflops_sustained"]), tflops_burst=float(p["bf16_tflops"]),
                    hbm=float(p["hbm_gbs"]), source="measured (MEASURED_PEAKS.json, bf16 sustained)")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                                  ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
def synthetic_inputs(seed, device, npts=NPTS):
    g = torch.Generator().manual_seed(seed)
    grid = torch.randn(1, *GRID, CHANNELS, generator=g) * 0.5
    q = torch.rand(1, npts, 3, generator=g) * (1 - 2e-6) + 1e-6
    return grid.to(device), q.to(device)


def make_model(device, seed=0):
    import space_time_pde_b200 as sp
    torch.manual_seed(seed)
    model = sp.ImNet(dim=3, in_features=CHANNELS, out_features=4, nf=NF, activation=sp.NONLINEARITIES[ACT])
    return model.to(device)


def cpu_reference_rate(sample_pts, repeats=1, warm_pts=128):
    """points/s of the reference algorithm (autograd per dif) on the host cores."""
    from oracle import jet_oracle as jo
    from oracle import ref_port as rp

    torch.set_num_threads(os.cpu_count() or 1)
    model = make_model("cpu")
    port = rp.SkipMLP([l.weight.detach().numpy() for l in model.fc], [l.bias.detach().numpy() for l in model.fc], ACT)
    iv, ov, eqs = jo.rb2_equations(**RB2)
    exprs = rp.compile_equations(eqs)
    grid, q = synthetic_inputs(1234, "cpu", max(sample_pts, warm_pts))
    rp.values_and_residuals(port, grid, q[:, :warm_pts], 0., 1., iv, ov, exprs)   # warm-up
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        rp.values_and_residuals(port, grid, q[:, :sample_pts], 0., 1., iv, ov, exprs)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sample_pts / best, best


def gpu_eager_rate(device, batch=2048, batches=3):
    """Reference algorithm (oracle/ref_port.py, autograd per dif) as PyTorch eager ops on the GPU, pseudo-batched
    like the reference's evaluation loop (experiments/rb2d/evaluation.py:54-69)."""
    from oracle import jet_oracle as jo
    from oracle import ref_port as rp

    torch.backends.cuda.matmul.allow_tf32 = False
    model = make_model("cpu")
    port = rp.SkipMLP([l.weight.detach().numpy() for l in model.fc], [l.bias.detach().numpy() for l in model.fc], ACT).to(device)
    iv, ov, eqs = jo.rb2_equations(**RB2)
    exprs = rp.compile_equations(eqs)
    grid, q = synthetic_inputs(1234, device, batch * (batches + 1))
    rp.values_and_residuals(port, grid, q[:, :batch], 0., 1., iv, ov, exprs)            # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(1, batches + 1):
        y, res = rp.values_and_residuals(port, grid, q[:, i * batch:(i + 1) * batch], 0., 1., iv, ov, exprs)
        del y, res
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"value": batch * batches / dt, "unit": UNIT,
            "sample": f"{batches} pseudo-batches of {batch} points, torch eager fp32 on the same GPU (oracle/ref_port.py)"}


def reference_size_train_step(device, with_eager):
    """Training step of the reference's own size (experiments/rb2d/train.py defaults: 10 crops x 1024 query points,
    ImNet nf=32, latent 4x16x16x32, RB2 + continuity, L1 losses): ms per step of the fused path, of the same forward
    with the autograd re-evaluation backward, and (context, like torch_eager_gpu) of the reference algorithm as
    PyTorch eager ops on this GPU (oracle/ref_port.py).  The UNet3d encoder is out of scope: the latent grid is a leaf."""
    import space_time_pde_b200 as sp
    B, P, nf = 10, 1024, 32
    torch.manual_seed(0)
    model = sp.ImNet(dim=3, in_features=CHANNELS, out_features=4, nf=nf, activation=sp.NONLINEARITIES[ACT]).to(device)
    grid = (torch.randn(B, *GRID, CHANNELS) * 0.5).to(device).requires_grad_(True)
    q = torch.rand(B, P, 3, device=device)
    target = torch.randn(B, P, 4, device=device)
    layer = sp.get_rb2_pde_layer(**RB2)
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))

    def fused_step():
        model.zero_grad(set_to_none=True)
        grid.grad = None
        y, res = layer(q, return_residue=True)
        loss = torch.nn.functional.l1_loss(y, target) + 0.0125 * torch.stack(list(res.values())).abs().mean()
        loss.backward()
        return loss

    def time_it(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    out = {"config": f"{B} crops x {P} points, ImNet nf={nf} {ACT}, latent 4x16x16x32, RB2 + continuity, L1 losses"}
    out["fused_ms"] = time_it(fused_step, 20)
    out["points_per_s"] = B * P / (out["fused_ms"] * 1e-3)
    os.environ["STPDE_BACKWARD"] = "torch"
    os.environ["STPDE_RESIDUALS"] = "torch"
    try:
        out["torch_autograd_backward_ms"] = time_it(fused_step, 5)
    finally:
        os.environ.pop("STPDE_BACKWARD", None)
        os.environ.pop("STPDE_RESIDUALS", None)
    if with_eager:
        from oracle import jet_oracle as jo
        from oracle import ref_port as rp
        port = rp.SkipMLP([l.weight.detach().cpu().numpy() for l in model.fc],
                          [l.bias.detach().cpu().numpy() for l in model.fc], ACT).to(device)
        iv, ov, eqs = jo.rb2_equations(**RB2)
        exprs = rp.compile_equations(eqs)
        grid2 = grid.detach().clone().requires_grad_(True)

        def eager_step():
            port.zero_grad(set_to_none=True)
            grid2.grad = None
            y, res = rp.values_and_residuals(port, grid2, q, 0., 1., iv, ov, exprs)
            loss = torch.nn.functional.l1_loss(y, target) + 0.0125 * torch.stack(list(res.values())).abs().mean()
            loss.backward()
            return loss

        try:
            out["reference_algorithm_eager_gpu_ms"] = time_it(eager_step, 3)
        except Exception as exc:
            out["reference_algorithm_eager_gpu_ms"] = None
            out["eager_error"] = str(exc)[:100]
    return out


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    rate, _ = cpu_reference_rate(64, repeats=1, warm_pts=64)                    # calibration
    budget = 150.0 / max(1, args.steps + args.warmup)
    sample = int(min(4096, max(64, rate * budget)) // 64 * 64)
    from oracle import jet_oracle as jo
    from oracle import ref_port as rp
    model = make_model("cpu")
    port = rp.SkipMLP([l.weight.detach().numpy() for l in model.fc], [l.bias.detach().numpy() for l in model.fc], ACT)
    iv, ov, eqs = jo.rb2_equations(**RB2)
    exprs = rp.compile_equations(eqs)
    grid, q = synthetic_inputs(1234, "cpu", sample)
    for _ in range(args.warmup):
        rp.values_and_residuals(port, grid, q, 0., 1., iv, ov, exprs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rp.values_and_residuals(port, grid, q, 0., 1., iv, ov, exprs)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config("cpu"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} of {NPTS} query points per step (same seeded workload), "
                                       f"oracle/ref_port.py = reference algorithm (torch CPU, one autograd.grad per dif)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(precision):
    return {"workload": "rb2d config[1]: latent 4x16x16x32 (synthetic UNet3d-shaped), ImNet nf=128 Softplus, "
                        "2^20 query points per GPU, values + 3 RB2 transport residuals + continuity",
            "imnet_nf": NF, "latent_grid": list(GRID) + [CHANNELS], "points_per_gpu": NPTS, "jet_components": KC,
            "precision": precision,
            "flops_per_point": flops_per_point(NF, 3, CHANNELS, 4, KC),
            "l2": "per-chunk activation scratch (GBs) >> 126 MB L2: every step streams from HBM, no flush needed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("STPDE_PRECISION", "fp16x3"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--train-chunk", type=int, default=16384, help="query points per forward/backward chunk of the training leg")
    ap.add_argument("--train-workspace-mb", type=int, default=40960, help="workspace budget of the training leg (stash of one chunk)")
    ap.add_argument("--train-backward-precision", default="same", choices=["same", "fp16x3", "fp16"],
                    help="arithmetic of the reverse sweep's contractions (same = the forward's parity mode)")
    ap.add_argument("--train-steps", type=int, default=1,
                    help="timed training steps (forward + residuals + loss + fused backward [+ all-reduce]); 0 skips the leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch.distributed as dist
    import space_time_pde_b200 as sp
    from space_time_pde_b200 import _lib, jets

    assert torch.cuda.is_available(), "bench.py needs a GPU (the hot path has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    jets.set_default_precision(args.precision)
    lib = _lib.load()

    model = make_model(device)
    grid, q = synthetic_inputs(1234 + rank, device)
    layer = sp.get_rb2_pde_layer(**RB2)
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))

    def step():
        with torch.no_grad():
            return layer(q, return_residue=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    _lib.profile_read()                                   # reset launch counters
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = sum(c for _, c in _lib.profile_read().values())
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * NPTS * args.steps / (ms * 1e-3)

    # ---- e2e: public API, host (pinned) inputs, results read back to the host, every step ----
    q_host = q.cpu().pin_memory()
    grid_host = grid.cpu().pin_memory()
    y_host = torch.empty(1, NPTS, 4).pin_memory()
    res_host = torch.empty(4, 1, NPTS, 1).pin_memory()
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step():
        qd = q_host.to(device, non_blocking=True)
        gd = grid_host.to(device, non_blocking=True)
        layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, gd, pts, 0., 1.))
        with torch.no_grad():
            y, res = layer(qd, return_residue=True)
        y_host.copy_(y, non_blocking=True)
        res_host.copy_(torch.stack(list(res.values())), non_blocking=True)
        torch.cuda.synchronize()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * NPTS * e2e_steps / float(e2e_s.item())
    h2d = q_host.numel() * 4 + grid_host.numel() * 4
    d2h = y_host.numel() * 4 + res_host.numel() * 4
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))

    # ---- roofline: dominant kernel timed alone with CUDA events on its launch stream ----
    peaks = measured_peaks()
    lib.stpde_profile_enable(1)
    _lib.profile_read()
    step()
    prof = _lib.profile_read()
    lib.stpde_profile_enable(0)
    kernel_ms = {k: v[0] for k, v in prof.items() if v[1] > 0}
    dom = "gemm_layer1"
    n1, k1 = 8 * NF, 16 * NF
    dom_flops = 2.0 * n1 * k1 * NPTS * 8 * KC                   # algorithmic FLOPs of that layer per step
    dom_ms = kernel_ms.get(dom, float("nan"))
    achieved = dom_flops / (dom_ms * 1e-3) / 1e12
    fpt = flops_per_point(NF, 3, CHANNELS, 4, KC)
    step_tflops = (value / world) * fpt / 1e12
    passes = {"fp16x3": 3, "fp16": 1}.get(args.precision, 0)        # tensor-core MMAs per algorithmic product
    n_launch = max(1, int(prof.get(dom, (0, 1))[1]))
    traffic = None                                                  # DRAM bytes per launch from the committed ncu capture
    ncu_path = os.path.join(ROOT, "profiles", "ncu_dominant_kernel.json")
    if os.path.exists(ncu_path):
        cap = json.load(open(ncu_path)).get(args.precision)
        if cap:
            rows_per_launch = NPTS * 8 / n_launch
            traffic = cap["dram_bytes_per_launch"] * rows_per_launch / cap["rows_per_launch"]
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tflops"], "traffic": traffic, "peak_source": peaks["source"],
                "launches": n_launch, "avg_launch_ms": dom_ms / n_launch, "algorithmic_flops_per_launch": dom_flops / n_launch,
                "tensor_passes": passes, "executed_tensor_tflops": achieved * passes if passes else None,
                "executed_frac": achieved * passes / peaks["tflops"] if passes else None,
                "launch_ms_total": dom_ms, "kernel_share_of_step": dom_ms / sum(kernel_ms.values()),
                "step_achieved": step_tflops, "step_frac": step_tflops / peaks["tflops"],
                "kernel_ms": kernel_ms,
                "note": "achieved = algorithmic FLOPs (2*K*N*rows*jet components) / CUDA-event time; "
                        "fp32 FFMA and fp16x3 execute more than the algorithmic FLOPs, the fraction is of the "
                        "measured bf16 tensor peak"}

    # ---- training step: forward + residuals + L1 losses + fused CUDA backward to the latent grid
    #      and the decoder weights, then ONE all-reduce of [loss sums | counts | flat gradients] ----
    train = None
    if args.train_steps > 0:
        try:
            from space_time_pde_b200.parallel import StepReducer
            grid_t = grid.clone().requires_grad_(True)
            params = [grid_t] + list(model.parameters())
            for p_ in params[1:]:
                p_.requires_grad_(True)
            reducer = StepReducer(params)
            layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid_t, pts, 0., 1.))

            # Points are independent, so the step walks the batch in chunks that fit the training stash (the forward of
            # a chunk leaves its operand planes in the workspace and the chunk's backward reuses them: no recompute);
            # gradients accumulate in .grad across chunks exactly as one big backward would.
            os.environ["STPDE_WORKSPACE_MB"] = str(args.train_workspace_mb)
            jets.release_workspaces()
            jets.set_backward_precision(args.train_backward_precision)
            tchunk = args.train_chunk

            def train_step():
                for p_ in params:
                    p_.grad = None
                reg_sum = torch.zeros((), device=device)
                pde_sum = torch.zeros((), device=device)
                for s0 in range(0, NPTS, tchunk):
                    y, res = layer(q[:, s0:s0 + tchunk], return_residue=True)
                    reg = y.abs().sum()
                    pde = torch.stack(list(res.values())).abs().sum()
                    (reg / (world * 4 * NPTS) + 0.0125 * pde / (world * 4 * NPTS)).backward()
                    reg_sum += reg.detach()
                    pde_sum += pde.detach()
                return reducer.reduce({"reg": reg_sum, "pde": pde_sum}, {"reg": 4 * NPTS, "pde": 4 * NPTS})

            train_step()
            barrier()
            lib.stpde_profile_enable(1)
            _lib.profile_read()
            tv0, tv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tv0.record()
            for _ in range(args.train_steps):
                means = train_step()
            tv1.record()
            barrier()
            prof_t = _lib.profile_read()
            lib.stpde_profile_enable(0)
            tms = torch.tensor([tv0.elapsed_time(tv1)], device=device, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            tms = float(tms.item()) / args.train_steps
            # algorithmic FLOPs of a training step = 3 x the forward contractions (forward, dgrad, wgrad); the recompute
            # of the forward inside the backward is overhead, not counted
            train = {"value": world * NPTS / (tms * 1e-3), "unit": "points/s", "ms_per_step": tms, "steps": args.train_steps,
                     "chunk_points": tchunk, "workspace_mb": args.train_workspace_mb,
                     "backward_precision": args.precision if args.train_backward_precision == "same" else args.train_backward_precision,
                     "what": "values + RB2 residuals + L1 losses + fused CUDA reverse sweep (grid + decoder gradients), "
                             "chunks of the batch with the forward planes kept for the backward (no recompute)"
                             + (" + one NCCL all-reduce of the flat gradient buffer" if world > 1 else ""),
                     "algorithmic_tflops": 3.0 * NPTS * fpt / (tms * 1e-3) / 1e12,
                     "frac_of_peak": 3.0 * NPTS * fpt / (tms * 1e-3) / 1e12 / peaks["tflops"],
                     "loss_reg": float(means["reg"]), "loss_pde": float(means["pde"]),
                     "kernel_ms_per_step": {k: v[0] / args.train_steps for k, v in prof_t.items() if v[1] > 0},
                     "gpu_launches_per_step": int(sum(v[1] for v in prof_t.values()) / args.train_steps)}
            for p_ in params:
                p_.grad = None
            jets.release_workspaces()
            jets.set_backward_precision("same")
            layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
        except Exception as exc:   # the headline metric must survive a failing optional leg (e.g. out of memory)
            if world > 1:
                raise
            train = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
            lib.stpde_profile_enable(0)
            jets.release_workspaces()
            jets.set_backward_precision("same")
            layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
        os.environ.pop("STPDE_WORKSPACE_MB", None)

    if world > 1:
        dist.barrier()
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            rate, secs = cpu_reference_rate(8192, repeats=1)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"8192 of {NPTS} points, {secs:.1f} s, oracle/ref_port.py (torch CPU, autograd per dif), "
                             f"{cores} threads"}
            try:   # context only: the same reference algorithm as PyTorch eager ops on this GPU (true fp32, no TF32)
                eager = gpu_eager_rate(device)
            except Exception as exc:   # e.g. out of memory for the autograd tapes
                eager = {"error": str(exc)[:120]}
        if train is not None and world == 1:
            try:
                train["reference_size_step"] = reference_size_train_step(device, with_eager=not args.no_cpu_baseline)
            except Exception as exc:
                train["reference_size_step"] = {"error": str(exc)[:200]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args.precision), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "train_step": train}
        if cpu is not None:
            line["torch_eager_gpu"] = eager
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
