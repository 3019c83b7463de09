#!/usr/bin/env python
"""Benchmark of the decode + PDE-residual hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision fp32|fp16x3|fp16]

Headline (BASELINE config[1]): one "step" = one pass of the hot path over one batch of synthetic query points: values +
all Rayleigh-Benard residuals (3 transport equations + continuity) for 2^20 points against a 4x16x16x32 latent grid with
ImNet(nf=128, Softplus), float32 I/O.

  value : whole-job throughput with inputs resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through the public API with HOST (pinned) buffers; H2D of the query points and latent grid and
          D2H of values + residuals are inside the timed region
  roofline : dominant kernel (hidden layer 1 contraction) timed with CUDA events on its stream
  parity : rel-L-infinity of the timed configuration against the fp64 oracle on a 256-point slice (checker only)
  cpu_baseline : the reference's own modules (oracle/_ref, staged copy of the unmodified reference) or, if that is not
                 staged, the restatement oracle/ref_port.py, on the host cores, bounded sample
  configs : the other BASELINE configurations, each with its own roofline / parity figure (short legs):
      config2_train  config[1] as a training step (fused reverse sweep + one all-reduce)
      config3        paper training shape: 8 M points per step split over the ranks, ImNet nf=32, 16-bit MLP operands
      config4        4-d Navier-Stokes strings, ImNet nf=256, K = 8 jet components
      config5        sweep 1e4 / 1e6 / 6.4e7 points, latent 32^3 x 128, FIXED TOTAL split over the ranks (strong scaling)
      eval_grid      evaluation.py's structured 192 x 128 x 512 query grid in one call
Under torchrun each rank processes its own 2^20 points for the headline (weak scaling, no data-path collective).
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


def make_model(device, seed=0, nf=NF, channels=CHANNELS, dim=3):
    import space_time_pde_b200 as sp
    torch.manual_seed(seed)
    model = sp.ImNet(dim=dim, in_features=channels, out_features=4, nf=nf, activation=sp.NONLINEARITIES[ACT])
    return model.to(device)


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's own modules when the staged copy exists (oracle/_ref), else the restatement
# ----------------------------------------------------------------------------------------------
class CpuReference:
    """values + RB2 residuals of the headline configuration on the host cores."""

    def __init__(self):
        from oracle import stage_ref
        torch.set_num_threads(os.cpu_count() or 1)
        model = make_model("cpu")
        if stage_ref.available():
            self.kind = "reference"
            self.what = ("oracle/_ref = the reference's own unmodified src/ + experiments/rb2d modules "
                         "(PDELayer -> query_local_implicit_grid -> ImNet, one torch.autograd.grad per dif)")
            self.pipe = stage_ref.ReferencePipeline(NF, CHANNELS, ACT, model.state_dict(), RB2)
            self.run = lambda grid, q: self.pipe(grid, q, 0., 1.)
        else:
            from oracle import jet_oracle as jo
            from oracle import ref_port as rp
            self.kind = "port"
            self.what = "oracle/ref_port.py = restatement of the reference algorithm (torch CPU, one autograd.grad per dif)"
            port = rp.SkipMLP([l.weight.detach().numpy() for l in model.fc], [l.bias.detach().numpy() for l in model.fc], ACT)
            iv, ov, eqs = jo.rb2_equations(**RB2)
            exprs = rp.compile_equations(eqs)
            self.run = lambda grid, q: rp.values_and_residuals(port, grid, q, 0., 1., iv, ov, exprs)

    def rate(self, sample_pts, repeats=1, warm_pts=128):
        grid, q = synthetic_inputs(1234, "cpu", max(sample_pts, warm_pts))
        self.run(grid, q[:, :warm_pts])                                   # warm-up (sympy lambdas, thread pool)
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            self.run(grid, q[:, :sample_pts])
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return sample_pts / best, best


def gpu_eager_rate(device, batch=2048, batches=3):
    """Reference algorithm (oracle/ref_port.py, autograd per dif) as PyTorch eager ops on the GPU, pseudo-batched
    like the reference's evaluation loop (experiments/rb2d/evaluation.py:54-69).  Context only."""
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


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    ref = CpuReference()
    rate, _ = ref.rate(64, repeats=1, warm_pts=64)                              # calibration
    budget = 150.0 / max(1, args.steps + args.warmup)
    sample = int(min(4096, max(64, rate * budget)) // 64 * 64)
    grid, q = synthetic_inputs(1234, "cpu", sample)
    for _ in range(args.warmup):
        ref.run(grid, q)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.run(grid, q)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config("cpu"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": ref.kind,
                             "sample": f"{sample} of {NPTS} query points per step (same seeded workload), {ref.what}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(precision):
    return {"workload": "rb2d config[1]: latent 4x16x16x32 (synthetic UNet3d-shaped), ImNet nf=128 Softplus, "
                        "2^20 query points per GPU, values + 3 RB2 transport residuals + continuity",
            "imnet_nf": NF, "latent_grid": list(GRID) + [CHANNELS], "points_per_gpu": NPTS, "jet_components": KC,
            "precision": precision,
            "arithmetic": {"fp16x3": "tcgen05 kind::f16, operands split hi+lo in fp16 (22 significant bits), 3 MMAs per "
                                     "product, fp32 accumulation in TMEM; fp32 I/O and jets",
                           "fp16": "tcgen05 kind::f16 single pass (11-bit operands), fp32 accumulation; fp32 I/O and jets",
                           "fp32": "FP32 FFMA on the CUDA cores"}.get(precision, precision),
            "flops_per_point": flops_per_point(NF, 3, CHANNELS, 4, KC),
            "l2": "per-chunk activation scratch (GBs) >> 126 MB L2: every step streams from HBM, no flush needed"}


# ----------------------------------------------------------------------------------------------
# helpers of the GPU legs
# ----------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, device, rank, world):
        self.device, self.rank, self.world = device, rank, world

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_ms(self, ms):
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=self.device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return float(ms)

    def all_ok(self, ok):
        """True only if `ok` holds on EVERY rank (an optional path must be taken or skipped by all ranks alike: its timing
        runs barriers and an all-reduce)."""
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([1.0 if ok else 0.0], device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return bool(t.item() > 0.5)
        return bool(ok)

    def time(self, fn, steps, warmup):
        """ms per step: CUDA events on the current stream, barrier + synchronize on both sides, max over ranks."""
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        return self.max_ms(e0.elapsed_time(e1)) / steps


def rel_linf(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


def oracle_parity(model, grid, q, y, res, n, rb2_kwargs=None, equations=None, act=ACT):
    """rel-Linf of (values, residuals) on the first n points of batch 0 against the fp64 numpy oracle (CHECKER: runs
    outside every timed region)."""
    from oracle import jet_oracle as jo
    Ws = [l.weight.detach().cpu().numpy() for l in model.fc]
    bs = [l.bias.detach().cpu().numpy() for l in model.fc]
    qn = q[:1, :n].detach().cpu().numpy()
    yj = jo.query_jet(grid[:1].detach().cpu().numpy(), qn, 0., 1., Ws, bs, act)
    if equations is None:
        iv, ov, eqs = jo.rb2_equations(**rb2_kwargs)
    else:
        iv, ov, eqs = equations
    ref = jo.pde_residuals(yj, qn, iv, ov, eqs)
    out = {"points": n, "values": rel_linf(y[:1, :n].detach().cpu().numpy(), yj.v)}
    out["residuals"] = max(rel_linf(res[k][:1, :n].detach().cpu().numpy(), ref[k]) for k in ref)
    out["rel_linf_vs_fp64"] = max(out["values"], out["residuals"])
    return out


def imnet_widths(nf):
    """Hidden widths of ImNet (reference src/implicit_net.py:28-33): nf*16, nf*8, nf*4, nf*2, nf."""
    return [nf * 16, nf * 8, nf * 4, nf * 2, nf]


def plane_bytes_per_point(nf, d, kc, precision, training=False, fused_final=True):
    """ALGORITHMIC HBM bytes per query point of the layer-by-layer design: every activation plane is
    written once and read once by the next layer; a training step adds the fp32 pre-activations (written by the forward,
    read by the dgrad epilogues; fp16 when forward and reverse sweep both run in the single-pass mode), the zbar planes and the wgrad operand reads.  Planes are [kc][rows][width padded to 64]
    fp16, hi + lo in the 3-pass mode.  At ImNet nf = 32 this, not the tensor pipe, bounds the step."""
    P = 2 if precision == "fp16x3" else 1
    w = imnet_widths(nf)
    ld = [(x + 63) // 64 * 64 for x in w]
    np_last = (w[-1] + 15) // 16 * 16
    a = lambda l: ld[l] * kc * 2 * P                 # activation / zbar plane bytes per row of layer l
    z = lambda l: ld[l] * kc * (2 if precision == "fp16" else 4)   # saved pre-activations: fp16 in the single-pass mode, else fp32
    n = len(w)
    b = a(0)                                         # layer 0 writes its planes
    for l in range(1, n):
        b += a(l - 1)                                # operand read
        if l < n - 1:
            b += a(l)
        elif training or not fused_final:
            b += np_last * kc * 4                    # fp32 plane of the last hidden layer
        if training:
            b += z(l)
    if training:
        b += z(n - 1) + np_last * kc * 4 + a(n - 1)              # blend_backward
        for l in range(n - 1, 0, -1):
            b += a(l - 1) + a(l)                                 # wgrad operands
            b += a(l)                                            # dgrad operand
            if l >= 2:
                b += z(l - 1) + a(l - 1)                         # saved pre-activations in, zbar planes out
    return b * (1 << d)


def hbm_roofline_of(rate_per_gpu, bytes_per_point, peaks):
    gbs = rate_per_gpu * bytes_per_point / 1e9
    return {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
            "bytes_per_point": bytes_per_point,
            "what": "algorithmic plane bytes per point (each plane written once, read once; training adds the saved "
                    "pre-activations, the zbar planes and the wgrad operand reads) over the measured copy bandwidth"}


def roofline_of(rate_per_gpu, fpt, peaks, passes, mult=1.0):
    tf = rate_per_gpu * fpt * mult / 1e12
    return {"bound": "tensor", "achieved": tf, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": tf / peaks["tflops"],
            "flops_per_point": fpt * mult, "tensor_passes": passes,
            "executed_frac": tf * passes / peaks["tflops"] if passes else None}


def graphed_chunk_step(layer, params, q, target, chunk, loss_of_sums, device):
    """One chunk of a training step (fused loss sums + reverse sweep, gradients ACCUMULATED into pre-existing .grad tensors)
    recorded into CUDA graphs and replayed for every chunk of the batch: the ~85 launches of a chunk stop paying a launch
    gap each (8 % of the BASELINE config-3 step).  Two graphs, because the call-invariant setup belongs to the FIRST chunk
    after a weight update only: graph A rebuilds it (split weights, per-vertex table), graph B reuses it.  The library never
    allocates, synchronises or reads device values on the host, so torch.cuda.graph can record the calls; their status words
    stay on the device and are checked once per step.  Returns step() -> (reg_sum, pde_sum) or raises (the caller falls
    back to the eager chunk loop)."""
    from space_time_pde_b200 import jets
    n = q.shape[1]
    if n % chunk:
        raise ValueError("graphed chunks need equal chunk sizes")
    static_q = q[:, :chunk].clone()
    static_t = target[:, :chunk].clone() if target is not None else None
    reg_acc = torch.zeros((), device=device)
    pde_acc = torch.zeros((), device=device)
    for p_ in params:
        p_.grad = torch.zeros_like(p_)                            # AccumulateGrad adds in place: the graphs accumulate
    static_grads = [p_.grad for p_ in params]                     # ... into exactly these tensors

    def body():
        y, sums, _ = layer.loss_sums(static_q, static_t, "l1")
        loss_of_sums(sums).backward()
        reg_acc.add_(sums[0].detach())
        pde_acc.add_(sums[1].detach())

    old_cache = os.environ.get("STPDE_SETUP_CACHE")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    try:
        os.environ["STPDE_SETUP_CACHE"] = "0"
        with torch.cuda.stream(side):
            for _ in range(2):
                body()
        torch.cuda.current_stream().wait_stream(side)
        g_first = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_first, stream=side):
            body()
        os.environ["STPDE_SETUP_CACHE"] = "1"
        with torch.cuda.stream(side):
            for _ in range(2):                                    # (the second call finds the first one's setup key)
                body()
        torch.cuda.current_stream().wait_stream(side)
        g_rest = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_rest, stream=side):
            body()
    finally:
        if old_cache is None:
            os.environ.pop("STPDE_SETUP_CACHE", None)
        else:
            os.environ["STPDE_SETUP_CACHE"] = old_cache
    torch.cuda.synchronize()

    def step():
        for p_, g_ in zip(params, static_grads):
            g_.zero_()
            p_.grad = g_                                          # (whatever the caller did to .grad in between)
        reg_acc.zero_()
        pde_acc.zero_()
        for i, s0 in enumerate(range(0, n, chunk)):
            static_q.copy_(q[:, s0:s0 + chunk])
            if static_t is not None:
                static_t.copy_(target[:, s0:s0 + chunk])
            (g_first if i == 0 else g_rest).replay()
        jets.check_captured()                                     # status words of the captured calls (one sync per step)
        return reg_acc, pde_acc

    step.graphs = (g_first, g_rest)
    return step


# ----------------------------------------------------------------------------------------------
# legs for the other BASELINE configurations
# ----------------------------------------------------------------------------------------------
def leg_config2_train(ctx, args, model, grid, q, layer, peaks, lib):
    """config[1] as a training step: forward + residuals + L1 losses + fused CUDA backward to the
    latent grid and the decoder weights, then ONE all-reduce of [loss sums | counts | flat gradients]."""
    import space_time_pde_b200 as sp
    from space_time_pde_b200 import _lib, jets
    from space_time_pde_b200.parallel import StepReducer
    device, world = ctx.device, ctx.world
    fpt = flops_per_point(NF, 3, CHANNELS, 4, KC)
    grid_t = grid.clone().requires_grad_(True)
    params = [grid_t] + list(model.parameters())
    reducer = StepReducer(params)
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid_t, pts, 0., 1.))
    # Points are independent, so the step walks the batch in chunks that fit the training stash (the forward of a chunk
    # leaves its operand planes in the workspace and the chunk's backward reuses them: no recompute); gradients
    # accumulate in .grad across chunks exactly as one big backward would.
    os.environ["STPDE_WORKSPACE_MB"] = str(args.train_workspace_mb)
    jets.release_workspaces()
    jets.set_backward_precision(args.train_backward_precision)
    tchunk = args.train_chunk
    out = {}

    def train_step():
        for p_ in params:
            p_.grad = None
        reg_sum = torch.zeros((), device=device)
        pde_sum = torch.zeros((), device=device)
        # deferred_checks: the calls' status words are read once at the end of the step (raised there), so the host
        # launches chunk n+1 while the device still works on chunk n instead of idling through two round trips per chunk
        with sp.deferred_checks():
            for s0 in range(0, NPTS, tchunk):
                # residual programs + L1 reductions in ONE kernel (per-CTA partial sums; no [4,b,p] tensor, no torch.stack)
                y, sums, _ = layer.loss_sums(q[:, s0:s0 + tchunk], None, "l1")
                (sums[0] / (world * 4 * NPTS) + 0.0125 * sums[1] / (world * 4 * NPTS)).backward()
                reg_sum += sums[0].detach()
                pde_sum += sums[1].detach()
        out["means"] = reducer.reduce({"reg": reg_sum, "pde": pde_sum}, {"reg": 4 * NPTS, "pde": 4 * NPTS})

    try:
        train_step()
        ctx.barrier()
        lib.stpde_profile_enable(1)
        _lib.profile_read()
        tms = ctx.time(train_step, args.train_steps, 0)
        prof_t = _lib.profile_read()
        lib.stpde_profile_enable(0)
        means = out["means"]
        eager_ms, graph_info = tms, None
        if args.graph_chunks:
            # the same step with each chunk replayed from a CUDA graph (same kernels, same losses; falls back to the loop)
            eager_means = {k: float(v) for k, v in means.items()}
            denom = float(world * 4 * NPTS)
            gstep = None
            try:                                                  # phase 1 (local): record the graphs
                gstep = graphed_chunk_step(layer, params, q, None, tchunk,
                                           lambda sums: sums[0] / denom + 0.0125 * sums[1] / denom, device)
            except Exception as exc:                              # noqa: BLE001 - optional path
                graph_info = {"error": str(exc)[:200]}
                jets._captured.clear()
            if ctx.all_ok(gstep is not None):                     # phase 2 (collective): every rank or none

                def graph_step():
                    reg_sum, pde_sum = gstep()
                    out["means"] = reducer.reduce({"reg": reg_sum.clone(), "pde": pde_sum.clone()},
                                                  {"reg": 4 * NPTS, "pde": 4 * NPTS})

                graph_step()
                gms = ctx.time(graph_step, args.train_steps, 0)
                diff = max(abs(float(out["means"][k]) - eager_means[k]) / max(abs(eager_means[k]), 1e-30) for k in eager_means)
                graph_info = {"ms_per_step": gms, "loss_rel_diff_vs_eager_loop": diff}
                if diff < 1e-5 and gms < tms:
                    tms = gms
            elif graph_info is None:
                graph_info = {"error": "graph capture failed on another rank"}
            gstep = None
            jets._captured.clear()
        res = {"value": world * NPTS / (tms * 1e-3), "unit": "points/s", "ms_per_step": tms, "steps": args.train_steps,
               "chunk_points": tchunk, "workspace_mb": args.train_workspace_mb,
               "eager_chunk_loop_ms_per_step": eager_ms, "cuda_graph_chunks": graph_info,
               "backward_precision": args.precision if args.train_backward_precision == "same" else args.train_backward_precision,
               "what": "config[1] training step: values + RB2 residuals + L1 losses + fused CUDA reverse sweep (grid + decoder "
                       "gradients), chunks of the batch with the forward planes kept for the backward (no recompute)"
                       + (" + one NCCL all-reduce of the flat gradient buffer" if world > 1 else ""),
               "roofline": roofline_of(NPTS / (tms * 1e-3), fpt, peaks, 3 if args.precision == "fp16x3" else 1, mult=3.0),
               "loss_reg": float(means["reg"]), "loss_pde": float(means["pde"]),
               "kernel_ms_per_step": {k: v[0] / args.train_steps for k, v in prof_t.items() if v[1] > 0},
               "gpu_launches_per_step": int(sum(v[1] for v in prof_t.values()) / args.train_steps)}
    finally:
        lib.stpde_profile_enable(0)
        for p_ in params:
            p_.grad = None
        jets.release_workspaces()
        jets.set_backward_precision("same")
        os.environ.pop("STPDE_WORKSPACE_MB", None)
        layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
    return res


def leg_config3(ctx, args, peaks):
    """BASELINE config[2] (paper training configuration): 8 M query points per step = 8 crops x 2^20 points SPLIT over
    the ranks, ImNet nf=32 Softplus, normalised RB2 equations, alpha_pde = 0.0125, 16-bit MLP operands (single-pass fp16
    tensor-core contractions, fp32 accumulation, fp32 jets) in both sweeps, L1 losses, one all-reduce per step."""
    import space_time_pde_b200 as sp
    from space_time_pde_b200 import jets
    from space_time_pde_b200.parallel import StepReducer
    device, rank, world = ctx.device, ctx.rank, ctx.world
    B, P_total, nf = 8, (1 << 20), 32
    p_rank = P_total // world
    rb2 = dict(mean=[0.1, -0.2, 0.05, 0.3], std=[1.1, 0.9, 1.3, 0.7], t_crop=2., z_crop=1., x_crop=2., prandtl=1.,
               rayleigh=1e6, use_continuity=True)
    model = make_model(device, seed=3, nf=nf)
    g = torch.Generator().manual_seed(300)
    grid = (torch.randn(B, *GRID, CHANNELS, generator=g) * 0.5).to(device).requires_grad_(True)
    gq = torch.Generator().manual_seed(301 + rank)
    q = (torch.rand(B, p_rank, 3, generator=gq) * (1 - 2e-6) + 1e-6).to(device)
    target = torch.randn(B, p_rank, 4, generator=gq).to(device)
    layer = sp.get_rb2_pde_layer(**rb2)
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
    params = [grid] + list(model.parameters())
    reducer = StepReducer(params)
    n_glob = float(B * P_total * 4)
    chunk = min(p_rank, max(2048, args.config3_chunk // B))          # points per crop and chunk
    old_prec = jets.DEFAULT_PRECISION
    os.environ["STPDE_WORKSPACE_MB"] = str(args.train_workspace_mb)
    jets.release_workspaces()
    jets.set_default_precision("fp16")
    jets.set_backward_precision("fp16")
    out = {}

    def step():
        for p_ in params:
            p_.grad = None
        reg_sum = torch.zeros((), device=device)
        pde_sum = torch.zeros((), device=device)
        with sp.deferred_checks():                                # status words read once per step, not twice per chunk
            for s0 in range(0, p_rank, chunk):
                y, sums, _ = layer.loss_sums(q[:, s0:s0 + chunk], target[:, s0:s0 + chunk], "l1")
                (sums[0] / n_glob + 0.0125 * sums[1] / n_glob).backward()
                reg_sum += sums[0].detach()
                pde_sum += sums[1].detach()
        out["means"] = reducer.reduce({"reg": reg_sum, "pde": pde_sum}, {"reg": n_glob / world, "pde": n_glob / world})

    try:
        from space_time_pde_b200 import _lib
        lib = _lib.load()
        step()                                                    # warm-up
        ctx.barrier()
        lib.stpde_profile_enable(1)
        _lib.profile_read()
        ms = ctx.time(step, 1, 0)
        prof = _lib.profile_read()
        lib.stpde_profile_enable(0)
        eager_ms, graph_info = ms, None
        if args.graph_chunks:
            # the same step with each chunk replayed from a CUDA graph (same kernels, same losses; falls back to the loop)
            eager_means = {k: float(v) for k, v in out["means"].items()}
            gstep = None
            try:                                                  # phase 1 (local): record the graphs
                gstep = graphed_chunk_step(layer, params, q, target, chunk,
                                           lambda sums: sums[0] / n_glob + 0.0125 * sums[1] / n_glob, device)
            except Exception as exc:                              # noqa: BLE001 - optional path
                graph_info = {"error": str(exc)[:200]}
                jets._captured.clear()
            if ctx.all_ok(gstep is not None):                     # phase 2 (collective): every rank or none

                def graph_step():
                    reg_sum, pde_sum = gstep()
                    out["means"] = reducer.reduce({"reg": reg_sum.clone(), "pde": pde_sum.clone()},
                                                  {"reg": n_glob / world, "pde": n_glob / world})

                graph_step()
                gms = ctx.time(graph_step, 1, 0)
                diff = max(abs(float(out["means"][k]) - eager_means[k]) / max(abs(eager_means[k]), 1e-30) for k in eager_means)
                graph_info = {"ms_per_step": gms, "loss_rel_diff_vs_eager_loop": diff}
                if diff < 1e-5 and gms < ms:
                    ms = gms
            elif graph_info is None:
                graph_info = {"error": "graph capture failed on another rank"}
            gstep = None
            jets._captured.clear()
            out["means"] = {k: torch.tensor(v) for k, v in eager_means.items()}
        with torch.no_grad():
            y, res = layer(q[:1, :512], return_residue=True)
        par = oracle_parity(model, grid, q, y, res, 256, rb2_kwargs=rb2)
        fpt = flops_per_point(nf, 3, CHANNELS, 4, KC)
        rate = B * P_total / (ms * 1e-3)
        return {"workload": f"paper training step: {B} crops x 2^20 points per step split over {world} rank(s) "
                            f"({B} x {p_rank} per rank), ImNet nf={nf} Softplus, normalised RB2 + continuity, L1 losses, "
                            "alpha_pde 0.0125, fused reverse sweep, one all-reduce of [loss sums | counts | gradients]",
                "imnet_nf": nf, "jet_components": KC, "dtype": "f16 operands (single tcgen05 pass), f32 accumulate / jets / I/O",
                "value": rate, "unit": "points/s (training step)", "ms_per_step": ms, "scaling": "strong (fixed 8 M points)",
                "eager_chunk_loop_ms_per_step": eager_ms, "cuda_graph_chunks": graph_info,
                "points_per_rank": B * p_rank, "chunk_points": B * chunk,
                "roofline": roofline_of(rate / world, fpt, peaks, 1, mult=3.0),
                "roofline_hbm": hbm_roofline_of(rate / world, plane_bytes_per_point(nf, 3, KC, "fp16", training=True), peaks),
                "parity": dict(par, mode="relaxed 16-bit mode: forward values + residuals vs fp64 oracle"),
                "kernel_ms_per_step": {k: round(v[0], 2) for k, v in prof.items() if v[1] > 0},
                "gpu_launches_per_step": int(sum(v[1] for v in prof.values())),
                "loss_reg": float(out["means"]["reg"]), "loss_pde": float(out["means"]["pde"])}
    finally:
        for p_ in params:
            p_.grad = None
        jets.set_default_precision(old_prec)
        jets.set_backward_precision("same")
        jets.release_workspaces()
        os.environ.pop("STPDE_WORKSPACE_MB", None)


def ns4d_layer():
    import space_time_pde_b200 as sp
    layer = sp.PDELayer(in_vars="x, y, z, t", out_vars="u, v, w, p")
    lap = lambda f: f"(dif(dif({f},x),x)+dif(dif({f},y),y)+dif(dif({f},z),z))"
    adv = lambda f: f"(u*dif({f},x)+v*dif({f},y)+w*dif({f},z))"
    eqs = {}
    for f in "uvw":
        eqs["mom_" + f] = (f"dif({f},t)+{adv(f)}+dif(p,{'xyz'['uvw'.index(f)]})-0.01*{lap(f)}", None)
    eqs["continuity"] = ("dif(u,x)+dif(v,y)+dif(w,z)", None)
    for k, (s, _) in eqs.items():
        layer.add_equation(s, k)
    return layer, (("x", "y", "z", "t"), ("u", "v", "w", "p"), eqs)


def leg_config4(ctx, args, peaks, precision):
    """BASELINE config[3]: custom PDELayer strings, unsteady 3-D incompressible Navier-Stokes with Laplacians
    (d = 4: x, y, z, t), latent 8^4 x 32, ImNet nf=256, K = 8 jet components; forward + residuals."""
    import space_time_pde_b200 as sp
    device, rank = ctx.device, ctx.rank
    nf, npts = 256, args.config4_points
    model = make_model(device, seed=4, nf=nf, dim=4)
    g = torch.Generator().manual_seed(400)
    grid = (torch.randn(1, 8, 8, 8, 8, CHANNELS, generator=g) * 0.5).to(device)
    q = (torch.rand(1, npts, 4, generator=torch.Generator().manual_seed(401 + rank)) * (1 - 2e-6) + 1e-6).to(device)
    layer, equations = ns4d_layer()
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
    kc = 8

    def step():
        with torch.no_grad():
            return layer(q, return_residue=True)

    ms = ctx.time(step, 1, 1)
    y, res = step()
    par = oracle_parity(model, grid, q, y, res, 24, equations=equations)
    fpt = flops_per_point(nf, 4, CHANNELS, 4, kc)
    rate = ctx.world * npts / (ms * 1e-3)
    return {"workload": f"4-d Navier-Stokes strings (3 momentum equations with Laplacians + continuity), latent 8^4 x 32, "
                        f"ImNet nf={nf} Softplus, {npts} of the configuration's 4 M query points per GPU (throughput is "
                        "flat in p: the call walks identical chunks), forward + residuals",
            "imnet_nf": nf, "dim": 4, "jet_components": kc, "dtype": f"f32 I/O, {precision} contractions",
            "value": rate, "unit": UNIT, "ms_per_step": ms, "scaling": "weak", "points_per_gpu": npts,
            "roofline": roofline_of(rate / ctx.world, fpt, peaks, {"fp16x3": 3, "fp16": 1}.get(precision, 0)),
            "parity": par}


def leg_config5(ctx, args, peaks, precision, lib):
    """BASELINE config[4]: throughput sweep, latent 32^3 x 128, ImNet nf=32, RB2; the TOTAL point count is fixed and split
    over the ranks (strong scaling), so the per-rank fixed cost of a call is what limits the small batches."""
    import space_time_pde_b200 as sp
    from space_time_pde_b200 import _lib
    device, rank, world = ctx.device, ctx.rank, ctx.world
    nf, c = 32, 128
    model = make_model(device, seed=5, nf=nf, channels=c)
    g = torch.Generator().manual_seed(500)
    grid = (torch.randn(1, 32, 32, 32, c, generator=g) * 0.3).to(device)
    layer = sp.get_rb2_pde_layer(**RB2)
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
    fpt = flops_per_point(nf, 3, c, 4, KC)
    rows, par = [], None
    for total in args.config5_points:
        p_rank = max(1, total // world)
        q = torch.rand(1, p_rank, 3, device=device) * (1 - 2e-6) + 1e-6

        def step():
            with torch.no_grad():
                return layer(q, return_residue=True)

        steps = 10 if total <= 100_000 else 4 if total <= 2_000_000 else 1
        ms = ctx.time(step, steps, 3 if total <= 2_000_000 else 1)
        rate = p_rank * world / (ms * 1e-3)
        row = {"total_points": total, "points_per_rank": p_rank, "ms_per_step": ms, "value": rate, "unit": UNIT,
               "roofline_frac": rate / world * fpt / 1e12 / peaks["tflops"]}
        if total <= 100_000:
            # what a call costs before the first point is decoded: per-call setup kernels (vertex precompute over the
            # 32^3 x 128 grid, weight packing / splitting) measured by the library's own event slots
            lib.stpde_profile_enable(1)
            _lib.profile_read()
            step()
            prof = _lib.profile_read()
            lib.stpde_profile_enable(0)
            row["kernel_ms"] = {k: round(v[0], 4) for k, v in prof.items() if v[1] > 0}
            row["fixed_cost"] = ("the per-call setup (per-vertex bias/latent table over 32768 vertices x 992 features, weight "
                                 "split: 0.7 ms) is cached across no-grad calls on the same grid / weights; what remains is "
                                 "~10 launches whose pipeline fill (TMEM allocation, barrier init, first TMA round trips) "
                                 "dominates at this size and does not shrink with the rank count")
        if par is None:
            y, res = step()
            par = oracle_parity(model, grid, q, y, res, 128, rb2_kwargs=RB2)
        rows.append(row)
        del q
    return {"workload": f"sweep: latent 32x32x32x128, ImNet nf={nf} Softplus, RB2 + continuity, forward + residuals; fixed "
                        f"TOTAL points split over {world} rank(s)",
            "imnet_nf": nf, "jet_components": KC, "dtype": f"f32 I/O, {precision} contractions", "scaling": "strong",
            "flops_per_point": fpt, "sweep": rows, "value": rows[-1]["value"], "unit": UNIT,
            "roofline": roofline_of(rows[-1]["value"] / world, fpt, peaks, {"fp16x3": 3, "fp16": 1}.get(precision, 0)),
            "roofline_hbm": hbm_roofline_of(rows[-1]["value"] / world, plane_bytes_per_point(nf, 3, KC, precision), peaks),
            "parity": par}


def leg_eval_grid(ctx, args, peaks, precision):
    """SURVEY 8(f) rank 3 - experiments/rb2d/evaluation.py:46-74,222-240: the structured evaluation grid
    linspace(eps, max - eps) of 192 x 128 x 512 points with tensor bounds maxs = [t_max, 1, 4] and a stride-0 expanded
    batch, decoded in ONE call (the reference walks 1 260 pseudo-batches of 10 000 points)."""
    import space_time_pde_b200 as sp
    device = ctx.device
    nf = 32
    nt, nz, nx = args.eval_grid
    model = make_model(device, seed=6, nf=nf)
    g = torch.Generator().manual_seed(600)
    latent = (torch.randn(1, CHANNELS, nt // 4, nz // 8, nx // 8, generator=g) * 0.5).to(device).permute(0, 2, 3, 4, 1)
    t_max = float(nt / 16)                                        # evaluation.py:225: t_max = eval_tres / nt (crop units, nt = 16)
    eps = 1e-6
    mins = torch.zeros(3, dtype=torch.float32, device=device)
    maxs = torch.tensor([t_max, 1.0, 4.0], dtype=torch.float32, device=device)
    seqs = [torch.linspace(eps, float(m) - eps, n) for m, n in zip((t_max, 1.0, 4.0), (nt, nz, nx))]
    coord = torch.stack(torch.meshgrid(*seqs, indexing="ij"), dim=-1).reshape(-1, 3).to(device)
    n_query = coord.shape[0]
    layer = sp.get_rb2_pde_layer(t_crop=t_max, z_crop=1., x_crop=4., prandtl=1., rayleigh=1e6, use_continuity=True)
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, latent, pts, mins, maxs))
    batch = coord[None].expand(1, n_query, 3)                     # stride-0 batch dimension, as evaluation.py:59

    def one_call():
        with torch.no_grad():
            return layer(batch, return_residue=True)

    ms = ctx.time(one_call, 1, 1)
    pb = 10_000
    n_pb = 64                                                     # a bounded sample of the 1 260 pseudo-batches

    def pseudo_batches():
        with torch.no_grad():
            for i in range(n_pb):
                layer(coord[i * pb:(i + 1) * pb][None].expand(1, pb, 3), return_residue=True)

    ms_pb = ctx.time(pseudo_batches, 1, 1)
    return {"workload": f"evaluation grid {nt} x {nz} x {nx} = {n_query} structured points (tie points on the clip bounds), "
                        f"bounds [0, ({t_max}, 1, 4)] as tensors, stride-0 expanded batch, latent {list(latent.shape)} "
                        f"(permuted view), ImNet nf={nf}, values + RB2 residuals, ONE call",
            "value": n_query / (ms * 1e-3), "unit": UNIT, "ms": ms, "dtype": f"f32 I/O, {precision} contractions",
            "roofline": roofline_of(n_query / (ms * 1e-3), flops_per_point(nf, 3, CHANNELS, 4, KC), peaks,
                                    {"fp16x3": 3, "fp16": 1}.get(precision, 0)),
            "roofline_hbm": hbm_roofline_of(n_query / (ms * 1e-3), plane_bytes_per_point(nf, 3, KC, precision), peaks),
            "pseudo_batch_loop": {"value": n_pb * pb / (ms_pb * 1e-3), "unit": UNIT, "batch": pb, "batches_timed": n_pb,
                                  "what": "the same decode called in the reference's 10 000-point pseudo-batches "
                                          "(evaluation.py:54-69) - per-call setup is paid every batch"}}


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
    from space_time_pde_b200 import jets as _jets
    # (i) without the two host round trips per step for the calls' status words (deferred checks, STPDE_ASYNC=1)
    os.environ["STPDE_ASYNC"] = "1"
    try:
        out["fused_async_ms"] = time_it(fused_step, 20)
        _jets.check_pending(wait=True)
    finally:
        os.environ.pop("STPDE_ASYNC", None)
    # (ii) the whole step (forward, residuals, losses, reverse sweep) recorded into ONE CUDA graph and replayed: the
    # library never allocates or synchronises, so torch.cuda.graph can capture it; status words are read afterwards
    try:
        eager_loss = float(fused_step().detach())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fused_step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            graph_loss = fused_step()
        out["cuda_graph_ms"] = time_it(graph.replay, 20)
        _jets.check_captured(clear=True)
        out["cuda_graph_loss_abs_diff"] = abs(float(graph_loss.detach()) - eager_loss)
        out["points_per_s_cuda_graph"] = B * P / (out["cuda_graph_ms"] * 1e-3)
        del graph
    except Exception as exc:
        out["cuda_graph_error"] = str(exc)[:200]
        _jets._captured.clear()
    _jets.release_workspaces()
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


def guarded(ctx, name, fn):
    """A failing optional leg (e.g. out of memory) must not take the headline down; under torchrun it must fail on every
    rank alike, so there it raises."""
    try:
        t0 = time.perf_counter()
        out = fn()
        out["leg_wall_s"] = round(time.perf_counter() - t0, 2)
        return out
    except Exception as exc:
        if ctx.world > 1:
            raise
        torch.cuda.synchronize()
        return {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("STPDE_PRECISION", "fp16x3"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--train-chunk", type=int, default=16384, help="query points per forward/backward chunk of the training leg")
    ap.add_argument("--train-workspace-mb", type=int, default=40960, help="workspace budget of the training legs (stash of one chunk)")
    ap.add_argument("--train-backward-precision", default="same", choices=["same", "fp16x3", "fp16"],
                    help="arithmetic of the reverse sweep's contractions (same = the forward's parity mode)")
    ap.add_argument("--train-steps", type=int, default=1,
                    help="timed training steps (forward + residuals + loss + fused backward [+ all-reduce]); 0 skips the leg")
    ap.add_argument("--legs", default="config2_train,config3,config4,config5,eval_grid",
                    help="comma-separated optional legs ('' = headline only)")
    ap.add_argument("--config3-chunk", type=int, default=65536, help="points (all crops) per chunk of the config-3 step")
    ap.add_argument("--graph-chunks", type=int, default=1,
                    help="1: the training legs also time their step with every chunk replayed from a CUDA graph (0: eager loop only)")
    ap.add_argument("--config4-points", type=int, default=1 << 20, help="query points per GPU of the config-4 leg (of 4 M)")
    ap.add_argument("--config5-points", type=lambda s: [int(float(x)) for x in s.split(",")], default=[10_000, 1_000_000, 64_000_000])
    ap.add_argument("--eval-grid", type=lambda s: tuple(int(x) for x in s.split("x")), default=(192, 128, 512))
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
    ctx = Ctx(device, rank, world)
    jets.set_default_precision(args.precision)
    lib = _lib.load()

    model = make_model(device)
    grid, q = synthetic_inputs(1234 + rank, device)
    layer = sp.get_rb2_pde_layer(**RB2)
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))

    def step():
        with torch.no_grad():
            return layer(q, return_residue=True)

    warm = max(3, args.warmup)
    for _ in range(warm):
        step()
    ctx.barrier()
    _lib.profile_read()                                   # reset launch counters
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    ctx.barrier()
    clocks = sampler.stop()
    launches = sum(c for _, c in _lib.profile_read().values())
    ms = ctx.max_ms(ev0.elapsed_time(ev1))
    value = world * NPTS * args.steps / (ms * 1e-3)

    # ---- parity of the timed configuration (checker, outside the timed region) ----
    y, res = step()
    parity = oracle_parity(model, grid, q, y, res, 256, rb2_kwargs=RB2)
    parity["what"] = "values + 4 RB2 residuals of the timed configuration vs the fp64 numpy oracle, first 256 points"
    del y, res

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
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    ctx.barrier()
    e2e_s = ctx.max_ms(time.perf_counter() - t0)
    e2e_value = world * NPTS * e2e_steps / e2e_s
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

    # ---- the other BASELINE configurations (short legs; every rank takes part) ----
    legs = [x for x in args.legs.split(",") if x]
    configs = {}
    if "config2_train" in legs and args.train_steps > 0:
        configs["config2_train"] = guarded(ctx, "config2_train",
                                           lambda: leg_config2_train(ctx, args, model, grid, q, layer, peaks, lib))
    del q
    torch.cuda.empty_cache()
    if "config3" in legs:
        configs["config3"] = guarded(ctx, "config3", lambda: leg_config3(ctx, args, peaks))
    if "config4" in legs:
        configs["config4"] = guarded(ctx, "config4", lambda: leg_config4(ctx, args, peaks, args.precision))
    if "config5" in legs:
        configs["config5"] = guarded(ctx, "config5", lambda: leg_config5(ctx, args, peaks, args.precision, lib))
    if "eval_grid" in legs:
        configs["eval_grid"] = guarded(ctx, "eval_grid", lambda: leg_eval_grid(ctx, args, peaks, args.precision))
    jets.release_workspaces()
    torch.cuda.empty_cache()

    if world > 1:
        dist.barrier()
    if rank == 0:
        cpu, eager = None, None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            ref = CpuReference()
            rate, secs = ref.rate(8192 if ref.kind == "port" else 4096, repeats=1)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": ref.kind,
                   "sample": f"{int(rate * secs + 0.5)} of {NPTS} points, {secs:.1f} s, {ref.what}, {cores} threads"}
            try:   # context only: the same reference algorithm as PyTorch eager ops on this GPU (true fp32, no TF32)
                eager = gpu_eager_rate(device)
            except Exception as exc:   # e.g. out of memory for the autograd tapes
                eager = {"error": str(exc)[:120]}
            if "config2_train" in configs and "error" not in configs["config2_train"]:
                try:
                    configs["config2_train"]["reference_size_step"] = reference_size_train_step(device, with_eager=True)
                except Exception as exc:
                    configs["config2_train"]["reference_size_step"] = {"error": str(exc)[:200]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args.precision), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps},
                "gpu_launches": int(launches), "roofline": roofline, "parity": parity, "cpu_baseline": cpu,
                "configs": configs, "train_step": configs.get("config2_train")}
        if eager is not None:
            line["torch_eager_gpu"] = eager
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
