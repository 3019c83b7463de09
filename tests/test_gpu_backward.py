"""GPU tests (-m gpu) of the fused reverse sweep ``stpde_jet_backward``.

Checker: the jets re-evaluated with differentiable torch ops in float64 (space_time_pde_b200/_torch_jets.py, itself pinned
against the reference algorithm's autograd in tests/test_gpu_parity.py::test_training_step_gradients and, on CPU, in
tests/test_host_logic.py) differentiated by torch.autograd for the same cotangents (gy, gjets).
Tolerance: rel-L-infinity per gradient tensor 5e-5 in the split-precision mode (the gradient is a sum over ~10^4 rows of
products that each carry the 2^-22 operand rounding, and tiny adjoints sit in the subnormal range of the lo plane), 5e-2 in the single-pass fp16 mode."""
import os

import numpy as np
import pytest
import torch

import space_time_pde_b200 as sp
from space_time_pde_b200 import _torch_jets, jets
from space_time_pde_b200.equations import JetSpec
from tests.helpers import GRAD_CASES, load_case, load_grads, pde_layer_for, record, rel_linf
from tests.test_host_logic import bounds, build_model

pytestmark = pytest.mark.gpu
BWD_TOLS = {"fp32": 5e-5, "fp16x3": 5e-5, "fp16": 5e-2}


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need CUDA (the hot path has no CPU fallback)"
    return torch.device("cuda:0")


def make_decoder(gen, d, c, o, nf, dev):
    widths = [16 * nf, 8 * nf, 4 * nf, 2 * nf, nf, o]
    D = d + c
    Ws, bs = [], []
    for l, w in enumerate(widths):
        fan_in = D if l == 0 else widths[l - 1] + (D if l < len(widths) - 1 else 0)
        bound = 1.0 / np.sqrt(fan_in)
        Ws.append(((torch.rand(w, fan_in, generator=gen) * 2 - 1) * bound).to(dev))
        bs.append(((torch.rand(w, generator=gen) * 2 - 1) * bound).to(dev))
    return Ws, bs


def reference_grads(grid, q, lo, hi, Ws, bs, act, beta, spec, gy, gj, dtype=torch.float64):
    """Autograd of the torch jets for the cotangents (gy, gj); dtype=float32 gives the plain-PyTorch-fp32 evaluation of
    the same function, whose distance from the float64 one is the noise floor the gates are pinned to."""
    g64 = grid.to(dtype).requires_grad_(True)
    W64 = [w.to(dtype).requires_grad_(True) for w in Ws]
    b64 = [b.to(dtype).requires_grad_(True) for b in bs]
    beta_t = torch.tensor(beta, dtype=dtype, device=grid.device, requires_grad=(act == "swish"))
    grads = None
    p = q.shape[1]
    step = 512                                           # bounded autograd tape
    for s in range(0, p, step):
        sl = slice(s, min(p, s + step))
        y, j = _torch_jets.query_jets(g64, q[:, sl].to(dtype), lo.to(grid.device), hi.to(grid.device), W64, b64, act, beta_t, spec)
        loss = (y * gy[:, sl].to(dtype)).sum()
        if j is not None:
            loss = loss + (j * gj[:, :, sl].to(dtype)).sum()
        leaves = [g64] + W64 + b64 + ([beta_t] if act == "swish" else [])
        gs = torch.autograd.grad(loss, leaves, allow_unused=True)
        gs = [torch.zeros_like(t) if g is None else g for g, t in zip(gs, leaves)]
        grads = gs if grads is None else [a + b for a, b in zip(grads, gs)]
    n = len(Ws)
    if dtype == torch.float64:
        reference_grads.last_gbeta = float(grads[1 + 2 * n]) if act == "swish" else None
    return grads[0], grads[1:1 + n], grads[1 + n:1 + 2 * n]


def run_case(dev, d, gshape, c, o, nf, act, first, second, p, precision, seed=0, beta=1.0, gscale=1.0):
    gen = torch.Generator().manual_seed(seed)
    Ws, bs = make_decoder(gen, d, c, o, nf, dev)
    grid = (torch.randn(1, *gshape, c, generator=gen) * 0.5).to(dev)
    q = (torch.rand(1, p, d, generator=gen) * (1 - 2e-6) + 1e-6).to(dev)
    spec = JetSpec(tuple(first), tuple(second))
    gy = (torch.randn(1, p, o, generator=gen) * gscale).to(dev)
    gj = (torch.randn(max(spec.n_jet, 1), 1, p, o, generator=gen) * gscale * 0.05).to(dev) if spec.n_jet else None
    lo, hi = jets.bounds_tensors(0., 1., d, dev)
    ggrid, gW, gB, gbeta_t = jets.raw_backward(grid, q, lo, hi, Ws, bs, act, beta, spec, precision, gy, gj)
    torch.cuda.synchronize()
    gbeta = float(gbeta_t)
    rgrid, rW, rB = reference_grads(grid, q, lo, hi, Ws, bs, act, beta, spec, gy, gj)
    errs = {"grid": rel_linf(ggrid.cpu().numpy(), rgrid.cpu().numpy())}
    if act == "swish":
        errs["beta"] = abs(gbeta - reference_grads.last_gbeta) / max(abs(reference_grads.last_gbeta), 1e-30)
    for l in range(len(Ws)):
        kh = 0 if l == 0 else Ws[l - 1].shape[0]
        a, b = gW[l].cpu().numpy(), rW[l].cpu().numpy()
        den = np.abs(b).max()
        if kh:
            errs[f"W{l}.act"] = float(np.abs(a[:, :kh] - b[:, :kh]).max() / den)
        if l < len(Ws) - 1:
            errs[f"W{l}.xrel"] = float(np.abs(a[:, kh:kh + d] - b[:, kh:kh + d]).max() / den)
            errs[f"W{l}.lat"] = float(np.abs(a[:, kh + d:] - b[:, kh + d:]).max() / den)
        errs[f"b{l}"] = rel_linf(gB[l].cpu().numpy(), rB[l].cpu().numpy())
    print(f"bwd {act} d={d} nf={nf} K={1 + len(first) + len(second)} {precision}: " +
          " ".join(f"{k}={v:.1e}" for k, v in errs.items()))
    # noise floor: plain PyTorch float32 autograd of the same jets against the float64 one (whole tensors)
    torch.backends.cuda.matmul.allow_tf32 = False
    fgrid, fW, fB = reference_grads(grid, q, lo, hi, Ws, bs, act, beta, spec, gy, gj, dtype=torch.float32)
    noise = max([rel_linf(fgrid.cpu().numpy(), rgrid.cpu().numpy())] +
                [rel_linf(a.cpu().numpy(), b.cpu().numpy()) for a, b in zip(fW + fB, rW + rB)])
    run_case.noise = noise
    test = os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0]
    record(test, "max gradient error (fused CUDA sweep vs float64 autograd)", max(errs.values()), BWD_TOLS[precision])
    record(test, "torch float32 autograd vs float64 autograd (noise floor)", noise, 0.0)
    return errs


RB2 = ((0, 1, 2), ((1, 1), (2, 2)))
PRECS = [p for p in os.environ.get("STPDE_TEST_PRECISIONS", "fp16x3,fp16").split(",") if p in ("fp16x3", "fp16")]


@pytest.mark.parametrize("precision", PRECS)
@pytest.mark.parametrize("act", ["softplus", "tanh", "swish", "elu"])
def test_backward_rb2_spec_smooth_activations(act, precision, dev):
    errs = run_case(dev, 3, (3, 4, 5), 16, 4, 8, act, *RB2, p=2048, precision=precision, seed=3, beta=1.3)
    assert max(errs.values()) < BWD_TOLS[precision], errs


@pytest.mark.parametrize("act", ["relu", "leakyrelu"])
def test_backward_kinked_activations(act, dev):
    # sigma'' = 0: only the value / first-order paths carry gradient; a pre-activation within rounding distance of 0
    # flips sigma' at isolated rows, so the gate is looser (same convention as the forward tests)
    errs = run_case(dev, 3, (3, 4, 5), 16, 4, 8, act, *RB2, p=2048, precision="fp16x3", seed=4)
    assert max(errs.values()) < 2e-3, errs


@pytest.mark.parametrize("case", [
    (1, (9,), 8, 2, 8, (0,), ((0, 0),)),                         # d = 1, K = 3 (80-row tiles)
    (2, (5, 6), 8, 3, 8, (0, 1), ((0, 1),)),                     # d = 2, mixed partial, K = 4
    (3, (3, 4, 5), 16, 4, 8, (), ()),                            # values only, K = 1
    (3, (3, 4, 5), 16, 4, 8, (0, 1, 2), ()),                     # gradient only, K = 4
    (3, (3, 4, 5), 16, 4, 8, (0, 1, 2), tuple((a, b) for a in range(3) for b in range(a, 3))),   # full Hessian, K = 10
    (4, (3, 3, 3, 3), 8, 4, 8, (0, 1, 2, 3), ((0, 0), (1, 1), (2, 2))),                           # d = 4, K = 8
])
def test_backward_dims_and_jet_specs(case, dev):
    d, gshape, c, o, nf, first, second = case
    errs = run_case(dev, d, gshape, c, o, nf, "softplus", first, second, p=1024, precision="fp16x3", seed=5)
    assert max(errs.values()) < BWD_TOLS["fp16x3"], errs


@pytest.mark.parametrize("nf", [32, 64])
def test_backward_wide_decoders(nf, dev):
    # several 256-feature tiles per layer, 256-wide wgrad N tiles, multiple K slices
    errs = run_case(dev, 3, (4, 6, 6), 32, 4, nf, "softplus", *RB2, p=1024, precision="fp16x3", seed=6)
    assert max(errs.values()) < BWD_TOLS["fp16x3"], errs


def test_backward_tiny_cotangents_are_rescaled(dev):
    # mean-reduced losses over 10^6 points give |gy| ~ 1e-7: the adjoint scale keeps them inside the fp16 planes
    errs = run_case(dev, 3, (3, 4, 5), 16, 4, 8, "softplus", *RB2, p=1024, precision="fp16x3", seed=7, gscale=1e-7)
    assert max(errs.values()) < BWD_TOLS["fp16x3"], errs


def test_backward_multi_chunk_matches_single_chunk(dev, monkeypatch):
    errs1 = run_case(dev, 3, (3, 4, 5), 16, 4, 8, "tanh", *RB2, p=3000, precision="fp16x3", seed=8)
    jets.release_workspaces()
    monkeypatch.setenv("STPDE_WORKSPACE_MB", "64")          # forces several chunks (and a ragged last one)
    errs2 = run_case(dev, 3, (3, 4, 5), 16, 4, 8, "tanh", *RB2, p=3000, precision="fp16x3", seed=8)
    jets.release_workspaces()
    assert max(errs1.values()) < BWD_TOLS["fp16x3"], errs1
    assert max(errs2.values()) < BWD_TOLS["fp16x3"], errs2


def test_training_step_fused_vs_torch_route(dev, monkeypatch):
    """PDELayer + loss.backward(): the fused CUDA backward against the autograd re-evaluation on the same inputs."""
    torch.manual_seed(0)
    model = sp.ImNet(dim=3, in_features=16, out_features=4, nf=16, activation=sp.NONLINEARITIES["softplus"]).to(dev)
    grid0 = (torch.randn(2, 4, 6, 5, 16) * 0.5).to(dev)
    q = torch.rand(2, 1500, 3, device=dev)
    layer = sp.get_rb2_pde_layer(t_crop=2., z_crop=1., x_crop=1., use_continuity=True)

    def grads(route):
        monkeypatch.setenv("STPDE_BACKWARD", route)
        grid = grid0.clone().requires_grad_(True)
        model.zero_grad()
        layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
        y, res = layer(q)
        loss = y.abs().mean() + 0.0125 * torch.stack(list(res.values())).abs().mean()
        loss.backward()
        return [grid.grad.clone()] + [p.grad.clone() for p in model.parameters()]

    fused, ref = grads("fused"), grads("torch")
    for i, (a, b) in enumerate(zip(fused, ref)):
        assert rel_linf(a.cpu().numpy(), b.cpu().numpy()) < 1e-4, i


def test_training_step_replays_from_a_cuda_graph(dev):
    """The library never allocates, synchronises or reads device values on the host, so a whole training step
    (forward, residuals, loss, fused reverse sweep) can be recorded with torch.cuda.graph and replayed: the
    reference-size step is launch-latency bound.  Status words are read after the replay (``check_captured``)."""
    torch.manual_seed(0)
    model = sp.ImNet(dim=3, in_features=16, out_features=4, nf=16, activation=sp.NONLINEARITIES["softplus"]).to(dev)
    grid = (torch.randn(2, 4, 6, 5, 16) * 0.5).to(dev).requires_grad_(True)
    q = torch.rand(2, 700, 3, device=dev)
    layer = sp.get_rb2_pde_layer(t_crop=2., z_crop=1., x_crop=1., use_continuity=True)
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))

    def step():
        model.zero_grad(set_to_none=True)
        grid.grad = None
        y, res = layer(q)
        loss = y.abs().mean() + 0.0125 * torch.stack(list(res.values())).abs().mean()
        loss.backward()
        return loss

    eager_loss = float(step().detach())
    eager = [grid.grad.clone()] + [p.grad.clone() for p in model.parameters()]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        loss = step()
    captured = [grid.grad] + [p.grad for p in model.parameters()]
    try:
        with torch.no_grad():
            q_new = torch.rand(2, 700, 3, device=dev)
            q_old = q.clone()
            q.copy_(q_new)                 # static input buffer: new points, same graph
        graph.replay()
        jets.check_captured()
        assert abs(float(loss.detach()) - eager_loss) > 0          # the replay really used the new points
        q.copy_(q_old)
        graph.replay()
        jets.check_captured()
        assert abs(float(loss.detach()) - eager_loss) < 1e-6 * max(1.0, abs(eager_loss))
        for i, (a, b) in enumerate(zip(captured, eager)):
            assert rel_linf(a.cpu().numpy(), b.cpu().numpy()) < 1e-5, i      # (atomics: summation order differs)
        assert len(jets._captured) == 2                   # forward + backward status words, still on the device
    finally:
        jets._captured.clear()
        del graph
        jets.release_workspaces()


def test_backward_fallback_routes(dev):
    """Gradients w.r.t. the query points still work (torch route), and so does a decoder the kernel does not cover."""
    torch.manual_seed(1)
    model = sp.ImNet(dim=3, in_features=8, out_features=2, nf=8, activation=sp.NONLINEARITIES["tanh"]).to(dev)
    grid = (torch.randn(1, 3, 4, 5, 8) * 0.5).to(dev)
    q = torch.rand(1, 64, 3, device=dev).requires_grad_(True)
    y = sp.query_local_implicit_grid(model, grid, q, 0., 1.)
    y.sum().backward()
    assert q.grad is not None and torch.isfinite(q.grad).all()


def test_training_forward_stash_is_reused_and_invalidated(dev):
    """stpde_jet_forward_train keeps the chunk's planes; the backward with the current token skips the recompute and
    gives the same gradients as the recomputing backward; a later training forward invalidates the token."""
    gen = torch.Generator().manual_seed(9)
    d, c, o, nf, p = 3, 16, 4, 8, 1000
    Ws, bs = make_decoder(gen, d, c, o, nf, dev)
    grid = (torch.randn(1, 3, 4, 5, c, generator=gen) * 0.5).to(dev)
    q = (torch.rand(1, p, d, generator=gen) * (1 - 2e-6) + 1e-6).to(dev)
    spec = JetSpec(*RB2)
    gy = torch.randn(1, p, o, generator=gen).to(dev)
    gj = (torch.randn(spec.n_jet, 1, p, o, generator=gen) * 0.05).to(dev)
    lo, hi = jets.bounds_tensors(0., 1., d, dev)
    y0, j0 = jets.raw_forward(grid, q, lo, hi, Ws, bs, "tanh", 1.0, spec, "fp16x3")
    tok = []
    y1, j1 = jets.raw_forward(grid, q, lo, hi, Ws, bs, "tanh", 1.0, spec, "fp16x3", stash_out=tok)
    assert len(tok) == 1 and tok[0] > 0
    assert rel_linf(y1.cpu().numpy(), y0.cpu().numpy()) < 2e-6          # same arithmetic, pair kernel for every layer
    assert rel_linf(j1.cpu().numpy(), j0.cpu().numpy()) < 2e-6
    ref = jets.raw_backward(grid, q, lo, hi, Ws, bs, "tanh", 1.0, spec, "fp16x3", gy, gj)
    reuse = jets.raw_backward(grid, q, lo, hi, Ws, bs, "tanh", 1.0, spec, "fp16x3", gy, gj, stash_token=tok[0])
    again = jets.raw_backward(grid, q, lo, hi, Ws, bs, "tanh", 1.0, spec, "fp16x3", gy, gj, stash_token=tok[0])
    for a, b, c_ in zip([ref[0]] + ref[1] + ref[2], [reuse[0]] + reuse[1] + reuse[2], [again[0]] + again[1] + again[2]):
        assert rel_linf(b.cpu().numpy(), a.cpu().numpy()) < 2e-6         # atomics reorder the sums, nothing else differs
        assert rel_linf(c_.cpu().numpy(), a.cpu().numpy()) < 2e-6
    # another training forward (different points) replaces the stash: the old token must not be honoured
    q2 = (torch.rand(1, p, d, generator=gen) * (1 - 2e-6) + 1e-6).to(dev)
    tok2 = []
    jets.raw_forward(grid, q2, lo, hi, Ws, bs, "tanh", 1.0, spec, "fp16x3", stash_out=tok2)
    stale = jets.raw_backward(grid, q, lo, hi, Ws, bs, "tanh", 1.0, spec, "fp16x3", gy, gj, stash_token=tok[0])
    assert rel_linf(stale[0].cpu().numpy(), ref[0].cpu().numpy()) < 2e-6
    assert tok2[0] != tok[0]


def test_single_pass_stash_keeps_fp16_preactivations_consistent(dev):
    """Single-pass training: the forward leaves fp16 pre-activation planes in the stash and tells the reverse sweep so
    (desc.reserved[2]).  Every stash combination must give the gradients of its own recompute route: fp16 forward + fp16
    sweep (fp16 planes), fp16x3 forward + fp16 sweep (fp32 planes read by a single-pass sweep), and a parity-mode sweep
    must NOT reuse a single-pass stash."""
    gen = torch.Generator().manual_seed(21)
    d, c, o, nf, p = 3, 16, 4, 32, 1500
    Ws, bs = make_decoder(gen, d, c, o, nf, dev)
    grid = (torch.randn(1, 3, 4, 5, c, generator=gen) * 0.5).to(dev)
    q = (torch.rand(1, p, d, generator=gen) * (1 - 2e-6) + 1e-6).to(dev)
    spec = JetSpec(*RB2)
    gy = torch.randn(1, p, o, generator=gen).to(dev)
    gj = (torch.randn(spec.n_jet, 1, p, o, generator=gen) * 0.05).to(dev)
    lo, hi = jets.bounds_tensors(0., 1., d, dev)
    flat = lambda r: [r[0]] + r[1] + r[2]
    ref16 = jets.raw_backward(grid, q, lo, hi, Ws, bs, "softplus", 1.0, spec, "fp16", gy, gj)          # recompute, fp16 planes
    ref48 = jets.raw_backward(grid, q, lo, hi, Ws, bs, "softplus", 1.0, spec, "fp16x3", gy, gj)
    for fwd_prec, bwd_prec, ref, tol in (("fp16", "fp16", ref16, 2e-6), ("fp16x3", "fp16", ref16, 5e-2),
                                         ("fp16", "fp16x3", ref48, 2e-6)):
        tok = []
        jets.raw_forward(grid, q, lo, hi, Ws, bs, "softplus", 1.0, spec, fwd_prec, stash_out=tok)
        got = jets.raw_backward(grid, q, lo, hi, Ws, bs, "softplus", 1.0, spec, bwd_prec, gy, gj, stash_token=tok[0])
        for a, b in zip(flat(ref), flat(got)):
            assert rel_linf(b.cpu().numpy(), a.cpu().numpy()) < tol, (fwd_prec, bwd_prec)
    # and the single-pass gradients themselves stay inside the mode's gate against the parity-mode sweep
    for a, b in zip(flat(ref48), flat(ref16)):
        assert rel_linf(b.cpu().numpy(), a.cpu().numpy()) < BWD_TOLS["fp16"]


def test_training_setup_cache_follows_weight_updates(dev, monkeypatch):
    """Chunks of one training step reuse the split weights / per-vertex table of the previous training forward in the stash
    (same tensors, same versions: desc.reserved[1]); an in-place optimizer update bumps the versions and the next forward
    rebuilds them.  Gradients with the cache equal the gradients without it (STPDE_SETUP_CACHE=0)."""
    torch.manual_seed(31)
    model = sp.ImNet(dim=3, in_features=16, out_features=4, nf=32, activation=sp.NONLINEARITIES["softplus"]).to(dev)
    grid = (torch.randn(2, 3, 4, 5, 16) * 0.5).to(dev).requires_grad_(True)
    q = torch.rand(2, 1536, 3, device=dev) * (1 - 2e-6) + 1e-6
    layer = sp.get_rb2_pde_layer(t_crop=2., z_crop=1., x_crop=1., use_continuity=True)
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
    opt = torch.optim.SGD(model.parameters(), lr=0.05)

    def two_steps():
        out = []
        for _ in range(2):
            opt.zero_grad(set_to_none=True)
            grid.grad = None
            for s0 in range(0, 1536, 512):                       # three chunks per step
                y, sums, counts = layer.loss_sums(q[:, s0:s0 + 512], None, "l1")
                (sums[0] / counts[0] + 0.0125 * sums[1] / counts[1]).backward()
            out.append([grid.grad.clone()] + [p_.grad.clone() for p_ in model.parameters()])
            opt.step()                                           # in-place: versions change, the cache must not survive
        return out

    state = {k: v.clone() for k, v in model.state_dict().items()}
    with_cache = two_steps()
    model.load_state_dict(state)
    monkeypatch.setenv("STPDE_SETUP_CACHE", "0")
    without = two_steps()
    for step_a, step_b in zip(with_cache, without):
        for a, b in zip(step_a, step_b):
            assert rel_linf(a.cpu().numpy(), b.cpu().numpy()) < 5e-6
    # the second step really used the updated weights
    assert rel_linf(with_cache[1][1].cpu().numpy(), with_cache[0][1].cpu().numpy()) > 1e-4


def test_chunks_replayed_from_cuda_graphs_accumulate_like_the_eager_loop(dev):
    """bench.graphed_chunk_step: every chunk of a step replayed from a CUDA graph (graph A rebuilds the call-invariant
    setup, graph B reuses it) accumulates the same gradients and loss sums as the eager chunk loop - also after an
    in-place weight update between two steps (the first chunk's graph re-splits the weights)."""
    import bench
    torch.manual_seed(41)
    model = sp.ImNet(dim=3, in_features=16, out_features=4, nf=32, activation=sp.NONLINEARITIES["softplus"]).to(dev)
    grid = (torch.randn(2, 3, 4, 5, 16) * 0.5).to(dev).requires_grad_(True)
    q = torch.rand(2, 1536, 3, device=dev) * (1 - 2e-6) + 1e-6
    target = torch.randn(2, 1536, 4, device=dev)
    layer = sp.get_rb2_pde_layer(t_crop=2., z_crop=1., x_crop=1., use_continuity=True)
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
    params = [grid] + list(model.parameters())
    loss_of = lambda sums: sums[0] * 1e-3 + 0.0125 * sums[1] * 1e-3

    def eager():
        for p_ in params:
            p_.grad = None
        tot = torch.zeros(2, device=dev)
        for s0 in range(0, 1536, 512):
            y, sums, _ = layer.loss_sums(q[:, s0:s0 + 512], target[:, s0:s0 + 512], "l1")
            loss_of(sums).backward()
            tot += torch.stack([sums[0].detach(), sums[1].detach()])
        return tot, [p_.grad.clone() for p_ in params]

    for round_ in range(2):
        tot_e, g_e = eager()
        gstep = bench.graphed_chunk_step(layer, params, q, target, 512, loss_of, dev)
        for _ in range(2):                                       # replaying twice gives the same step twice
            reg, pde = gstep()
            assert rel_linf(torch.stack([reg, pde]).cpu().numpy(), tot_e.cpu().numpy()) < 1e-6
            for a, b in zip(g_e, [p_.grad for p_ in params]):
                assert rel_linf(b.cpu().numpy(), a.cpu().numpy()) < 5e-6
        if round_ == 0:
            # in-place update (as an optimizer does): the captured graphs read the weights through the same pointers
            with torch.no_grad():
                for p_ in model.parameters():
                    p_.add_(0.01 * torch.randn_like(p_))
            tot_u, g_u = eager()
            reg, pde = gstep()
            assert rel_linf(torch.stack([reg, pde]).cpu().numpy(), tot_u.cpu().numpy()) < 1e-6
            for a, b in zip(g_u, [p_.grad for p_ in params]):
                assert rel_linf(b.cpu().numpy(), a.cpu().numpy()) < 5e-6
        del gstep
        jets.check_captured(clear=True)
    for p_ in params:
        p_.grad = None


def test_chunked_training_accumulates_like_one_batch(dev):
    """Walking the batch in chunks (each with its own stash) accumulates the same .grad as one big backward."""
    torch.manual_seed(2)
    model = sp.ImNet(dim=3, in_features=16, out_features=4, nf=8, activation=sp.NONLINEARITIES["softplus"]).to(dev)
    grid = (torch.randn(1, 4, 6, 5, 16) * 0.5).to(dev).requires_grad_(True)
    q = torch.rand(1, 3000, 3, device=dev)
    layer = sp.get_rb2_pde_layer(t_crop=2., z_crop=1., x_crop=1., use_continuity=True)
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))

    def run(chunk):
        model.zero_grad()
        grid.grad = None
        for s0 in range(0, q.shape[1], chunk):
            y, res = layer(q[:, s0:s0 + chunk])
            (y.abs().sum() + 0.0125 * torch.stack(list(res.values())).abs().sum()).backward()
        return [grid.grad.clone()] + [p_.grad.clone() for p_ in model.parameters()]

    whole, parts = run(3000), run(700)
    for a, b in zip(parts, whole):
        assert rel_linf(a.cpu().numpy(), b.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("name", GRAD_CASES)
def test_loss_backward_matches_reference_golden_gradients(name, dev):
    """PDELayer + training-style loss + loss.backward() on the GPU (fused forward, fused reverse sweep) against the
    gradients of the REAL reference (float64, tests/golden/make_golden_grads.py).  Gate per tensor: rel-L-infinity
    max(1e-5, 2 * err(reference float32, reference float64)) with the reference's own float32 noise stored in the
    fixture (``noise_*``, measured 4e-7 .. 8e-7 on these cases, so the gate is 1e-5); the golden losses are means, so
    the cotangents are ~1e-3 .. 1e-2 and exercise the adjoint rescaling."""
    c, g = load_case(name), load_grads(name)
    o = c["Ws"][5].shape[0]
    model = build_model(c, o).to(dev)
    grid = torch.tensor(c["grid"]).to(dev).requires_grad_(True)
    q = torch.tensor(c["q"]).to(dev)
    xmin, xmax = bounds(c)
    layer = pde_layer_for(sp, name, c)
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, xmin, xmax))
    y, res = layer(q, return_residue=True)
    loss = y.abs().mean() + 0.0125 * torch.stack(list(res.values())).abs().mean()
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    errs = {"grid": rel_linf(grid.grad.cpu().numpy(), g["g_grid"])}
    for i in range(6):
        errs[f"W{i}"] = rel_linf(model.fc[i].weight.grad.cpu().numpy(), g[f"g_W{i}"])
        errs[f"b{i}"] = rel_linf(model.fc[i].bias.grad.cpu().numpy(), g[f"g_b{i}"])
    if "g_beta" in g:   # learnable Swish beta (reference src/nonlinearities.py:5-13): gradient from the same fused sweep
        errs["beta"] = rel_linf(model.activ.beta.grad.cpu().numpy().reshape(1), g["g_beta"])
    print(name, " ".join(f"{k}={v:.1e}" for k, v in errs.items()))
    for k, v in errs.items():
        key = "noise_" + ("grid" if k == "grid" else k)
        gate = max(1e-5, 2 * float(g[key])) if key in g else 1e-5
        assert record("golden_gradients:" + name, k, v, gate) < gate, (k, errs)


def test_single_pass_backward_option(dev):
    """STPDE_BACKWARD_PRECISION=fp16: mixed-precision style reverse sweep (one fp16 pass) on a parity-mode forward."""
    errs = run_case(dev, 3, (3, 4, 5), 16, 4, 8, "softplus", *RB2, p=2048, precision="fp16", seed=3)
    assert max(errs.values()) < BWD_TOLS["fp16"], errs
    jets.set_backward_precision("fp16")
    try:
        torch.manual_seed(3)
        model = sp.ImNet(dim=3, in_features=16, out_features=4, nf=8, activation=sp.NONLINEARITIES["softplus"]).to(dev)
        grid = (torch.randn(1, 4, 6, 5, 16) * 0.5).to(dev).requires_grad_(True)
        q = torch.rand(1, 1000, 3, device=dev)
        layer = sp.get_rb2_pde_layer(t_crop=2., z_crop=1., x_crop=1., use_continuity=True)
        layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))

        def grads():
            model.zero_grad()
            grid.grad = None
            y, res = layer(q)
            (y.abs().mean() + 0.0125 * torch.stack(list(res.values())).abs().mean()).backward()
            return [grid.grad.clone()] + [p_.grad.clone() for p_ in model.parameters()]

        low = grads()
        jets.set_backward_precision("same")
        full = grads()
    finally:
        jets.set_backward_precision("same")
    for a, b in zip(low, full):
        assert rel_linf(a.cpu().numpy(), b.cpu().numpy()) < 5e-2


def test_residual_program_backward_matches_torch_route(dev, monkeypatch):
    """loss.backward() with the residual arithmetic in CUDA (stpde_residuals + stpde_residuals_backward) against the
    same step with the lambdified torch arithmetic (STPDE_RESIDUALS=torch)."""
    torch.manual_seed(5)
    model = sp.ImNet(dim=3, in_features=16, out_features=4, nf=8, activation=sp.NONLINEARITIES["softplus"]).to(dev)
    grid0 = (torch.randn(2, 4, 6, 5, 16) * 0.5).to(dev)
    q = torch.rand(2, 700, 3, device=dev)
    layer = sp.get_rb2_pde_layer(mean=[0.1, -0.2, 0.05, 0.3], std=[1.1, 0.9, 1.3, 0.7], t_crop=2., z_crop=1., x_crop=2.,
                                 use_continuity=True)

    def grads(route):
        monkeypatch.setenv("STPDE_RESIDUALS", route)
        grid = grid0.clone().requires_grad_(True)
        model.zero_grad()
        layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
        y, res = layer(q, return_residue=True)
        loss = (y ** 2).mean() + 0.0125 * (torch.stack(list(res.values())) ** 2).mean()
        loss.backward()
        return [loss.detach(), grid.grad.clone()] + [p_.grad.clone() for p_ in model.parameters()]

    kern, ref = grads("kernel"), grads("torch")
    for i, (a, b) in enumerate(zip(kern, ref)):
        assert rel_linf(a.cpu().numpy(), b.cpu().numpy()) < 1e-5, i


def test_encoder_gradients_through_permuted_latent_grid(dev, monkeypatch):
    """The reference feeds the decoder a PERMUTED (non-contiguous) view of the UNet3d output (train.py:58-60).  A small
    conv encoder stands in for it: its weight gradients through the fused path must match the torch route."""
    torch.manual_seed(11)
    enc = torch.nn.Conv3d(4, 16, 3, padding=1).to(dev)
    model = sp.ImNet(dim=3, in_features=16, out_features=4, nf=8, activation=sp.NONLINEARITIES["softplus"]).to(dev)
    lres = torch.randn(2, 4, 4, 6, 5, device=dev)
    q = torch.rand(2, 900, 3, device=dev)
    layer = sp.get_rb2_pde_layer(t_crop=2., z_crop=1., x_crop=1., use_continuity=True)

    def grads(route):
        monkeypatch.setenv("STPDE_BACKWARD", route)
        enc.zero_grad()
        model.zero_grad()
        latent = enc(lres).permute(0, 2, 3, 4, 1)                  # [b, T, Z, X, C], non-contiguous
        assert not latent.is_contiguous()
        layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, latent, pts, 0., 1.))
        y, res = layer(q, return_residue=True)
        (y.abs().mean() + 0.0125 * torch.stack(list(res.values())).abs().mean()).backward()
        return [enc.weight.grad.clone(), enc.bias.grad.clone(), model.fc[0].weight.grad.clone()]

    fused, ref = grads("fused"), grads("torch")
    for a, b in zip(fused, ref):
        assert rel_linf(a.cpu().numpy(), b.cpu().numpy()) < 1e-4


def test_swish_beta_gradient_through_module(dev, monkeypatch):
    """The learnable beta of the reference's Swish (src/nonlinearities.py:5-13) gets its gradient from the fused sweep."""
    torch.manual_seed(12)
    model = sp.ImNet(dim=3, in_features=16, out_features=4, nf=8, activation=sp.NONLINEARITIES["swish"]).to(dev)
    with torch.no_grad():
        model.activ.beta.fill_(1.2)
    grid0 = (torch.randn(1, 4, 6, 5, 16) * 0.5).to(dev)
    q = torch.rand(1, 1200, 3, device=dev)
    layer = sp.get_rb2_pde_layer(t_crop=2., z_crop=1., x_crop=1., use_continuity=True)

    def grads(route):
        monkeypatch.setenv("STPDE_BACKWARD", route)
        grid = grid0.clone().requires_grad_(True)
        model.zero_grad()
        layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
        y, res = layer(q, return_residue=True)
        (y.abs().mean() + 0.0125 * torch.stack(list(res.values())).abs().mean()).backward()
        return [model.activ.beta.grad.clone().reshape(1), grid.grad.clone(), model.fc[2].weight.grad.clone()]

    fused, ref = grads("fused"), grads("torch")
    assert fused[0].abs().item() > 0
    for a, b in zip(fused, ref):
        assert rel_linf(a.cpu().numpy(), b.cpu().numpy()) < 1e-4


def test_backward_empty_and_tiny_batches(dev):
    """Edge cases of the point dimension: p = 0 (all gradients exactly zero) and p = 1."""
    gen = torch.Generator().manual_seed(13)
    d, c, o, nf = 3, 16, 4, 8
    Ws, bs = make_decoder(gen, d, c, o, nf, dev)
    grid = (torch.randn(1, 3, 4, 5, c, generator=gen) * 0.5).to(dev)
    spec = JetSpec(*RB2)
    lo, hi = jets.bounds_tensors(0., 1., d, dev)
    q0 = torch.empty(1, 0, d, device=dev)
    ggrid, gW, gB, _ = jets.raw_backward(grid, q0, lo, hi, Ws, bs, "softplus", 1.0, spec, "fp16x3",
                                      torch.empty(1, 0, o, device=dev), torch.empty(spec.n_jet, 1, 0, o, device=dev))
    assert ggrid.abs().max() == 0 and all(g.abs().max() == 0 for g in gW + gB)
    # the public API with an empty batch: shapes only, through forward, residual programs and backward
    model = sp.ImNet(dim=3, in_features=c, out_features=o, nf=nf, activation=sp.NONLINEARITIES["softplus"]).to(dev)
    g2 = grid.clone().requires_grad_(True)
    layer = sp.get_rb2_pde_layer(t_crop=2., z_crop=1., x_crop=1., use_continuity=True)
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, g2, pts, 0., 1.))
    y, res = layer(q0)
    assert y.shape == (1, 0, o) and all(v.shape == (1, 0, 1) for v in res.values())
    (y.sum() + torch.stack(list(res.values())).sum()).backward()
    assert g2.grad is not None and g2.grad.abs().max() == 0
    errs = run_case(dev, d, (3, 4, 5), c, o, nf, "softplus", *RB2, p=1, precision="fp16x3", seed=14)
    assert max(errs.values()) < BWD_TOLS["fp16x3"], errs


@pytest.mark.parametrize("loss_type", ["l1", "l2", "huber"])
def test_fused_loss_sums_match_the_reference_training_losses(loss_type, dev, monkeypatch):
    """PDELayer.loss_sums (residual programs + loss reductions in one kernel) against the reference's
    formulation of the same step (experiments/rb2d/train.py:70-75: loss_func(pred, target), loss_func(stack(residues), 0))
    evaluated with torch ops on the SAME fused forward: sums / counts == the mean losses to 1e-6, and the gradients of
    alpha_reg * reg + alpha_pde * pde w.r.t. the latent grid and the decoder agree."""
    F = {"l1": torch.nn.functional.l1_loss, "l2": torch.nn.functional.mse_loss, "huber": torch.nn.functional.smooth_l1_loss}[loss_type]
    torch.manual_seed(21)
    model = sp.ImNet(dim=3, in_features=16, out_features=4, nf=8, activation=sp.NONLINEARITIES["softplus"]).to(dev)
    grid0 = (torch.randn(2, 4, 6, 5, 16) * 0.5).to(dev)
    q = torch.rand(2, 3000, 3, device=dev)
    target = torch.randn(2, 3000, 4, device=dev) * 2.0            # |y - target| on both sides of 1: both huber branches
    layer = sp.get_rb2_pde_layer(t_crop=2., z_crop=1., x_crop=1., use_continuity=True)

    def run(fused):
        grid = grid0.clone().requires_grad_(True)
        model.zero_grad()
        layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
        if fused:
            y, sums, counts = layer.loss_sums(q, target, loss_type)
            reg, pde = sums[0] / counts[0], sums[1] / counts[1]
        else:
            y, res = layer(q, return_residue=True)
            reg = F(y, target)
            stacked = torch.stack([d for d in res.values()], dim=0)
            pde = F(stacked, torch.zeros_like(stacked))
        (1.0 * reg + 0.0125 * pde).backward()
        return float(reg), float(pde), [grid.grad.clone()] + [p.grad.clone() for p in model.parameters()]

    reg_f, pde_f, g_f = run(True)
    reg_t, pde_t, g_t = run(False)
    assert abs(reg_f - reg_t) < 1e-6 * abs(reg_t) and abs(pde_f - pde_t) < 1e-6 * abs(pde_t)
    for a, b in zip(g_f, g_t):
        assert rel_linf(a.cpu().numpy(), b.cpu().numpy()) < 1e-5
    # no target (zeros) and no-grad evaluation
    with torch.no_grad():
        y, sums, counts = layer.loss_sums(q, None, loss_type)
        y2, res = layer(q)
        st = torch.stack(list(res.values()))
        assert abs(float(sums[1] / counts[1]) - float(F(st, torch.zeros_like(st)))) < 1e-6 * float(F(st, torch.zeros_like(st)))
        assert abs(float(sums[0] / counts[0]) - float(F(y2, torch.zeros_like(y2)))) < 1e-6 * float(F(y2, torch.zeros_like(y2)))
