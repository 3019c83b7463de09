"""Fused decode + jets: the torch-facing wrapper around ``stpde_jet_forward`` (include/stpde.h).

``fused_query`` replaces, in ONE kernel sequence per chunk of points,
  reference src/local_implicit_grid.py:47-61  (corner gather, ImNet x 2^d, blend) and
  reference src/pde.py:8-9                    (every ``torch.autograd.grad`` the equations trigger).
"""
from __future__ import annotations

import contextlib
import ctypes
import itertools
import os
import threading
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .equations import JetSpec

_state = threading.local()
_test_backend = None          # tests only: CPU stand-in for the kernel (see _torch_jets.py)
# "fp16x3": tcgen05 tensor cores with fp16 hi/lo split operands (fp32 parity, the default);
# "fp32": FP32 FFMA on the CUDA cores (reference-exact arithmetic); "fp16": single-pass tensor cores (relaxed).
DEFAULT_PRECISION = os.environ.get("STPDE_PRECISION", "fp16x3")


# STPDE_ASYNC=1 keeps the host from waiting for every call's status word.  The word is copied to pinned host memory behind
# the call instead, and every later call (or ``check_pending()``) raises for the calls that have finished since - errors
# are delayed, never dropped.
_pending: list = []


def _async_mode() -> bool:
    return os.environ.get("STPDE_ASYNC", "0") == "1" or getattr(_state, "deferred", 0) > 0


@contextlib.contextmanager
def deferred_checks():
    """No call inside the block waits for its status word (as under ``STPDE_ASYNC=1``); leaving the block synchronises on
    all of them and raises for the first error.  For loops over chunks of one step: the host keeps launching while the
    device works instead of idling through two round trips per chunk.  (The backward cannot repeat itself with more
    adjoint headroom in this mode: an fp16 overflow of an adjoint is reported, not repaired.)"""
    _state.deferred = getattr(_state, "deferred", 0) + 1
    try:
        yield
    except BaseException:
        _state.deferred -= 1
        raise
    else:
        _state.deferred -= 1
        if _state.deferred == 0:
            check_pending(wait=True)


def _raise_for_status(flags: int, where: str) -> None:
    if flags & 1:
        raise IndexError("query point addressed a cell outside the latent grid "
                         "(reference regular_nd_grid_interpolation.py:52 ignores xmin; use xmin = 0)" + where)
    if flags & 2:
        if "backward" in where:
            raise _lib.StpdeError(-6, "an adjoint left the fp16 range of the split-precision tensor-core backward; "
                                      "set STPDE_BACKWARD=torch to use the autograd re-evaluation" + where)
        raise _lib.StpdeError(-6, "activation left the fp16 range of the split-precision tensor-core path; "
                                  "use precision='fp32'" + where)


def _defer_status(status: torch.Tensor, where: str) -> None:
    host = torch.empty(1, dtype=torch.int32, pin_memory=True)
    host.copy_(status, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(status.device))
    _pending.append((ev, host, where))


def check_pending(wait: bool = False) -> None:
    """Raise for every asynchronous call whose status word has arrived (``wait=True``: synchronise on all of them)."""
    keep = []
    err = None
    for ev, host, where in _pending:
        if wait:
            ev.synchronize()
        if ev.query():
            if err is None and int(host[0]) & 3:
                err = (int(host[0]), where)
        else:
            keep.append((ev, host, where))
    _pending[:] = keep
    if err is not None:
        _raise_for_status(err[0], f" [reported late: {err[1]} call under STPDE_ASYNC=1]")


# CUDA-graph capture (torch.cuda.graph around a whole training / inference step): nothing may wait for the device while
# the stream is capturing, so the calls' status words stay on the device; ``check_captured()`` reads them after a replay.
# The backward cannot retry with more adjoint headroom inside a graph either - an overflow is reported, not repaired.
_captured: list = []


def _capturing(device: torch.device) -> bool:
    return device.type == "cuda" and torch.cuda.is_current_stream_capturing()


def check_captured(clear: bool = False) -> None:
    """Raise if a call recorded into a CUDA graph flagged an error during the latest replay (synchronises)."""
    flags = [(int(t.item()), where) for t, where in _captured]
    if clear:
        _captured.clear()
    for f, where in flags:
        if f & 3:
            _raise_for_status(f, f" [{where} call recorded in a CUDA graph]")


def set_test_backend(fn) -> None:
    """Install a stand-in jet provider (tests of the host logic on machines without a GPU)."""
    global _test_backend
    _test_backend = fn


# Arithmetic of the reverse sweep's contractions: "same" follows the forward precision (fp32 / fp16x3 -> 3-pass
# split, the parity mode); "fp16" runs dgrad / wgrad as single fp16 passes (mixed-precision training: gradients
# accurate to ~1e-3, 3x fewer tensor-core MMAs).
BACKWARD_PRECISION = os.environ.get("STPDE_BACKWARD_PRECISION", "same")


def set_backward_precision(name: str) -> None:
    global BACKWARD_PRECISION
    if name not in ("same", "fp16", "fp16x3"):
        raise ValueError("backward precision must be 'same', 'fp16x3' or 'fp16'")
    BACKWARD_PRECISION = name


def set_default_precision(name: str) -> None:
    global DEFAULT_PRECISION
    if name not in _lib.PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
    DEFAULT_PRECISION = name


# ----------------------------------------------------------------------------------------------
# jet request context: PDELayer announces which partials it needs before calling forward_method
# ----------------------------------------------------------------------------------------------
class JetRequest:
    def __init__(self, spec: JetSpec):
        self.spec = spec
        self.records: List[Tuple[torch.Tensor, Optional[torch.Tensor], torch.Tensor]] = []

    def __enter__(self):
        stack = getattr(_state, "stack", None)
        if stack is None:
            stack = _state.stack = []
        stack.append(self)
        return self

    def __exit__(self, *exc):
        _state.stack.pop()
        return False

    def lookup(self, y: torch.Tensor):
        for rec in self.records:
            if rec[0] is y:
                return rec
        return None


def active_request() -> Optional[JetRequest]:
    stack = getattr(_state, "stack", None)
    return stack[-1] if stack else None


# ----------------------------------------------------------------------------------------------
# workspace cache (one growing byte buffer per device; the C ABI never allocates)
# ----------------------------------------------------------------------------------------------
def _slot(device: torch.device):
    """Scratch is owned per (device, CUDA stream): two streams of one device may have calls in flight at the same time,
    and a call's scratch (workspace, training stash) is only ordered on the stream it was launched on."""
    return (device, torch.cuda.current_stream(device).cuda_stream) if device.type == "cuda" else (device, 0)


_workspaces: Dict[tuple, torch.Tensor] = {}
# What the call-invariant region of a device's workspace currently holds (packed / split weights, per-vertex table):
# identity + version of the latent grid and of every decoder parameter, shapes, precision.  An inference call that finds
# its own key here skips the per-call setup kernels (desc.reserved[1] = 1) - evaluation loops decode thousands of
# pseudo-batches against the same grid and weights (reference experiments/rb2d/evaluation.py:54-69).
# STPDE_SETUP_CACHE=0 disables it (in-place edits through ``.data`` do not bump tensor versions).
_setup_keys: Dict[tuple, tuple] = {}


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    key = _slot(device)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        _workspaces.pop(key, None)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
        _setup_keys.pop(key, None)
    return ws


def release_workspaces() -> None:
    _workspaces.clear()
    _stashes.clear()
    _setup_keys.clear()


class _Stash:
    """Workspace that holds the operand planes of the latest training forward on a device (reverse-mode layout).

    ``token`` identifies the forward call whose planes are in ``ws``; a backward may skip its recompute only when
    its own token is still the current one (any later training forward on the device replaces it)."""

    def __init__(self, ws: torch.Tensor):
        self.ws = ws
        self.token = 0
        self.precision = ""      # arithmetic of the forward that filled it (a single-pass forward leaves no lo planes)
        self.setup_key = None    # (grid, weights, ...) whose call-invariant setup sits in the workspace's fixed region


_stashes: Dict[tuple, _Stash] = {}
_stash_tokens = itertools.count(1)


def _stash(device: torch.device, nbytes: int) -> _Stash:
    key = _slot(device)
    st = _stashes.get(key)
    if st is None or st.ws.numel() < nbytes:
        _stashes.pop(key, None)
        st = _Stash(torch.empty(nbytes, dtype=torch.uint8, device=device))
        _stashes[key] = st
    return st


def bounds_tensors(xmin, xmax, dim: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """xmin / xmax conversion of reference rgi.py:39-45 (python scalar -> float32 ones * value, ...)."""
    import numpy as np

    if isinstance(xmin, (int, float)) or isinstance(xmax, (int, float)):
        lo = float(xmin) * torch.ones([dim], dtype=torch.float32)
        hi = float(xmax) * torch.ones([dim], dtype=torch.float32)
    elif isinstance(xmin, (list, tuple, np.ndarray)) or isinstance(xmax, (list, tuple, np.ndarray)):
        lo, hi = torch.as_tensor(np.asarray(xmin)), torch.as_tensor(np.asarray(xmax))
    else:
        lo, hi = xmin, xmax
    lo = lo.detach().to("cpu", torch.float32).reshape(-1)
    hi = hi.detach().to("cpu", torch.float32).reshape(-1)
    if lo.numel() != dim or hi.numel() != dim:
        raise ValueError(f"xmin/xmax must have {dim} entries")
    return lo, hi


def make_desc(grid: torch.Tensor, q: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor, widths: Sequence[int],
              act: str, act_param: float, spec: JetSpec, precision: str) -> _lib.StpdeDesc:
    d = _lib.StpdeDesc()
    dim = q.shape[-1]
    d.batch, d.npts, d.dim = int(q.shape[0]), int(q.shape[1]), dim
    for k in range(dim):
        d.grid_size[k] = int(grid.shape[1 + k])
        d.xmin[k] = float(lo[k])
        d.xmax[k] = float(hi[k])
    d.channels = int(grid.shape[-1])
    d.n_layers = len(widths)
    for l, w in enumerate(widths):
        d.widths[l] = int(w)
    d.act_kind = _lib.ACT_CODES[act]
    d.act_param = float(act_param)
    d.n_first = len(spec.first)
    for i, k in enumerate(spec.first):
        d.first_dirs[i] = k
    d.n_second = len(spec.second)
    for i, (a, b) in enumerate(spec.second):
        d.second_pairs[i][0] = a
        d.second_pairs[i][1] = b
    d.precision = _lib.PRECISIONS[precision]
    return d


def _i64(values) -> ctypes.Array:
    return (ctypes.c_int64 * len(values))(*[int(v) for v in values])


def raw_forward(grid: torch.Tensor, q: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor,
                Ws: Sequence[torch.Tensor], bs: Sequence[torch.Tensor], act: str, act_param: float,
                spec: JetSpec, precision: str, check: bool = True, stash_out: Optional[list] = None):
    """One C-ABI call per <=10-component sub-spec; returns y [b,p,o], jets [n_jet,b,p,o] or None.

    ``stash_out`` (a list) asks for the TRAINING forward (``stpde_jet_forward_train``): when the whole batch fits
    one chunk of the reverse-mode workspace its operand planes are kept there and the stash token is appended to
    the list, so that the backward of this call can skip the recompute."""
    if _test_backend is not None and not q.is_cuda:
        return _test_backend(grid, q, lo, hi, Ws, bs, act, act_param, spec)
    if not (grid.is_cuda and q.is_cuda):
        raise RuntimeError("the fused decode path needs CUDA tensors (there is no CPU fallback); "
                           f"got grid on {grid.device}, query_pts on {q.device}")
    lib = _lib.load()
    device = q.device
    if grid.dtype != torch.float32 or q.dtype != torch.float32:
        raise TypeError("the fused decode path is float32 (reference arithmetic); got "
                        f"{grid.dtype} / {q.dtype}")
    Wc = [w.detach().to(device=device, dtype=torch.float32).contiguous() for w in Ws]
    Bc = [v.detach().to(device=device, dtype=torch.float32).contiguous() for v in bs]
    widths = [w.shape[0] for w in Wc]
    b, p, dim = q.shape
    o = widths[-1]
    y = torch.empty(b, p, o, dtype=torch.float32, device=device)
    jets = torch.empty(spec.n_jet, b, p, o, dtype=torch.float32, device=device) if spec.n_jet else None
    status = torch.zeros(1, dtype=torch.int32, device=device)
    wptr = (ctypes.c_void_p * len(Wc))(*[w.data_ptr() for w in Wc])
    bptr = (ctypes.c_void_p * len(Bc))(*[v.data_ptr() for v in Bc])
    gstr, qstr = _i64(grid.stride()), _i64(q.stride())
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        subs = spec.split(_lib.MAX_COMPONENTS)
        if stash_out is not None and len(subs) == 1 and precision != "fp32" and len(Wc) >= 3 and b * p > 0:
            desc = make_desc(grid, q, lo, hi, widths, act, act_param, spec, precision)
            nbytes = lib.stpde_backward_workspace_bytes(ctypes.byref(desc))
            if nbytes and lib.stpde_backward_chunk_points(ctypes.byref(desc), nbytes) >= b * p:
                st = _stash(device, nbytes)
                st.token = 0
                # chunks of one training step see the same grid / weights: the split weights and the per-vertex table of
                # the previous training forward in this workspace are still valid (same tensors, same versions)
                key = None
                if os.environ.get("STPDE_SETUP_CACHE", "1") != "0":
                    key = (st.ws.data_ptr(), grid.data_ptr(), grid._version, tuple(grid.shape), tuple(grid.stride()),
                           tuple((w.data_ptr(), w._version) for w in list(Wc) + list(Bc)), tuple(widths), act,
                           float(act_param), precision, tuple(float(v) for v in lo), tuple(float(v) for v in hi))
                desc.reserved[1] = 1 if (key is not None and st.setup_key == key) else 0
                st.setup_key = None
                rc = lib.stpde_jet_forward_train(ctypes.byref(desc), grid.data_ptr(), gstr, q.data_ptr(), qstr, wptr,
                                                 bptr, y.data_ptr(), jets.data_ptr() if jets is not None else None,
                                                 st.ws.data_ptr(), st.ws.numel(), status.data_ptr(), stream)
                _lib.check(rc)
                st.token = next(_stash_tokens)
                st.precision = precision
                st.setup_key = key
                stash_out.append(st.token)
                subs = []
        for sub in subs:
            desc = make_desc(grid, q, lo, hi, widths, act, act_param, sub, precision)
            nbytes = lib.stpde_workspace_bytes(ctypes.byref(desc))
            if nbytes == 0:
                raise _lib.StpdeError(-1, lib.stpde_last_error().decode())
            ws = _workspace(device, nbytes)
            if sub is spec:
                sub_jets = jets
            else:
                sub_jets = torch.empty(sub.n_jet, b, p, o, dtype=torch.float32, device=device)
            # call-invariant setup (weight packing / splitting, per-vertex table) is skipped when the workspace still
            # holds it for exactly these tensors (inference only: a training step changes the weights anyway)
            key = None
            if not torch.is_grad_enabled() and len(subs) == 1 and os.environ.get("STPDE_SETUP_CACHE", "1") != "0":
                key = (ws.data_ptr(), grid.data_ptr(), grid._version, tuple(grid.shape), tuple(grid.stride()),
                       tuple((w.data_ptr(), w._version) for w in list(Wc) + list(Bc)), tuple(widths), act, float(act_param),
                       precision, tuple(float(v) for v in lo), tuple(float(v) for v in hi))
            desc.reserved[1] = 1 if (key is not None and _setup_keys.get(_slot(device)) == key) else 0
            _setup_keys[_slot(device)] = key
            rc = lib.stpde_jet_forward(ctypes.byref(desc), grid.data_ptr(), gstr, q.data_ptr(), qstr, wptr, bptr,
                                       y.data_ptr(), sub_jets.data_ptr() if sub_jets is not None else None,
                                       ws.data_ptr(), ws.numel(), status.data_ptr(), stream)
            _lib.check(rc)
            if sub is not spec:
                nf = len(spec.first)
                jets[:nf] = sub_jets[:nf]
                for i, pair in enumerate(sub.second):
                    jets[nf + spec.second.index(pair)] = sub_jets[nf + i]
    if _capturing(device):
        _captured.append((status, "forward"))
    elif check:
        if not _async_mode():
            _raise_for_status(int(status.item()), "")
        else:
            check_pending()
            _defer_status(status, "forward")
    return y, jets


def raw_backward(grid: torch.Tensor, q: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor,
                 Ws: Sequence[torch.Tensor], bs: Sequence[torch.Tensor], act: str, act_param: float,
                 spec: JetSpec, precision: str, gy: torch.Tensor, gjets: Optional[torch.Tensor],
                 need_grid: bool = True, check: bool = True, stash_token: int = 0):
    """Fused reverse sweep (``stpde_jet_backward``): returns (grid_grad | None, [dW_l], [db_l], gbeta); ``gbeta`` is the
    gradient of the Swish beta (a 1-element tensor, zero for other activations).

    ``stash_token``: token of the training forward whose planes may still sit in the device's stash workspace; if it
    is still current the forward is not recomputed.

    Replaces what ``loss.backward()`` does in the reference (experiments/rb2d/train.py:77) for the decode +
    PDE-derivative part of the graph: autograd's backward through every ``torch.autograd.grad`` call of
    src/pde.py:8 (create_graph=True) and through src/local_implicit_grid.py:47-61.
    """
    lib = _lib.load()
    device = q.device
    Wc = [w.detach().to(device=device, dtype=torch.float32).contiguous() for w in Ws]
    Bc = [v.detach().to(device=device, dtype=torch.float32).contiguous() for v in bs]
    widths = [w.shape[0] for w in Wc]
    gy = gy.detach().to(torch.float32).contiguous()
    if spec.n_jet:
        gjets = gjets.detach().to(torch.float32).contiguous()
    gW = [torch.empty_like(w) for w in Wc]
    gB = [torch.empty_like(v) for v in Bc]
    ggrid = torch.empty(grid.shape, dtype=torch.float32, device=device) if need_grid else None
    status = torch.zeros(1, dtype=torch.int32, device=device)
    gbeta = torch.zeros(1, dtype=torch.float32, device=device)
    wptr = (ctypes.c_void_p * len(Wc))(*[w.data_ptr() for w in Wc])
    bptr = (ctypes.c_void_p * len(Bc))(*[v.data_ptr() for v in Bc])
    gwptr = (ctypes.c_void_p * len(gW))(*[w.data_ptr() for w in gW])
    gbptr = (ctypes.c_void_p * len(gB))(*[v.data_ptr() for v in gB])
    gstr, qstr = _i64(grid.stride()), _i64(q.stride())
    capturing = _capturing(device)
    sync = check and not capturing and not _async_mode()
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        desc = make_desc(grid, q, lo, hi, widths, act, act_param, spec, precision)
        nbytes = lib.stpde_backward_workspace_bytes(ctypes.byref(desc))
        if nbytes == 0:
            raise _lib.StpdeError(-1, lib.stpde_last_error().decode())
        st = _stashes.get(_slot(device))
        reuse = 1 if (stash_token and st is not None and st.token == stash_token and st.ws.numel() >= nbytes
                      and (precision == "fp16" or st.precision != "fp16")) else 0
        ws = st.ws if reuse else _workspace(device, nbytes)
        if not reuse:
            _setup_keys.pop(_slot(device), None)   # the shared workspace is about to be overwritten
        # The adjoints travel through fp16 hi/lo planes behind a power-of-two scale; if one overflows (status bit 1)
        # the sweep is repeated with 6 more bits of headroom.
        desc.reserved[2] = 1 if (reuse and st.precision == "fp16") else 0   # the stash holds fp16 pre-activation planes
        for headroom in (0, 6, 12, 24):
            desc.reserved[0] = headroom
            status.zero_()
            rc = lib.stpde_jet_backward(ctypes.byref(desc), grid.data_ptr(), gstr, q.data_ptr(), qstr, wptr, bptr,
                                        gy.data_ptr(), gjets.data_ptr() if spec.n_jet else None, gwptr, gbptr,
                                        ggrid.data_ptr() if ggrid is not None else None, gbeta.data_ptr(),
                                        ws.data_ptr(), ws.numel(), reuse, status.data_ptr(), stream)
            _lib.check(rc)
            if not sync:
                # asynchronous mode / graph capture: no retry with more headroom is possible without reading the flag
                # back; the overflow is reported by a later call (check_captured() for a graph) instead of being dropped
                if capturing:
                    _captured.append((status, "backward"))
                elif check:
                    check_pending()
                    _defer_status(status, "backward")
                break
            if not (int(status.item()) & 2):
                break
        else:
            _raise_for_status(2, " (backward, after 4 retries with more headroom)")
    return ggrid, gW, gB, gbeta


def fused_backward_supported(q: torch.Tensor, spec: JetSpec, n_layers: int, needs_q: bool, needs_beta: bool) -> bool:
    """The CUDA reverse sweep covers grid / weight / bias / Swish-beta gradients of decoders with >= 3 linear layers."""
    if os.environ.get("STPDE_BACKWARD", "fused") == "torch":
        return False
    return (q.is_cuda and n_layers >= 3 and not needs_q
            and 1 + len(spec.first) + len(spec.second) <= _lib.MAX_COMPONENTS)


def _ddp_average(ddp, grads) -> None:
    """Average decoder-parameter gradients over the process group of a DistributedDataParallel wrapper, in place, with
    ONE all-reduce of a flat buffer (what DDP's reducer does when ``DDP.forward`` runs; the fused path bypasses it).
    ``ddp.no_sync()`` is honoured (gradient accumulation)."""
    if ddp is None or not grads or not getattr(ddp, "require_backward_grad_sync", True):
        return
    import torch.distributed as dist

    group = getattr(ddp, "process_group", None)
    world = dist.get_world_size(group)
    if world <= 1:
        return
    flat = torch.cat([g.reshape(-1).to(torch.float32) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= world
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].reshape(g.shape))
        off += n


class FusedJetQuery(torch.autograd.Function):
    """(grid, q, *params) -> (y, jets).

    Backward: ``stpde_jet_backward`` (fused CUDA reverse sweep) for the gradients w.r.t. the latent grid and the
    decoder weights / biases / the learnable Swish beta; gradients w.r.t. the query points (not needed by the
    reference training loop) re-evaluate the jets with differentiable torch ops (see _torch_jets.py)."""

    @staticmethod
    def forward(ctx, grid, q, lo, hi, act, act_param_t, spec, precision, n_layers, *params):
        ctx.ddp, ctx.train = None, True
        if isinstance(n_layers, tuple):          # (n_layers, DistributedDataParallel wrapper | None, caller's grad mode)
            n_layers, ctx.ddp, ctx.train = n_layers
        Ws, bs = params[:n_layers], params[n_layers:]
        beta = float(act_param_t.detach()) if act_param_t is not None else 1.0
        needs = ctx.needs_input_grad
        stash_out = None
        # (needs_input_grad is True for parameters even under torch.no_grad(): the grad mode of the CALLER decides
        #  whether this is a training forward - inference of a small batch must not pay for keeping the planes)
        if ctx.train and (needs[0] or any(needs[9:])) and fused_backward_supported(q, spec, n_layers, needs[1],
                                                                                  act_param_t is not None and needs[5]):
            stash_out = []            # training forward: keep the operand planes for the backward when they fit
        y, jets = raw_forward(grid, q, lo, hi, Ws, bs, act, beta, spec, precision, stash_out=stash_out)
        ctx.stash_token = stash_out[0] if stash_out else 0
        ctx.save_for_backward(grid, q, act_param_t if act_param_t is not None else torch.empty(0), *params)
        ctx.meta = (lo, hi, act, spec, n_layers, act_param_t is not None)
        ctx.precision = precision
        if jets is None:
            jets = y.new_empty(0)
            ctx.mark_non_differentiable(jets)
        return y, jets

    @staticmethod
    def backward(ctx, gy, gjets):
        from ._torch_jets import query_jets

        grid, q, beta_t, *params = ctx.saved_tensors
        lo, hi, act, spec, n_layers, has_beta = ctx.meta
        needs = ctx.needs_input_grad
        result = [None] * (9 + len(params))
        if not any(needs):
            return tuple(result)
        if fused_backward_supported(q, spec, n_layers, needs[1], has_beta and needs[5]):
            beta = float(beta_t.detach()) if has_beta else 1.0
            gj = gjets if (spec.n_jet > 0 and gjets is not None and gjets.numel() > 0) else None
            if spec.n_jet > 0 and gj is None:
                gj = torch.zeros(spec.n_jet, *gy.shape, dtype=gy.dtype, device=gy.device)
            bprec = ctx.precision if BACKWARD_PRECISION == "same" else BACKWARD_PRECISION
            ggrid, gW, gB, gbeta = raw_backward(grid, q, lo, hi, params[:n_layers], params[n_layers:], act, beta, spec,
                                                bprec, gy, gj, need_grid=needs[0], stash_token=ctx.stash_token)
            _ddp_average(ctx.ddp, list(gW) + list(gB) + [gbeta])
            if needs[0]:
                result[0] = ggrid
            if has_beta and needs[5]:
                result[5] = gbeta.reshape(beta_t.shape).to(beta_t.dtype)
            for i, g in enumerate(list(gW) + list(gB)):
                if needs[9 + i]:
                    result[9 + i] = g.to(params[i].dtype)
            return tuple(result)
        # Points are independent, so the gradient is accumulated over chunks of the point dimension; the chunk is
        # sized so that the autograd tape of the torch re-evaluation (~12 live tensors of rows x components x
        # widest layer) stays below ~2 GB.
        b, p, dim = q.shape
        widest = max(int(t.shape[0]) for t in params[:n_layers])
        per_point = (1 << dim) * (1 + len(spec.first) + len(spec.second)) * widest * 4 * 12 * b
        step = max(64, min(p, int((2 << 30) // max(per_point, 1))))
        lo_d, hi_d = lo.to(q.device), hi.to(q.device)
        have_jets = spec.n_jet > 0 and gjets is not None and gjets.numel() > 0
        acc = {}
        q_grad = torch.zeros_like(q) if needs[1] else None
        for s in range(0, p, step):
            sl = slice(s, min(p, s + step))
            with torch.enable_grad():
                grid_ = grid.detach().requires_grad_(needs[0])
                q_ = q[:, sl].detach().requires_grad_(needs[1])
                beta_ = beta_t.detach().requires_grad_(needs[5]) if has_beta else None
                params_ = [t.detach().requires_grad_(needs[9 + i]) for i, t in enumerate(params)]
                y, jets = query_jets(grid_, q_, lo_d, hi_d, params_[:n_layers], params_[n_layers:], act, beta_, spec)
                outs, gouts = [y], [gy[:, sl]]
                if have_jets and jets is not None:
                    outs.append(jets)
                    gouts.append(gjets[:, :, sl])
                wanted = [(0, grid_)] * needs[0] + [(1, q_)] * needs[1] + ([(5, beta_)] if has_beta and needs[5] else [])
                wanted += [(9 + i, t) for i, t in enumerate(params_) if needs[9 + i]]
                grads = torch.autograd.grad(outs, [t for _, t in wanted], gouts, allow_unused=True) if wanted else []
            for (slot, _), g in zip(wanted, grads):
                if g is None:
                    continue
                if slot == 1:
                    q_grad[:, sl] = g
                else:
                    acc[slot] = g if slot not in acc else acc[slot] + g
        _ddp_average(ctx.ddp, [g for slot, g in acc.items() if slot >= 5])
        for slot, g in acc.items():
            result[slot] = g
        if needs[1]:
            result[1] = q_grad
        return tuple(result)


def fused_query(grid: torch.Tensor, q: torch.Tensor, xmin, xmax, layers, act: str, act_param,
                spec: Optional[JetSpec] = None, precision: Optional[str] = None, ddp=None):
    """y [b,p,o] (and jets [n_jet,b,p,o] when ``spec`` asks for derivatives).

    ``ddp``: the DistributedDataParallel wrapper the decoder layers came from, if any (its process group receives the
    gradient averaging the bypassed ``DDP.forward`` would have armed)."""
    spec = spec or JetSpec()
    precision = precision or DEFAULT_PRECISION
    lo, hi = bounds_tensors(xmin, xmax, q.shape[-1], q.device)
    Ws = [l.weight for l in layers]
    bs = [l.bias for l in layers]
    nl = (len(layers), ddp, torch.is_grad_enabled())
    y, jets = FusedJetQuery.apply(grid, q, lo, hi, act, act_param, spec, precision, nl, *Ws, *bs)
    return y, (jets if spec.n_jet else None)
