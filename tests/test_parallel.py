"""World-size-2 gloo test of the point-sharding helpers (CPU; the N>1 path of bench.py / training)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from space_time_pde_b200.parallel import StepReducer, shard_bounds, shard_points


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                       # same data and model on every rank
    model = torch.nn.Linear(3, 2)
    pts = torch.rand(1, 101, 3)
    tgt = torch.rand(1, 101, 2)
    p_loc, t_loc = shard_points(pts, tgt)
    loss_sum = (model(p_loc) - t_loc).abs().sum()              # L1 sum over the local shard
    # global mean = sum of sums / sum of counts; backprop the local sum scaled by the GLOBAL count
    (loss_sum / (pts.shape[1] * 2)).backward()
    means = StepReducer(model.parameters()).reduce({"reg": loss_sum}, {"reg": p_loc.shape[1] * 2})
    if rank == 0:
        ref = torch.nn.Linear(3, 2)
        ref.load_state_dict(model.state_dict())
        ref_loss = (ref(pts) - tgt).abs().mean()
        ref_loss.backward()
        out.put((float(means["reg"]), float(ref_loss), float((model.weight.grad - ref.weight.grad).abs().max())))
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_two_rank_step_matches_single_process():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    mean, ref, gerr = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert abs(mean - ref) < 1e-6 and gerr < 1e-6


def _ddp_worker(rank, world, port, out):
    """DDP(ImNet) through the fused route (torch stand-in backend on CPU): the decoder gradients must come out
    identical on both ranks and equal to the average of the per-rank gradients, as DDP's reducer would give
    (reference experiments/rb2d/train_ddp.py: imnet = DDP(imnet))."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import space_time_pde_b200 as sp
    from space_time_pde_b200 import _torch_jets, jets

    jets.set_test_backend(lambda grid, q, lo, hi, Ws, bs, act, beta, spec: _torch_jets.query_jets(
        grid, q, lo, hi, list(Ws), list(bs), act, torch.tensor(beta), spec))
    torch.manual_seed(0)
    imnet = sp.ImNet(dim=3, in_features=8, out_features=4, nf=4, activation=sp.NONLINEARITIES["softplus"])
    ddp = torch.nn.parallel.DistributedDataParallel(imnet)
    grid = torch.randn(1, 3, 4, 5, 8) * 0.5
    gen = torch.Generator().manual_seed(100 + rank)            # different points per rank
    q = torch.rand(1, 64, 3, generator=gen)
    layer = sp.get_rb2_pde_layer(t_crop=2., z_crop=1., x_crop=1., use_continuity=True)

    def local_loss(model):
        layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
        y, res = layer(q)
        return y.abs().mean() + 0.0125 * torch.stack(list(res.values())).abs().mean()

    local_loss(ddp).backward()
    got = torch.cat([p.grad.reshape(-1) for p in imnet.parameters()]).clone()
    imnet.zero_grad()
    local_loss(imnet).backward()                                 # no wrapper: purely local gradients
    want = torch.cat([p.grad.reshape(-1) for p in imnet.parameters()]).clone()
    dist.all_reduce(want)
    want /= world
    gathered = [torch.zeros_like(got) for _ in range(world)]
    dist.all_gather(gathered, got)
    if rank == 0:
        out.put((float((gathered[0] - gathered[1]).abs().max()), float((got - want).abs().max() / want.abs().max())))
    dist.destroy_process_group()


def test_ddp_wrapped_decoder_gradients_are_averaged_across_ranks():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    across, err = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert across == 0.0 and err < 1e-6


def _all_ok_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    ctx = bench.Ctx(torch.device("cpu"), rank, world)
    # an optional path (CUDA-graph replay of the training chunks) is taken by every rank or by none
    res = (ctx.all_ok(True), ctx.all_ok(rank != 1), ctx.all_ok(False))
    # the reducer writes the reduced gradients INTO pre-existing .grad tensors (graph replays accumulate into them)
    w = torch.nn.Parameter(torch.ones(3))
    w.grad = torch.full((3,), float(rank + 1))
    keep = w.grad
    StepReducer([w]).reduce({"reg": torch.tensor(1.0)}, {"reg": 1.0})
    out.put((rank, res, keep is w.grad, w.grad.tolist()))
    dist.destroy_process_group()


def test_optional_paths_are_taken_by_all_ranks_or_none_and_grads_stay_in_place():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_all_ok_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(out.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res, same_tensor, grad in got:
        assert res == (True, False, False)
        assert same_tensor and grad == [3.0, 3.0, 3.0]          # 1 + 2 summed over the ranks, written in place
