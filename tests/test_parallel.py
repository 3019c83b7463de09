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
