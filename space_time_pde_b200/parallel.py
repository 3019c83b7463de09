"""Point sharding across ranks.

Query points are independent given (latent grid, decoder weights), so the forward + residual path needs no
data-path collective: every rank decodes its own slice of the point dimension.  A training step needs exactly one
all-reduce of a flat float32 buffer  [loss sums | counts | flat gradients]  (mean losses are recovered as sum / count,
which equals the single-process mean over the whole batch).  Reference analogue: DDP over crops with gloo
(experiments/rb2d/train_ddp.py:48,361-366,402-406).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of n items for `rank`; the first n % world ranks get one extra item."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_points(points: torch.Tensor, *others: torch.Tensor, rank: int = None, world: int = None):
    """Slice [b, p, ...] tensors along the point dimension for this rank."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_bounds(points.shape[1], rank, world)
    out = tuple(t[:, lo:hi] for t in (points,) + others)
    return out if others else out[0]


class StepReducer:
    """Pack loss sums / counts / gradients into ONE flat buffer and all-reduce it once per step."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]

    def reduce(self, sums: Dict[str, torch.Tensor], counts: Dict[str, float]) -> Dict[str, torch.Tensor]:
        """sums: local loss SUMS (already backpropagated); returns global MEANS, grads become global sums."""
        names = sorted(sums)
        device = self.params[0].device if self.params else next(iter(sums.values())).device
        scalars = torch.stack([sums[k].detach().float().reshape(()).to(device) for k in names] +
                              [torch.tensor(float(counts[k]), device=device) for k in names])
        grads = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in self.params]
        flat = torch.cat([scalars] + grads)
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        k = len(names)
        means = {name: flat[i] / flat[k + i] for i, name in enumerate(names)}
        off = 2 * k
        for p in self.params:
            n = p.numel()
            g = flat[off:off + n].reshape(p.shape).to(p.dtype)
            if p.grad is not None and p.grad.shape == g.shape and p.grad.dtype == g.dtype:
                p.grad.copy_(g)          # in place: a step replayed from CUDA graphs accumulates into these very tensors
            else:
                p.grad = g.clone()
            off += n
        return means
