"""Multi-GPU plumbing for the latent-ODE solve: shard trajectories, reduce parameter gradients once.

Trajectories are independent (``Dynamics.forward`` has no cross-trajectory term,
``models/blackbox_ode.py:97-109``), so the batch axis shards with no forward communication.  The only
exchange of a training step is the sum of the parameter gradients, done as ONE ``all_reduce`` over a
single flat fp32 buffer (dynamics + initial-state net + whatever else the caller registers).  The
reference has no distributed code at all; this is where ``training_*.py`` would gain it
(``run_batch``, ``training_cvs.py:147-157``).

One process per GPU; ``torch.distributed`` backend ``nccl`` on GPUs, ``gloo`` in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_total: int, rank: int, world_size: int):
    """Contiguous row range [lo, hi) of rank ``rank``; the first ``n_total % world_size`` ranks get one extra."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rows(x: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    lo, hi = shard_bounds(x.shape[0], rank, world_size)
    return x[lo:hi]


class FlatGradReducer:
    """Sum the ``.grad`` of a fixed parameter list across ranks with one collective.

    The flat buffer is allocated once; ``reduce()`` packs, all-reduces (SUM) and unpacks in place.
    Parameters whose ``.grad`` is ``None`` contribute zeros (every rank must issue the same collective).
    """

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no parameters to reduce")
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        self.group = group

    def reduce(self, average: bool = False):
        # pack: one concatenation kernel instead of one copy per parameter
        views = list(self.flat.split([p.numel() for p in self.params]))
        have = [p.grad is not None for p in self.params]
        if all(have):
            torch.cat([p.grad.reshape(-1) for p in self.params], out=self.flat)
        else:
            self.flat.zero_()
            src = [p.grad.reshape(-1) for p, h in zip(self.params, have) if h]
            if src:
                torch._foreach_copy_([v for v, h in zip(views, have) if h], src)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            if average:
                self.flat.div_(dist.get_world_size(self.group))
        # unpack: one multi-tensor copy
        for p, v, h in zip(self.params, views, have):
            if not h:
                p.grad = torch.empty_like(p)
        torch._foreach_copy_([p.grad.view(-1) if p.grad.is_contiguous() else p.grad for p in self.params],
                             [v if p.grad.is_contiguous() else v.view_as(p) for p, v in zip(self.params, views)])
        return self.flat


def dopri5_shard_options(local_batch: int, group=None, device=None, options=None):
    """``options`` for ``odeint(..., method="dopri5")`` on ONE shard of a trajectory-sharded batch.

    torchdiffeq's dopri5 takes one step size for the whole batch (error norm over all B*S elements),
    so a sharded solve has a real exchange step: per attempted step one all-reduce of two float64 sums.  The returned
    options make the solver run one pass per launch and sum the batch-wide norms over ``group`` in between
    (``torchdiffeq_api.Dopri5ShardSolve``); every rank then takes the step sequence of the unsharded solve."""
    n = torch.tensor([int(local_batch)], dtype=torch.int64, device=device)
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world > 1:
        dist.all_reduce(n, op=dist.ReduceOp.SUM, group=group)

    def reducer(x):
        if world > 1:
            dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
        return x

    out = dict(options or {})
    out["shard_reducer"] = reducer
    out["global_batch"] = int(n.item())
    return out
