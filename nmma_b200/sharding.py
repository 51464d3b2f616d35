"""Multi-GPU: contiguous row blocks per rank + one gather of the per-point log-likelihoods.

Evaluations are independent, so the path shards trivially (the reference parallelises the
same way, one point per MPI task: ``nmma/core/mpi_setup.py:651-683``).  ``points[N, P]`` is cut
into contiguous blocks ``[r*N/G, (r+1)*N/G)``; surrogate weights and the observation table are
replicated (~1-3 MB); every rank evaluates its block on its own GPU; the only collective is an
all-gather (or gather-to-root) of ``logL`` over NCCL/NVLink -- 8 bytes per point.  There is no
data-path collective inside the evaluation.  ``backend='gloo'`` serves CPU tests of the host
logic (the local evaluator is injected there).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def shard_bounds(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of rank ``rank``: sizes differ by at most one, earlier ranks take the extra."""
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n: int, world_size: int):
    return [shard_bounds(n, world_size, r)[1] - shard_bounds(n, world_size, r)[0] for r in range(world_size)]


class ShardedEvaluator:
    """Evaluate a global batch across the ranks of an initialised ``torch.distributed`` group.

    ``local_eval(points_local) -> logL_local`` is the per-rank evaluator (normally
    ``EMTransientLikelihood.log_likelihood_batch`` on CUDA tensors)."""

    def __init__(self, local_eval: Callable, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.local_eval = local_eval
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def local_block(self, points_global):
        lo, hi = shard_bounds(len(points_global), self.world, self.rank)
        return points_global[lo:hi]

    def gather(self, local_logl, n_global: Optional[int] = None, dst: Optional[int] = None):
        """All-gather (``dst is None``) or gather-to-``dst`` of the per-rank blocks, in rank order."""
        import torch
        dist = self.dist
        t = local_logl if isinstance(local_logl, torch.Tensor) else torch.from_numpy(np.asarray(local_logl, float))
        t = t.contiguous()
        if n_global is None:
            n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
            dist.all_reduce(n, group=self.group)
            n_global = int(n.item())
        sizes = shard_sizes(n_global, self.world)
        assert sizes[self.rank] == t.numel(), "local block does not match the contiguous partition"
        if len(set(sizes)) == 1:
            out = torch.empty(n_global, dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(out, t, group=self.group)   # ncclAllGather over NVLink
        else:   # ragged tail: pad every block to the largest one, gather, trim
            m = max(sizes)
            padded = torch.zeros(m, dtype=t.dtype, device=t.device)
            padded[:t.numel()] = t
            buf = torch.empty(m * self.world, dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(buf, padded, group=self.group)
            out = torch.cat([buf[r * m:r * m + sizes[r]] for r in range(self.world)])
        if dst is not None and self.rank != dst:
            return None
        return out

    def evaluate(self, points_global, dst: Optional[int] = None):
        """Every rank passes the same global ``points[N, P]`` (or builds its block itself and calls
        :meth:`gather`); returns the full ``logL[N]`` on every rank (or on ``dst`` only)."""
        local = self.local_eval(self.local_block(points_global))
        return self.gather(local, n_global=len(points_global), dst=dst)
