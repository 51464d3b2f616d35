"""N > 1 host logic on CPU: world_size-2 gloo group, contiguous row blocks, gather in rank order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nmma_b200.sharding import HostResultBuffer, ShardedEvaluator, shard_bounds, shard_sizes


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(shard_sizes(n, w)) - min(shard_sizes(n, w)) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pts = torch.from_numpy(np.random.default_rng(5).normal(size=(n, 6)))      # same global batch on every rank
        local_eval = lambda p: (p ** 2).sum(dim=1) + 1000.0 * rank * 0            # stand-in for the GPU evaluator
        ev = ShardedEvaluator(local_eval)
        full = ev.evaluate(pts)
        root = ev.evaluate(pts, dst=0)
        lo, hi = shard_bounds(n, world, rank)
        mine = ev.gather(local_eval(pts[lo:hi]))
        # prior sweep sharded by GLOBAL point index: rank r scores [lo_r, hi_r) of one counter-based sequence
        swept = ev.sweep(lambda first, m: torch.arange(first, first + m, dtype=torch.float64) * 0.5, n)
        # pipelined gather (CPU path of gather_overlapped): equal blocks, two slots, rank order
        m = 8
        local_bufs = [torch.empty(m, dtype=torch.float64) for _ in range(2)]
        full_bufs = [torch.empty(m * world, dtype=torch.float64) for _ in range(2)]
        steps = []
        for step in range(3):
            slot = ev.gather_overlapped(lambda out, s=step: out.copy_(torch.arange(m, dtype=torch.float64) + 100 * rank + 1000 * s),
                                        local_bufs, full_bufs, step)
            steps.append(full_bufs[slot].clone().numpy())
        ev.drain()
        # host consumer: every rank writes its slice of one shared-memory vector, rank 0 reads it in place
        hb = HostResultBuffer(n, rank, world, name=f"nmma_b200_test_{port}", register=False)
        dist.barrier()
        hb.attach()
        hb.local[:] = local_eval(pts[lo:hi]).numpy()
        dist.barrier()
        host_full = np.array(hb.full) if rank == 0 else None
        dist.barrier()
        hb.close()
        q.put((rank, full.numpy(), None if root is None else root.numpy(), mine.numpy(), swept.numpy(), steps, host_full))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [64, 101])
def test_world_size_2_gather(n):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = (torch.from_numpy(np.random.default_rng(5).normal(size=(n, 6))) ** 2).sum(dim=1).numpy()
    for rank, full, root, mine, swept, steps, host_full in res:
        assert np.array_equal(full, expect) and np.array_equal(mine, expect)
        assert (root is None) == (rank != 0)
        if root is not None:
            assert np.array_equal(root, expect)
            assert np.array_equal(host_full, expect)
        assert np.array_equal(swept, np.arange(n) * 0.5)
        for s, got in enumerate(steps):
            want = np.concatenate([np.arange(8.0) + 100 * r + 1000 * s for r in range(world)])
            assert np.array_equal(got, want)
