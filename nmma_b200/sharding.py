"""Multi-GPU: contiguous row blocks per rank + one gather of the per-point log-likelihoods.

Evaluations are independent, so the path shards trivially (the reference parallelises the
same way, one point per MPI task: ``nmma/core/mpi_setup.py:651-683``).  ``points[N, P]`` is cut
into contiguous blocks ``[r*N/G, (r+1)*N/G)``; surrogate weights and the observation table are
replicated (~1-5 MB); every rank evaluates its block on its own GPU; the only collective is an
all-gather (or gather-to-root) of ``logL`` over NCCL/NVLink -- 8 bytes per point.  There is no
data-path collective inside the evaluation, hence nothing to fuse with compute; what matters is that
the gather never sits on the critical path:

* device consumers (:meth:`ShardedEvaluator.gather_overlapped`): the all-gather of step *i* runs on a side
  stream under the kernels of step *i + 1* (double-buffered blocks, event-ordered);
* host consumers (:class:`HostResultBuffer`): no collective at all -- every rank's device-to-host copy lands
  in its slice of one page-locked shared-memory buffer that the consumer process (rank 0) reads in place.

``backend='gloo'`` serves CPU tests of the host logic (the local evaluator is injected there).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def bind_to_gpu_numa_node(device: int) -> Optional[str]:
    """Pin this process to the CPUs local to ``cuda:device`` (``/sys/bus/pci/devices/<bus id>/local_cpulist``) so that its
    page-locked buffers are allocated on, and its copies issued from, the GPU's own NUMA node.  With eight ranks pushing
    48 MB per step through one host, remote-node staging memory is what end-to-end scaling loses first.  Returns the CPU list
    it bound to, or None when the topology is not visible (containers without /sys, a restricted cpuset): never raises."""
    import os
    try:
        import torch
        props = torch.cuda.get_device_properties(device)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as fh:
            text = fh.read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed or allowed == os.sched_getaffinity(0):
            return None
        os.sched_setaffinity(0, allowed)
        return text
    except Exception:  # noqa: BLE001
        return None


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
        self._side = None            # side stream + events of gather_overlapped (CUDA only)
        self._pending = []           # [(event gather done, slot)]

    def local_block(self, points_global):
        lo, hi = shard_bounds(len(points_global), self.world, self.rank)
        return points_global[lo:hi]

    def gather(self, local_logl, n_global: Optional[int] = None, dst: Optional[int] = None, out=None):
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
            if out is None:
                out = torch.empty(n_global, dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(out, t, group=self.group)   # ncclAllGather over NVLink
        else:   # ragged tail: pad every block to the largest one, gather, trim
            m = max(sizes)
            padded = torch.zeros(m, dtype=t.dtype, device=t.device)
            padded[:t.numel()] = t
            buf = torch.empty(m * self.world, dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(buf, padded, group=self.group)
            res = torch.cat([buf[r * m:r * m + sizes[r]] for r in range(self.world)])
            out = res if out is None else out.copy_(res)
        if dst is not None and self.rank != dst:
            return None
        return out

    def evaluate(self, points_global, dst: Optional[int] = None):
        """Every rank passes the same global ``points[N, P]`` (or builds its block itself and calls
        :meth:`gather`); returns the full ``logL[N]`` on every rank (or on ``dst`` only)."""
        local = self.local_eval(self.local_block(points_global))
        return self.gather(local, n_global=len(points_global), dst=dst)

    # ---- gather off the critical path (device consumers) -------------------------------------------------
    def gather_overlapped(self, compute: Callable, local_bufs, full_bufs, step: int):
        """One pipelined step: ``compute(local_bufs[slot])`` enqueues this step's kernels on the current stream; the
        all-gather of the result runs on a side stream, ordered after them by an event, while the caller already
        enqueues the next step.  ``slot = step % len(local_bufs)``; a slot is reused only after its previous gather
        finished (event wait on the compute stream).  Call :meth:`drain` before reading ``full_bufs``.
        Equal block sizes on every rank (the sharded sweep pads its tail)."""
        import torch
        slot = step % len(local_bufs)
        if not local_bufs[slot].is_cuda:            # gloo / CPU: no streams to overlap
            compute(local_bufs[slot])
            self.dist.all_gather_into_tensor(full_bufs[slot], local_bufs[slot], group=self.group)
            return slot
        if self._side is None:
            self._side = torch.cuda.Stream(device=local_bufs[slot].device)
            self._ev_done = [None] * len(local_bufs)
        cur = torch.cuda.current_stream(local_bufs[slot].device)
        if self._ev_done[slot] is not None:
            cur.wait_event(self._ev_done[slot])      # the gather that last read this slot has finished
        compute(local_bufs[slot])
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self._side):
            self._side.wait_event(ready)
            self.dist.all_gather_into_tensor(full_bufs[slot], local_bufs[slot], group=self.group)
            done = torch.cuda.Event()
            done.record(self._side)
        self._ev_done[slot] = done
        return slot

    def drain(self):
        """Make the current stream wait for every outstanding overlapped gather."""
        import torch
        if self._side is not None:
            torch.cuda.current_stream(self._side.device).wait_stream(self._side)

    # ---- prior sweep sharded by global point index (BASELINE.json configs[4]) -----------------------------------
    def sweep(self, local_sweep: Callable, n_global: int, dst: Optional[int] = None):
        """``local_sweep(first_index, n) -> logL[n]`` draws and scores points ``first_index .. first_index + n - 1`` of the
        global sequence (counter-based Philox: point *i* is the same on any rank, ``nmma_b200_logl_sweep``); rank *r*
        takes the contiguous index range ``shard_bounds(n_global, world, r)``; one gather returns ``logL[n_global]``."""
        lo, hi = shard_bounds(n_global, self.world, self.rank)
        return self.gather(local_sweep(lo, hi - lo), n_global=n_global, dst=dst)


class HostResultBuffer:
    """``logL[n_global]`` in POSIX shared memory, page-locked by every rank (``cudaHostRegister``): rank *r* copies its
    block device-to-host straight into ``[lo_r, hi_r)`` and the host consumer (rank 0's sampler process) reads the
    whole vector in place -- the host-side analogue of the reference's MPI result gather
    (``nmma/core/mpi_setup.py:651-683``) without moving a byte between GPUs or between processes."""

    def __init__(self, n_global: int, rank: int, world: int, name: str, register: bool = True):
        from multiprocessing import shared_memory
        self.n, self.rank, self.world = int(n_global), rank, world
        nbytes = max(self.n, 1) * 8
        if rank == 0:
            try:
                shared_memory.SharedMemory(name=name).unlink()
            except FileNotFoundError:
                pass
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=nbytes)
        self._name = name
        self._registered = False
        self._register = register
        self._nbytes = nbytes

    def attach(self):
        """Call on every rank after a barrier that follows rank 0's construction."""
        from multiprocessing import shared_memory
        if self.rank != 0:
            self.shm = shared_memory.SharedMemory(name=self._name)
        self.full = np.ndarray((self.n,), dtype=np.float64, buffer=self.shm.buf)
        lo, hi = shard_bounds(self.n, self.world, self.rank)
        self.lo, self.hi = lo, hi
        self.local = self.full[lo:hi]
        if self._register and hi > lo:
            import torch
            rc = torch.cuda.cudart().cudaHostRegister(self.local.ctypes.data, self.local.nbytes, 0)
            self._registered = int(rc) == 0
        return self

    def close(self):
        import torch
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self.local.ctypes.data)
            self._registered = False
        self.full = self.local = None
        self.shm.close()
        if self.rank == 0:
            try:
                self.shm.unlink()
            except FileNotFoundError:
                pass
