"""Data-parallel training glue: one process per GPU, impressions sharded by rank.

The reference is single-process; north_star asks for data-parallel
training with the gradient all-reduce bucketed behind backward.  The flat gradient buffer
makes that two NCCL calls: the head bucket ([bn .. out_mlp] + delta, contiguous at the end
of the layout) is reduced while the encoder backward still runs, the encoder bucket right
after it.  With `sync_bn=True` the BatchNorm batch statistics (2x264 doubles forward, 2x264
backward) are all-reduced too, which makes N ranks x B impressions bit-for-bit the same
model as one process on the N*B batch; the default keeps per-replica statistics (standard
DDP behaviour, no collective in the forward).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def bind_to_local_cpus(device_index: int):
    """Pin this process (and the pinned host buffers it allocates afterwards: first touch) to the CPU cores NVML reports
    as local to GPU `device_index`.  With one process per GPU every rank then stages its batches from the memory of
    the socket its own PCIe root hangs off, instead of wherever the scheduler happened to start it.  Returns the
    sorted core list, or None when NVML or the affinity call is not available (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        handle = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + uuid) if not uuid.startswith('GPU-') else uuid)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:                                   # an optimisation only: never a reason to stop
        return None


def shard_range(total: int, rank: int, world: int):
    """Contiguous [begin, end) slice of `total` impressions owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class GradientBuckets:
    """Average the flat gradient buffer across ranks in two buckets."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self._avg = dist.get_backend(group) == 'nccl'
        self.handles = []

    def _reduce(self, t: torch.Tensor):
        if self._avg:
            self.handles.append(dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        else:                                   # gloo (CPU tests): SUM then scale
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.div_(self.world)

    def reduce(self, flat_grad: torch.Tensor, begin: int, end: int):
        if end > begin:
            self._reduce(flat_grad[begin:end])

    def wait(self):
        for h in self.handles:
            h.wait()
        self.handles = []


class DataParallel:
    """Attach to a UserModel: `DataParallel(model)`; afterwards model.forward / backward
    run the collectives described in the module docstring."""

    def __init__(self, model, group=None, sync_bn: bool = False, broadcast: bool = True):
        if not dist.is_initialized():
            raise RuntimeError('torch.distributed is not initialised')
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.sync_bn = sync_bn
        self.buckets = GradientBuckets(group)
        self.comm_stream: Optional[torch.cuda.Stream] = None
        self._head_begin = None
        self._rows_checked = None
        model._dp = self
        if broadcast:
            flat = model.flat_parameters()
            dist.broadcast(flat.buf, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            for b in (model.bn.running_mean, model.bn.running_var, model.bn.num_batches_tracked):
                dist.broadcast(b, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)

    # ---- BatchNorm statistics -----------------------------------------------------------
    def all_reduce_stats(self, sums: torch.Tensor, local_rows: int) -> int:
        """Sum the BatchNorm statistics over the ranks; returns the global row count.  Shards must be equal (the gradient
        average weights every rank the same): checked with one small all-reduce whenever the local row count changes."""
        if local_rows != self._rows_checked:
            t = torch.tensor([local_rows, -local_rows], dtype=torch.int64, device=sums.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            hi, lo = int(t[0]), -int(t[1])
            if hi != lo:
                raise RuntimeError(f'DataParallel(sync_bn=True): ranks hold between {lo} and {hi} candidate rows; shard the global '
                                   'batch evenly (drop or pad the ragged tail) -- unequal shards would bias the averaged gradients')
            self._rows_checked = local_rows
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
        return local_rows * self.world

    # ---- gradient buckets ---------------------------------------------------------------
    def _split(self, flat):
        if self._head_begin is None:
            self._head_begin = next(off for name, off, _, _ in flat.slots if name == 'bn.weight')
        return self._head_begin

    def _on_comm_stream(self, fn):
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream()
        cur = torch.cuda.current_stream()
        self.comm_stream.wait_stream(cur)          # the gradients just written are visible
        with torch.cuda.stream(self.comm_stream):
            fn()

    def reduce_head_bucket(self, g: torch.Tensor, flat):
        # [bn.weight .. out_mlp.fc2.bias]; delta (behind it in the layout) is averaged by the loss backward (reduce_delta)
        hb = self._split(flat)
        self._on_comm_stream(lambda: self.buckets.reduce(g, hb, flat.fixed))

    def reduce_delta(self, ddelta: torch.Tensor):
        """Average the per-user bias gradient across ranks, in stream order on the caller's stream (the result is handed to
        autograd right away, see engine._LossFn.backward)."""
        if ddelta.numel() == 0:
            return
        if self.buckets._avg:
            dist.all_reduce(ddelta, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(ddelta, op=dist.ReduceOp.SUM, group=self.group)
            ddelta.div_(self.world)

    def reduce_encoder_bucket(self, g: torch.Tensor, flat):
        hb = self._split(flat)
        self._on_comm_stream(lambda: self.buckets.reduce(g, 0, hb))

    def wait(self):
        self.buckets.wait()
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
