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

import struct
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


class PeerContext:
    """Symmetric (peer-mapped) memory of one rank for the fused data-parallel path (csrc/nrm_loss_adam.cu, "data parallelism over
    peer memory"): the flat gradient buffer every peer reads while it runs Adam, and a control block [BatchNorm sums 2 x 528 doubles |
    flag words].  `torch.distributed._symmetric_memory` allocates and maps the memory (CUDA VMM handles exchanged through the
    process group's store); the kernels only ever see the raw pointers packed into `self.ctx` (the C ABI's PeerCtx struct).
    Raises if the GPUs of the group cannot map each other's memory; callers fall back to NCCL."""

    def __init__(self, n_floats: int, device, group=None, rows: int = 0, delta_off: int = 0, delta_n: int = 0):
        """rows / delta_off / delta_n > 0: sparse exchange of the per-user bias gradient (`rows` (user id, value) entries per rank
        instead of delta's dense range [delta_off, delta_off + delta_n) of the flat buffers)."""
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        lib = _lib.load()
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError('PeerContext: at most 8 ranks (one NVSwitch domain)')
        stats_bytes = int(lib.nrm_peer_stats_bytes())
        flag_bytes = 4 * int(lib.nrm_peer_flag_words())
        self.grads = symm.empty(n_floats, dtype=torch.float32, device=device)
        self.control = symm.empty((stats_bytes + flag_bytes + 255) // 256 * 256, dtype=torch.uint8, device=device)
        self.sparse_rows = int(rows) if delta_n > 0 else 0
        srows = max(self.sparse_rows, 1)
        self.sparse = symm.empty(srows * 12 + 16, dtype=torch.uint8, device=device)       # [rows] int64 user ids | [rows] float32 values
        self.grads.zero_()
        self.control.zero_()
        self.sparse.zero_()
        torch.cuda.synchronize(device)
        self._hg = symm.rendezvous(self.grads, group)
        self._hc = symm.rendezvous(self.control, group)
        self._hs = symm.rendezvous(self.sparse, group)
        gp = [int(x) for x in self._hg.buffer_ptrs]
        cp = [int(x) for x in self._hc.buffer_ptrs]
        sp = [int(x) for x in self._hs.buffer_ptrs]
        self.sparse_uid = self.sparse[:srows * 8].view(torch.int64)
        self.sparse_val = self.sparse[srows * 8:srows * 12].view(torch.float32)
        self.acc = torch.zeros(max(int(delta_n), 1), dtype=torch.int64, device=device)      # fixed-point accumulator, zero between steps
        self.gridbar = torch.zeros(4, dtype=torch.int32, device=device)
        pad8 = lambda xs: xs + [0] * (8 - len(xs))
        raw = struct.pack('<iiq', self.rank, self.world, n_floats) + struct.pack('<8Q', *pad8(gp)) \
            + struct.pack('<8Q', *pad8([c + stats_bytes for c in cp])) + struct.pack('<8Q', *pad8(cp)) \
            + struct.pack('<8Q', *pad8(sp)) + struct.pack('<8Q', *pad8([x + srows * 8 for x in sp])) \
            + struct.pack('<QQqqq', self.acc.data_ptr(), self.gridbar.data_ptr(), int(delta_off), int(delta_n) if self.sparse_rows else 0, srows)
        assert len(raw) == int(lib.nrm_peer_ctx_bytes())
        self.ctx = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
        self.ticket = torch.zeros(4, dtype=torch.int32, device=device)
        self.stats_out = torch.zeros(2, stats_bytes // 16, dtype=torch.float64, device=device)     # [forward | backward] global sums
        self.stats_local = torch.zeros(2, stats_bytes // 16, dtype=torch.float64, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)                                  # every rank's flags are zero before anybody signals


class DataParallel:
    """Attach to a UserModel: `DataParallel(model)`; afterwards model.forward / backward
    run the collectives described in the module docstring."""

    def __init__(self, model, group=None, sync_bn: bool = False, broadcast: bool = True, peer_memory: bool = True):
        """peer_memory: let FusedTrainStep average the gradients inside its Adam kernel over peer-mapped memory (PeerContext) when
        the backend is NCCL and the GPUs can map each other; otherwise (and on the module path) NCCL all-reduce buckets."""
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
        self.peer_memory = bool(peer_memory) and dist.get_backend(group) == 'nccl'
        self.peer: Optional[PeerContext] = None
        self.peer_error: Optional[str] = None
        model._dp = self
        if broadcast:
            flat = model.flat_parameters()
            dist.broadcast(flat.buf, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            for b in (model.bn.running_mean, model.bn.running_var, model.bn.num_batches_tracked):
                dist.broadcast(b, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)

    def peer_context(self, n_floats: int, device, rows: int = 0, delta_off: int = 0, delta_n: int = 0) -> Optional[PeerContext]:
        """A NEW symmetric-memory context for one FusedTrainStep (its flags count that step object's optimizer steps, so contexts
        are never shared); None when peer mapping is off or impossible.  Collective: every rank must call it at the same point."""
        if not self.peer_memory or self.peer_error is not None:
            return None
        ok, ctx = 1, None
        try:
            ctx = PeerContext(n_floats, device, self.group, rows, delta_off, delta_n)
        except Exception as ex:                              # no P2P / fabric handles, old driver ...: NCCL stays the transport
            self.peer_error = f'{type(ex).__name__}: {ex}'
            ok = 0
        t = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        if int(t.item()) == 0:                               # all ranks or none
            ctx = None
            self.peer_error = self.peer_error or 'a peer rank could not map symmetric memory'
        self.peer = ctx
        return ctx

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
