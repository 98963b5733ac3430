"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed over NCCL / NVLink.

The reference is single-process; this is new capability along the two axes the path shards on:

* inference -- sequences are independent: `shard_range` gives each rank a contiguous slice, there is NO collective on
  the data path;
* training -- the batch is sharded; the one exchange per step is the average of the fp32 parameter gradients.
  `GradSync` hooks the backward of vp3d_b200.training: every large gradient (a convolution weight, 4-12 MB) is
  all-reduced asynchronously on NCCL's stream the moment its weight-gradient GEMM has been issued, so the transfer
  overlaps the remaining backward GEMMs of the earlier layers; the ~30 small tensors (BatchNorm affine parameters,
  shrink layer) travel as one flat bucket at the end. BatchNorm batch statistics stay per replica by default (like
  torch DDP without SyncBatchNorm); `enable_sync_bn` all-reduces the per-channel sums instead (18 tiny latency-bound
  exchanges per step). `broadcast_buffers` makes running statistics identical before a checkpoint.
"""
import torch
import torch.distributed as dist

from . import training

import os

# gradients of at least this many bytes are all-reduced on their own as soon as they exist; smaller ones share one flat
# bucket at the end of the backward (VP3D_DDP_LARGE_BYTES overrides, e.g. a huge value = one exchange after the backward)
LARGE_BYTES = int(os.environ.get('VP3D_DDP_LARGE_BYTES', 1 << 20))


def shard_range(n_items, rank, world):
    """Contiguous, balanced [lo, hi) slice of n_items for `rank` (first n_items % world ranks get one extra)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class GradSync:
    """Averages parameter gradients over the ranks of `group`, overlapped with the backward pass.
    `compress='bf16'`: gradients travel as bfloat16 (half the bytes on the wire; the sum over ranks is taken in bf16, a
    relative error of ~2^-9 per gradient element -- far below the ~5e-2 that 16-bit GEMM operands already put on these
    gradients) and are written back into the fp32 gradient tensors the optimiser reads."""

    def __init__(self, group=None, compress=None):
        assert compress in (None, 'bf16')
        self.compress = compress
        self.group = group
        self.world = dist.get_world_size(group)
        backend = dist.get_backend(group)
        self._avg = dist.ReduceOp.AVG if backend == 'nccl' else None   # gloo has no AVG: SUM then scale
        self._pending = []
        self._small = []
        self.bytes_reduced = 0
        self.collectives = 0

    # -- hook protocol used by training._StackTrainFn.backward ---------------------------------------------------
    def __call__(self, param, grad):
        if self.world == 1:
            return
        if grad.numel() * grad.element_size() >= LARGE_BYTES:
            self._launch(grad)
        else:
            self._small.append(grad)

    def _launch(self, t):
        op = self._avg if self._avg is not None else dist.ReduceOp.SUM
        wire = t.to(torch.bfloat16) if (self.compress == 'bf16' and t.dtype == torch.float32) else t
        work = dist.all_reduce(wire, op=op, group=self.group, async_op=True)
        self._pending.append((work, t, wire))
        self.bytes_reduced += wire.numel() * wire.element_size()
        self.collectives += 1

    def finish(self):
        """Flushes the small-tensor bucket and makes the current stream wait for every outstanding all-reduce."""
        if self.world == 1:
            return
        flat = None
        if self._small:
            flat = torch.cat([g.reshape(-1) for g in self._small])
            self._launch(flat)
        for work, t, wire in self._pending:
            work.wait()
            if wire is not t:
                t.copy_(wire)
            if self._avg is None:
                t.div_(self.world)
        if flat is not None:
            off = 0
            for g in self._small:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n
        self._pending, self._small = [], []


_active = None


def enable_grad_sync(group=None, reserve_sms=0, dynamic_schedule=None, compress=None):
    """Install gradient averaging for every vp3d_b200 training backward of this process. Returns the GradSync.
    `dynamic_schedule` (default off; env VP3D_DDP_SCHED=dynamic turns it on): the big GEMMs of the backward take their
    tiles dynamically (cluster launch control, vp3d_set_sched_mode) instead of walking a static persistent schedule, so
    SMs that NCCL's all-reduce CTAs occupy only shrink the pool of workers instead of leaving a fixed share of the tiles
    waiting behind them. Measured at 2 GPUs (profiles/README.md): within the box-to-box noise of the static schedule
    (1.96-2.01 vs 1.96-1.99 ms per step), so it is an option, not the default.
    `compress='bf16'` (env VP3D_DDP_COMPRESS=bf16): see GradSync.
    `reserve_sms` > 0 sizes the persistent GEMM grids for that many SMs fewer than the device has, so that NCCL's
    all-reduce CTAs (cap them with NCCL_MAX_CTAS <= reserve_sms before the process group is created) run beside the
    backward GEMMs instead of taking turns with them."""
    global _active
    if reserve_sms > 0 and torch.cuda.is_available():
        from . import native
        sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
        native.check(native.lib().vp3d_set_sm_limit(int(sms - reserve_sms)), 'set_sm_limit')
    if compress is None:
        compress = os.environ.get('VP3D_DDP_COMPRESS') or None
    _active = GradSync(group, compress=compress)
    if dynamic_schedule is None:
        dynamic_schedule = _active.world > 1 and os.environ.get('VP3D_DDP_SCHED', 'static') == 'dynamic'
    if torch.cuda.is_available():
        from . import native
        native.check(native.lib().vp3d_set_sched_mode(1 if dynamic_schedule else 0), 'set_sched_mode')
    training.grad_ready_hook = _active
    training.grad_finish_hook = _active.finish
    return _active


def disable_grad_sync():
    global _active
    _active = None
    if torch.cuda.is_available():
        from . import native
        native.check(native.lib().vp3d_set_sm_limit(0), 'set_sm_limit')
        native.check(native.lib().vp3d_set_sched_mode(0), 'set_sched_mode')
    training.grad_ready_hook = None
    training.grad_finish_hook = None


def enable_sync_bn(group=None, on=True):
    """Train-mode BatchNorm statistics over the global batch (all ranks) instead of per replica."""
    training.sync_bn_group = (group if group is not None else dist.group.WORLD) if on else None


def broadcast_buffers(model, src=0, group=None):
    """Running statistics / num_batches_tracked of rank `src` to every rank (call before saving a checkpoint)."""
    for b in model.buffers():
        dist.broadcast(b, src=src, group=group)


def broadcast_parameters(model, src=0, group=None):
    for p in model.parameters():
        dist.broadcast(p.data, src=src, group=group)
