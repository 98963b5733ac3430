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


class PeerGradSync:
    """Gradient averaging through NVLink / NVSwitch peer memory with the library's own kernel
    (vp3d_peer_allreduce_f32, csrc/allreduce.cu) instead of NCCL.

    All parameter gradients of a replica live in ONE symmetric fp32 buffer (torch.distributed._symmetric_memory: CUDA
    VMM allocations every rank of the node maps, plus the NVSwitch multicast address when the fabric has one). The
    backward writes each gradient straight into its slot (training.grad_alloc); the moment a large gradient has been
    issued its slot is all-reduced in place on a communication stream by a kernel of `ctas` CTAs with no shared memory
    -- two-shot over multimem.ld_reduce / multimem.st, or plain peer loads / stores without multicast -- while the
    persistent GEMMs of the earlier layers run on grids sized `ctas` SMs smaller (vp3d_set_sm_limit for the duration of
    the backward). The small tensors (BatchNorm affine, shrink layer) are adjacent in the buffer and travel as one
    slice at the end. Everything is a kernel launch on a stream: the step stays capturable as one CUDA graph.
    One rank computes each element and all ranks receive the same bits, so replicas stay bit-identical."""

    FLAG_FLOATS = 64 * 16          # kArMaxCtas x kArMaxRanks flag words at the start of every rank's buffer
    ALIGN = 32                     # slots start on 128-byte boundaries

    def __init__(self, params, group=None, ctas=None, reserve_sms=None, use_multicast=None, timeout_s=20.0):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import native
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.compress = None
        params = [p for p in params if p.requires_grad]
        assert params and all(p.is_cuda and p.dtype == torch.float32 for p in params), 'fp32 CUDA parameters'
        dev = params[0].device
        # CTAs of the exchange kernel (= SMs the backward's GEMM grids leave free). Measured (profiles/README.md): through
        # the switch 4 CTAs already move 67.8 MB in 0.16 ms at 8 GPUs (a rank reduces 1/8 of a slice) and the step is
        # 1.91 ms with 4 against 1.97 with 8; at 2 GPUs a rank reduces half of every slice and 8 CTAs are worth their SMs
        self.ctas = int(ctas if ctas is not None else os.environ.get('VP3D_DDP_CTAS', 8 if self.world <= 2 else 4))
        assert 2 <= self.ctas <= 64 and self.ctas % 2 == 0
        self.reserve = int(reserve_sms if reserve_sms is not None else os.environ.get('VP3D_DDP_RESERVE', self.ctas))
        up = lambda n: (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.slots = {}
        off = self.FLAG_FLOATS
        large = [p for p in params if p.numel() * 4 >= LARGE_BYTES]
        small = [p for p in params if p.numel() * 4 < LARGE_BYTES]
        for p in large:
            self.slots[id(p)] = (off, p.numel())
            off += up(p.numel())
        self.small_range = (off, 0)
        for p in small:
            self.slots[id(p)] = (off, p.numel())
            off += up(p.numel())
        self.small_range = (self.small_range[0], off - self.small_range[0])
        self.large_ids = {id(p) for p in large}
        self.total = off
        with torch.cuda.device(dev):
            self.buf = symm.empty(self.total, dtype=torch.float32, device=dev)
            self.handle = symm.rendezvous(self.buf, self.group.group_name)
            self.buf.zero_()
            torch.cuda.synchronize()
        dist.barrier(self.group)
        base_off = int(getattr(self.handle, 'offset', 0) or 0)
        ptrs = [int(a) + base_off for a in self.handle.buffer_ptrs]
        assert ptrs[self.rank] == self.buf.data_ptr(), 'symmetric buffer of this rank is not the tensor rendezvous()ed'
        self._peers = (C.c_void_p * self.world)(*ptrs)
        mc = int(getattr(self.handle, 'multicast_ptr', 0) or 0)
        if use_multicast is None:
            use_multicast = os.environ.get('VP3D_DDP_MULTICAST', '1') != '0'
        self.multicast = (mc + base_off) if (mc and use_multicast) else None
        self.timeout_s = float(timeout_s)
        self.comm = torch.cuda.Stream(dev)
        self._native = native
        self._dev = dev
        self.bytes_reduced = 0
        self.collectives = 0

    # -- training hooks ------------------------------------------------------------------------------------------
    def alloc(self, param):
        s = self.slots.get(id(param))
        if s is None:
            return None
        slot = self.buf[s[0]:s[0] + s[1]]
        if param.grad is not None and param.grad.data_ptr() == slot.data_ptr():
            # the previous step's gradient is still attached and lives in this slot (zero_grad(set_to_none=False),
            # gradient accumulation): move it out, autograd is going to add this step's gradient to it
            param.grad = param.grad.clone()
        return slot

    def begin(self):
        if self.reserve > 0:
            sms = torch.cuda.get_device_properties(self._dev).multi_processor_count
            self._native.check(self._native.lib().vp3d_set_sm_limit(int(sms - self.reserve)), 'set_sm_limit')

    def _exchange(self, off, n):
        """All-reduce of buf[off : off + n) on the communication stream, after everything issued so far on the current one."""
        cur = torch.cuda.current_stream(self._dev)
        ev = torch.cuda.Event()
        ev.record(cur)
        n4 = (n + 3) // 4 * 4
        a = self._native.AllReduceArgs()
        a.multicast = self.multicast
        a.peers = self._peers
        a.flags = self._peers                 # the flag words are the head of every buffer
        a.rank, a.world, a.offset, a.count = self.rank, self.world, off, n4
        a.scale, a.ctas, a.timeout_s = 1.0 / self.world, self.ctas, self.timeout_s
        import ctypes as C
        with torch.cuda.device(self._dev):
            self.comm.wait_event(ev)
            self._native.check(self._native.lib().vp3d_peer_allreduce_f32(C.byref(a), self.comm.cuda_stream),
                               'peer_allreduce')
        self.bytes_reduced += n4 * 4
        self.collectives += 1

    def __call__(self, param, grad):
        s = self.slots.get(id(param))
        if s is None:
            raise RuntimeError('vp3d_b200.ddp: gradient of a parameter that was not registered with PeerGradSync')
        off, n = s
        slot = self.buf[off:off + n]
        if grad.data_ptr() != slot.data_ptr():
            slot.copy_(grad.reshape(-1))       # a producer that did not take the slot (generic path): one copy
        if id(param) in self.large_ids:
            self._exchange(off, n)

    def finish(self):
        self._exchange(*self.small_range)
        torch.cuda.current_stream(self._dev).wait_stream(self.comm)
        if self.reserve > 0:
            self._native.check(self._native.lib().vp3d_set_sm_limit(0), 'set_sm_limit')
        return None


_active = None


def enable_grad_sync(group=None, reserve_sms=0, dynamic_schedule=None, compress=None, exchange=None, params=None):
    """Install gradient averaging for every vp3d_b200 training backward of this process. Returns the GradSync.
    `exchange`: 'nccl' = dist.all_reduce per gradient (GradSync); 'peer' = the library's own all-reduce kernel over
    NVLink peer memory (PeerGradSync; needs `params`, CUDA, one node); default: env VP3D_DDP_EXCHANGE, else 'peer' when
    `params` are given and symmetric memory can be set up, else 'nccl'.
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
    if exchange is None:
        exchange = 'nccl' if params is None else (os.environ.get('VP3D_DDP_EXCHANGE') or 'peer')
    assert exchange in ('nccl', 'peer')
    world = dist.get_world_size(group)
    if exchange == 'peer' and world > 1 and torch.cuda.is_available() and dist.get_backend(group) == 'nccl':
        if params is None:
            raise ValueError("enable_grad_sync(exchange='peer') needs params=model.parameters()")
        try:
            _active = PeerGradSync(list(params), group, reserve_sms=reserve_sms or None)
        except Exception as e:          # no symmetric memory on this system (no P2P / fabric): NCCL carries the exchange
            if os.environ.get('VP3D_DDP_EXCHANGE') == 'peer':
                raise
            import warnings
            warnings.warn('vp3d_b200.ddp: peer-memory gradient exchange unavailable (%s); using NCCL' % (e,))
            _active = None
        if _active is not None:
            from . import native
            native.check(native.lib().vp3d_set_sched_mode(0), 'set_sched_mode')
            training.grad_ready_hook = _active
            training.grad_finish_hook = _active.finish
            training.grad_begin_hook = _active.begin
            training.grad_alloc = _active.alloc
            # an optimiser update inside the backward (FusedAdam.update_in_backward) follows the exchange of its gradient
            # on the communication stream; the small tensors are exchanged at the end and left to optimizer.step()
            training.update_stream = lambda sync=_active: sync.comm
            training.update_filter = lambda q, sync=_active: id(q) in sync.large_ids
            return _active
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
    training.grad_begin_hook = None
    training.grad_alloc = None
    training.update_stream = None
    training.update_filter = (lambda q: False) if _active.world > 1 else None   # NCCL: updates wait for finish()
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
    training.grad_begin_hook = None
    training.grad_alloc = None
    training.update_stream = None
    training.update_filter = None


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
