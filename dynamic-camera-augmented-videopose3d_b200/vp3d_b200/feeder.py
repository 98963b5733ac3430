"""GPU-resident batch feeder: ChunkedGenerator's batch assembly (common/generators.py:11-137) fused with
the dynamic-camera projection, on the device.

The reference assembles every training batch on the host: a Python loop over the 1024 samples, np.pad edge padding,
a 3x3 @ (243,3,4) matmul per sample, float64 staging buffers, then a blocking .cuda() (generators.py:102-132,
run.py:458-464) -- 78 ms per batch, the real bottleneck of its training step. Here all sequences live on
the device once (world-space joints + one camera pose per frame), an epoch is the reference's own permutation of
(sequence, frame) pairs, and a batch costs one 12 KB index upload plus one kernel (vp3d_project_windows) that gathers
the padded windows, applies world -> camera -> image per frame and writes the (B, window, J, 2) batch, the
camera-space target and (optionally) the K @ [R|t] matrices directly.

Epoch order: `np.random.RandomState(seed).permutation(pairs)` exactly as generators.py:56,85, so sample order is
reproducible against the reference generator. Deviation (deliberate): the last, partial batch of an epoch contains
only its own samples -- the reference yields its full-size buffers with stale rows from the previous batch.
"""
import ctypes as C

import numpy as np
import torch

from . import native, ops


def build_pairs(lengths, chunk_length):
    """(n_chunks, 3) int64 table of (sequence, first frame, one-past-last frame). A sequence of n frames is cut into
    ceil(n / chunk_length) chunks centred on it (the cut overhangs both ends by half of the excess), which is the
    lineage rule of generators.py:39-45; the table is what RandomState.permutation shuffles row-wise each epoch."""
    tables = []
    for seq, n in enumerate(lengths):
        k = -(-int(n) // chunk_length)
        first = np.arange(k, dtype=np.int64) * chunk_length - (k * chunk_length - int(n)) // 2
        tables.append(np.stack([np.full(k, seq, dtype=np.int64), first, first + chunk_length], axis=1))
    return np.concatenate(tables) if tables else np.zeros((0, 3), dtype=np.int64)


class DeviceWindowFeeder:
    def __init__(self, world_3d, quats, trans, intrinsics, batch_size, chunk_length=1, pad=0, causal_shift=0,
                 shuffle=True, random_seed=1234, root_relative=True, linear=False, want_cameras=False, device='cuda',
                 endless=False):
        """world_3d: list of (T_i, J, 3) arrays/tensors; quats (T_i, 4) and trans (T_i, 3): one camera pose per frame;
        intrinsics: (n_seq, 9) [fx, fy, cx, cy, k1, k2, k3, p1, p2] per sequence."""
        assert len(world_3d) == len(quats) == len(trans) == len(intrinsics)
        dev = torch.device(device)
        as_t = lambda a: torch.as_tensor(np.asarray(a.cpu() if isinstance(a, torch.Tensor) else a), dtype=torch.float32)
        self.joints = int(world_3d[0].shape[-2])
        lens = [int(x.shape[0]) for x in world_3d]
        for x, q, t in zip(world_3d, quats, trans):
            assert x.shape[0] == q.shape[0] == t.shape[0] and x.shape[-1] == 3 and q.shape[-1] == 4 and t.shape[-1] == 3
        self.x = torch.cat([as_t(x) for x in world_3d]).contiguous().to(dev)
        self.q = torch.cat([as_t(q) for q in quats]).contiguous().to(dev)
        self.t = torch.cat([as_t(t) for t in trans]).contiguous().to(dev)
        self.cam = as_t(intrinsics).reshape(len(lens), 9).contiguous().to(dev)
        starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
        self.seq_start = torch.from_numpy(starts).to(dev)
        self.seq_len = torch.tensor(lens, dtype=torch.int64, device=dev)
        self.pairs = pairs = build_pairs(lens, chunk_length)
        self.batch_size, self.chunk_length, self.pad, self.causal_shift = batch_size, chunk_length, pad, causal_shift
        self.num_batches = (len(pairs) + batch_size - 1) // batch_size
        self.random = np.random.RandomState(random_seed)
        self.shuffle, self.endless = shuffle, endless
        self._resume = None      # (next batch, this epoch's order) of an interrupted endless epoch
        self.root_relative, self.linear, self.want_cameras = root_relative, linear, want_cameras
        self.dev = dev
        self.window = chunk_length + 2 * pad
        # pinned staging for the per-batch index upload (two buffers: the copy of batch i+1 may overlap batch i)
        self._idx_host = [torch.empty((batch_size, 2), dtype=torch.int64).pin_memory() for _ in range(2)]
        self._idx_dev = [torch.empty((batch_size, 2), dtype=torch.int64, device=dev) for _ in range(2)]
        self._idx_copied = [None, None]   # event recorded behind the H2D copy that last read each pinned buffer
        self._flip = 0
        self._bound = None       # (batch_2d, batch_3d) buffers every full batch is written into (bind_outputs)

    def bind_outputs(self, batch_2d=None, batch_3d=None):
        """Write every full batch into these caller-owned buffers instead of fresh tensors -- e.g. the static inputs of a
        captured training step (vp3d_b200.graphs.GraphedTrainStep), so that the kernel's output IS the graph's input and
        no device-to-device copy sits between the feeder and the step. Shapes (batch_size, window, J, 2) and
        (batch_size, chunk_length, J, 3), fp32, contiguous, on the feeder's device; None unbinds."""
        if batch_2d is None:
            self._bound = None
            return self
        J = self.joints
        want2, want3 = (self.batch_size, self.window, J, 2), (self.batch_size, self.chunk_length, J, 3)
        for t, want in ((batch_2d, want2), (batch_3d, want3)):
            assert tuple(t.shape) == want and t.dtype == torch.float32 and t.is_contiguous() and t.device == torch.device(
                self.dev), 'bind_outputs: expected a contiguous fp32 %s on %s' % (want, self.dev)
        self._bound = (batch_2d, batch_3d)
        return self

    # -- generator protocol of the reference ------------------------------------------------------------------
    def num_frames(self):
        return self.num_batches * self.batch_size

    def random_state(self):
        return self.random

    def set_random_state(self, random):
        self.random = random

    def epoch_order(self):
        """(first batch, chunk table in this epoch's order): a fresh draw from the RandomState (generators.py:85), or
        the remainder of an endless epoch that was interrupted (checkpoint / resume, run.py:436-445)."""
        if self._resume is not None:
            return self._resume
        return 0, (self.random.permutation(self.pairs) if self.shuffle else self.pairs)

    next_pairs = epoch_order     # the reference generator's name for it

    def assemble(self, chunks):
        """chunks: (n, 3) int array of (seq_i, start_3d, end_3d) -> (cams or None, batch_3d, batch_2d) on the device."""
        chunks = np.asarray(chunks)
        n = len(chunks)
        b = self._flip
        self._flip ^= 1
        host = self._idx_host[b]
        if self._idx_copied[b] is not None:
            # a host that runs batches ahead of the GPU must not overwrite indices a pending copy still has to read
            self._idx_copied[b].synchronize()
        host[:n] = torch.from_numpy(np.ascontiguousarray(chunks[:, :2], dtype=np.int64))
        idx = self._idx_dev[b]
        idx[:n].copy_(host[:n], non_blocking=True)
        ev = self._idx_copied[b] = self._idx_copied[b] or torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        seq32 = idx[:n, 0].to(torch.int32).contiguous()
        start = idx[:n, 1].contiguous()
        J = self.joints
        if self._bound is not None and n == self.batch_size:
            x2d, tgt = self._bound
        else:
            x2d = torch.empty((n, self.window, J, 2), dtype=torch.float32, device=self.dev)
            tgt = torch.empty((n, self.chunk_length, J, 3), dtype=torch.float32, device=self.dev)
        cams = torch.empty((n, self.window, 3, 4), dtype=torch.float32, device=self.dev) if self.want_cameras else None
        a = native.WindowArgs()
        a.x_world, a.q, a.t, a.cam = self.x.data_ptr(), self.q.data_ptr(), self.t.data_ptr(), self.cam.data_ptr()
        a.seq_start, a.seq_len = self.seq_start.data_ptr(), self.seq_len.data_ptr()
        a.sample_seq, a.sample_start = seq32.data_ptr(), start.data_ptr()
        a.batch, a.joints, a.chunk_length, a.pad, a.causal_shift = n, J, self.chunk_length, self.pad, self.causal_shift
        a.root_relative, a.linear = int(self.root_relative), int(self.linear)
        a.out2, a.target3 = x2d.data_ptr(), tgt.data_ptr()
        a.cam3x4 = None if cams is None else cams.data_ptr()
        with torch.cuda.device(self.dev):
            native.check(native.lib().vp3d_project_windows(C.byref(a), ops._stream()), 'project_windows')
        return cams, tgt, x2d

    def next_epoch(self):
        """Yields (batch_cam, batch_3d, batch_2d) like ChunkedGenerator.next_epoch (generators.py:102-132), as fp32 CUDA
        tensors (batch_cam is None unless want_cameras)."""
        bs = self.batch_size
        while True:
            first, order = self.epoch_order()
            for b in range(first, self.num_batches):
                if self.endless:
                    self._resume = (b + 1, order)
                yield self.assemble(order[b * bs:(b + 1) * bs])
            self._resume = None
            if not self.endless:
                return


class DeviceSequenceFeeder(DeviceWindowFeeder):
    """UnchunkedGenerator (common/generators.py:140-205) on the device: one whole sequence per iteration (batch 1), the
    2-D input edge-padded by (pad + causal_shift, pad - causal_shift) frames (:193-195) -- here the padded frames are
    *projected* edge frames, which is the same thing because padding repeats the edge frame's joints and camera --, the
    camera-space (root-relative) 3-D target of the un-padded frames, optionally the padded K @ [R|t] matrices (:184-198)
    and the sequence's camera-motion statistics passed through (`seq_info`, :200-205).

    The sequences stay resident in HBM; an iteration is one launch of vp3d_project_windows with chunk = sequence length,
    so the evaluation loop of run.py:697-705 has no host-side padding, no float64 staging and no H2D copy per sequence."""

    def __init__(self, world_3d, quats, trans, intrinsics, pad=0, causal_shift=0, root_relative=True, linear=False,
                 want_cameras=False, seq_info=None, device='cuda'):
        super().__init__(world_3d, quats, trans, intrinsics, batch_size=1, chunk_length=1, pad=pad,
                         causal_shift=causal_shift, shuffle=False, root_relative=root_relative, linear=linear,
                         want_cameras=want_cameras, device=device)
        self.lengths = [int(x.shape[0]) for x in world_3d]
        self.seq_info = seq_info if seq_info is not None else [{} for _ in self.lengths]
        assert len(self.seq_info) == len(self.lengths)
        self.seq_length = 1 + 2 * pad     # attribute of the reference class (:170)

    def num_frames(self):
        return sum(self.lengths)

    def sequence(self, i):
        """(batch_cam or None, batch_3d (1, T, J, 3), batch_2d (1, T + 2 pad, J, 2)) of sequence i."""
        n = self.lengths[i]
        self.chunk_length, self.window = n, n + 2 * self.pad
        try:
            return self.assemble(np.array([[i, 0, n]], dtype=np.int64))
        finally:
            self.chunk_length, self.window = 1, 1 + 2 * self.pad

    def next_epoch(self):
        for i in range(len(self.lengths)):
            cams, b3d, b2d = self.sequence(i)
            yield cams, b3d, b2d, self.seq_info[i]
