"""Streaming inference for causal TemporalModel (BASELINE configs[3]): S concurrent streams, one new frame per step.

The reference has no streaming mode; it re-runs the fully-convolutional model over the whole padded sequence
(run.py:697-711, generators.py:178-205). A causal model's output at frame t only needs, per layer, the layer's input
at t, t-d and t-2d (d = 1, 3, 9, 27, 81), so each layer keeps a ring of its last 2d+1 input frames and one step is ten
M = S GEMMs (34.7 GFLOP for S = 1024) instead of a 243-frame window per output.

Ring layout: ring[i] is [2*L_i slots][S_pad rows][C] in the operand type with L_i = 2*d_i + 1 and S_pad = S rounded up
to the 128-row GEMM tile; frame t lives in slot (t mod L_i) AND in slot (t mod L_i) + L_i. With the mirror, the three
taps of a step are always the arithmetic row progression q-2d, q-d, q of one flat [slot*S_pad + stream] view
(q = (t mod L) + L), which is exactly what one vp3d_conv_block_fwd launch with taps = 3, tap_row_step = d*S_pad
consumes -- no wrap-around case, no gather kernel; the producing GEMM stores every output box twice (slot and mirror).

The ring positions live on the DEVICE: vp3d_stream_advance ticks a frame counter and rewrites a small offsets table
that the GEMM launches read (vp3d_conv_args.dyn_offsets), so the ~15 launches of a step never change and are captured
once as a CUDA graph; a frame then costs one graph launch instead of ~20 Python-issued launches.

Few streams (S <= 8): a frame is ten matrix-vector products, and twelve dependent launches cost 0.13 ms whatever S is.
`vp3d_stream_step_fused` (csrc/stream.cu) runs the whole frame -- ring bookkeeping, ring write, all layers -- as ONE
cooperative kernel with grid barriers between the layers (one warp per output channel, weights streamed from L2), on the
same rings and packed operands; VP3D_STREAM_FUSED=0 keeps the GEMM launches.

Per-frame camera: `step_world()` takes world-space joints plus this frame's camera (quaternion, translation, intrinsics
with distortion) per stream and projects on the device (vp3d_project_points) before the stack.
"""
import ctypes as C
import os

import torch

from . import native, ops
from .temporal import N_TILE, packed_for, resolve_dtype

FUSED_MAX_STREAMS = 8     # vp3d_stream_step_fused serves 1..8 streams


class CausalStream:
    def __init__(self, model, n_streams, dtype=None, use_graph=True):
        assert not model._strided, 'streaming uses the dilated TemporalModel (weights are interchangeable with the 1f model)'
        assert all(s > 0 for s in model.causal_shift[1:]) or len(model.filter_widths) == 1, \
            'streaming needs TemporalModel(causal=True)'
        assert not model.training, 'streaming runs the eval-mode (folded BatchNorm) path'
        self.model = model
        self.S = int(n_streams)
        self.S_pad = (self.S + 127) // 128 * 128
        self.dt = resolve_dtype(dtype or getattr(model, 'operand_dtype', None))
        if self.dt == native.TF32:
            raise RuntimeError('vp3d_b200 streaming runs fp16 / bf16 operands')
        self.pk = pk = packed_for(model, self.dt)
        self.dev = dev = model.expand_conv.weight.device
        fw = model.filter_widths
        nb = len(fw) - 1
        self.taps = [fw[0]] + [model.layers_conv[2 * i].kernel_size[0] for i in range(nb)]
        self.dil = [1] + [model.layers_conv[2 * i].dilation[0] for i in range(nb)]
        self.L = [(self.taps[i] - 1) * self.dil[i] + 1 for i in range(nb + 1)]
        td = ops.torch_dtype(self.dt)
        widths = [pk.c_in_pad] + [pk.c_pad] * nb
        # ring i holds the INPUT of convolution i (i = 0: packed 2-D keypoints; i >= 1: output of the previous block)
        self.rings = [torch.zeros((2 * self.L[i], self.S_pad, widths[i]), dtype=td, device=dev) for i in range(nb + 1)]
        i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
        self.ring_len, self.ring_dil, self.ring_taps = i32(self.L), i32(self.dil), i32(self.taps)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.table = torch.zeros((nb + 1, 4), dtype=torch.int32, device=dev)
        # GEMM launches of one step: (A ring, residual ring, output ring); -1 = a plain buffer
        desc = [(0, -1, 1 if nb else -1)]
        for i in range(1, nb + 1):
            desc += [(i, -1, -1), (-1, i, i + 1 if i < nb else -1)]
        self.launch_desc = i32(desc)
        self.launch_table = torch.zeros((len(desc), 4), dtype=torch.int32, device=dev)
        C_ = pk.c_pad
        self.x_in = torch.zeros((self.S, pk.c_in), dtype=torch.float32, device=dev)
        self.y1 = torch.empty((self.S_pad, C_), dtype=td, device=dev)
        self.h_last = torch.empty((self.S_pad, C_), dtype=td, device=dev)
        self.y = torch.empty((self.S, pk.n_out), dtype=torch.float32, device=dev)
        # small M: 64-wide column tiles give 4x more CTAs than the 256-wide tile of the batch path
        self.block_n = 64 if (self.S_pad // 128) * (C_ // N_TILE) * 4 <= native.sm_count(dev) else N_TILE
        self.use_graph = use_graph
        self.graph = None
        self._warm = 0
        # few streams: the whole frame as one cooperative kernel (no graph needed: it is a single launch)
        self.fused = (self.S <= FUSED_MAX_STREAMS and os.environ.get('VP3D_STREAM_FUSED', '1') != '0' and
                      max(self.taps[i] * widths[i] for i in range(nb + 1)) <= 3072 and nb + 2 + nb <= 12)
        self.barrier = torch.zeros(128, dtype=torch.int64, device=dev)    # grid-barrier counters, 0 whenever step_dev is 0
        self._fused_layers = None

    def refresh_weights(self):
        """Re-resolves the packed eval operands (folded BatchNorm) from the module's CURRENT parameters and buffers. The
        pack is looked up once per stream object because checking ~60 tensor versions would cost a fifth of a 0.13 ms
        frame; call this (or reset()) after load_state_dict / further training of the module. A captured graph that
        baked the old operands' addresses is dropped and re-captured."""
        pk = packed_for(self.model, self.dt)
        if pk is not self.pk:
            self.pk = pk
            self.graph = None
            self._warm = 0
            self._fused_layers = None

    def reset(self):
        self.refresh_weights()
        for r in self.rings:
            r.zero_()
        self.step_dev.zero_()
        self.barrier.zero_()
        
    def prime(self, x0):
        """Left edge padding of the reference's generator (generators.py:193-195 replicates the first frame over the
        receptive field): feed frame 0 RF-1 times so that the first real step sees that history."""
        for _ in range(self.model.receptive_field() - 1):
            self.step(x0)

    # ------------------------------------------------------------------------------------------------------------
    def _issue(self):
        """All launches of one frame, reading self.x_in and writing self.y. Nothing here depends on the frame index on
        the host: ring positions come from the device table."""
        pk, S, Sp, dt, dev = self.pk, self.S, self.S_pad, self.dt, self.dev
        lib = native.lib()
        nb = len(self.model.filter_widths) - 1
        C_ = pk.c_pad
        with torch.cuda.device(dev):
            native.check(lib.vp3d_stream_advance(self.step_dev.data_ptr(), nb + 1, self.ring_len.data_ptr(),
                                                 self.ring_dil.data_ptr(), self.ring_taps.data_ptr(), Sp,
                                                 self.table.data_ptr(), self.launch_table.shape[0],
                                                 self.launch_desc.data_ptr(), self.launch_table.data_ptr(),
                                                 ops._stream()), 'stream_advance')
            native.check(lib.vp3d_ring_write(dt, self.x_in.data_ptr(), self.rings[0].data_ptr(),
                                             self.table[0].data_ptr(), S, pk.c_in, pk.c_in_pad, ops._stream()),
                         'ring_write')
        launch = 0

        def gemm(a, a_rows, k_pad, w, taps, dil, shift, out, out_rows_total, res=None, relu=True):
            nonlocal launch
            kw = {}
            if res is not None:
                kw = dict(res=res, res_view=(C_, 0, 1, 0))
            ops.conv_block(dt, a, (1, a_rows, k_pad, k_pad, a_rows * k_pad), w, taps, dil * Sp, k_pad, S, out,
                           (C_, out_rows_total * C_), block_n=self.block_n, scale=None, shift=shift, relu=relu,
                           dyn_offsets=self.launch_table[launch], out_rows_total=out_rows_total, **kw)
            launch += 1

        ring_rows = lambda i: self.rings[i].shape[0] * Sp
        out0 = self.rings[1] if nb else self.h_last
        gemm(self.rings[0], ring_rows(0), pk.c_in_pad, pk.w_expand, self.taps[0], self.dil[0], pk.bn_expand[1], out0,
             ring_rows(1) if nb else Sp)
        for i in range(1, nb + 1):
            gemm(self.rings[i], ring_rows(i), C_, pk.w_layers[2 * (i - 1)], self.taps[i], self.dil[i],
                 pk.bn_layers[2 * (i - 1)][1], self.y1, Sp)
            last = i == nb
            out = self.h_last if last else self.rings[i + 1]
            gemm(self.y1, Sp, C_, pk.w_layers[2 * (i - 1) + 1], 1, 0, pk.bn_layers[2 * (i - 1) + 1][1], out,
                 Sp if last else ring_rows(i + 1), res=self.rings[i])
        ops.conv_block(dt, self.h_last, (1, Sp, C_, C_, Sp * C_), pk.w_shrink, 1, 0, C_, S, self.y,
                       (pk.n_out, S * pk.n_out), block_n=64, scale=None, shift=pk.shrink_shift, relu=False, out_f32=True,
                       n_valid=pk.n_out)

    def _fused_desc(self):
        """vp3d_stream_layer array of one frame: the launches of _issue() as layer descriptions (built once per pack)."""
        if self._fused_layers is not None:
            return self._fused_layers
        pk, Sp = self.pk, self.S_pad
        nb = len(self.model.filter_widths) - 1
        C_ = pk.c_pad
        layers = []

        def layer(a, a_ring, k_pad, w, taps, dil, shift, out, out_ring, res=None, res_ring=-1, relu=True, out_f32=False,
                  n=C_, n_valid=C_, out_stride=C_):
            layers.append(native.StreamLayer(
                a=a.data_ptr(), w=w.data_ptr(), shift=shift.data_ptr(), res=res.data_ptr() if res is not None else None,
                out=out.data_ptr(), a_ring=a_ring, res_ring=res_ring, out_ring=out_ring, k_per_tap=k_pad, taps=taps,
                tap_row_step=dil * Sp, n=n, n_valid=n_valid, relu=1 if relu else 0, out_f32=1 if out_f32 else 0,
                res_row_stride=C_, out_row_stride=out_stride))

        layer(self.rings[0], 0, pk.c_in_pad, pk.w_expand, self.taps[0], self.dil[0], pk.bn_expand[1],
              self.rings[1] if nb else self.h_last, 1 if nb else -1)
        for i in range(1, nb + 1):
            layer(self.rings[i], i, C_, pk.w_layers[2 * (i - 1)], self.taps[i], self.dil[i], pk.bn_layers[2 * (i - 1)][1],
                  self.y1, -1)
            last = i == nb
            layer(self.y1, -1, C_, pk.w_layers[2 * (i - 1) + 1], 1, 0, pk.bn_layers[2 * (i - 1) + 1][1],
                  self.h_last if last else self.rings[i + 1], -1 if last else i + 1, res=self.rings[i], res_ring=i)
        layer(self.h_last, -1, C_, pk.w_shrink, 1, 0, pk.shrink_shift, self.y, -1, relu=False, out_f32=True,
              n=pk.w_shrink.shape[0], n_valid=pk.n_out, out_stride=pk.n_out)
        arr = (native.StreamLayer * len(layers))(*layers)
        self._fused_layers = (arr, len(layers))
        return self._fused_layers

    def _issue_fused(self, x_in):
        """One cooperative kernel for the whole frame (S <= 8). x_in: (S, c_in) fp32 contiguous CUDA."""
        pk = self.pk
        arr, n_layers = self._fused_desc()
        with torch.cuda.device(self.dev):
            native.check(native.lib().vp3d_stream_step_fused(
                self.dt, self.step_dev.data_ptr(), self.barrier.data_ptr(), len(self.rings), self.ring_len.data_ptr(),
                self.ring_dil.data_ptr(), self.ring_taps.data_ptr(), self.S_pad, x_in.data_ptr(), pk.c_in, pk.c_in_pad,
                self.rings[0].data_ptr(), self.S, arr, n_layers, ops._stream()), 'stream_step_fused')

    def step(self, x_t):
        """x_t: (S, J, F) fp32 CUDA, the frame every stream has just received -> (S, J_out, 3) fp32 (a fresh tensor)."""
        ops.require_cuda(x_t)
        m, pk, S = self.model, self.pk, self.S
        assert x_t.shape[0] == S and x_t.numel() == S * pk.c_in
        if self.fused:
            self._issue_fused(ops.f32c(x_t).reshape(S, pk.c_in))
            return self.y.view(S, m.num_joints_out, 3).clone()
        self.x_in.copy_(x_t.reshape(S, pk.c_in))
        if not self.use_graph:
            self._issue()
        elif self.graph is None:
            self._issue()                       # first frames run eagerly (lazy initialisations, attribute setup) ...
            self._warm += 1
            if self._warm >= 2:                 # ... then the step is captured once
                torch.cuda.synchronize(self.dev)
                g = torch.cuda.CUDAGraph()
                step_before = self.step_dev.clone()
                with torch.cuda.graph(g):
                    self._issue()
                # capture does not execute: restore nothing, but make sure the counter was not advanced by the capture
                assert torch.equal(step_before, self.step_dev)
                self.graph = g
        else:
            self.graph.replay()
        return self.y.view(S, m.num_joints_out, 3).clone()

    def step_world(self, X_world, q, t, camera_params):
        """X_world (S, J, 3) world-space joints of this frame, q (S, 4) / t (S, 3) this frame's camera pose per stream,
        camera_params (S, 9) intrinsics incl. distortion -> (S, J_out, 3). world -> camera -> image on the device."""
        J = X_world.shape[1]
        mode = native.PT_WORLD_TO_CAMERA | native.PT_PROJECT | native.PT_FAST
        _, x2d = ops.project_points(X_world, q=q, t=t, cam=camera_params, pts_per_q=J, pts_per_cam=J, mode=mode, want2=True)
        return self.step(x2d)
