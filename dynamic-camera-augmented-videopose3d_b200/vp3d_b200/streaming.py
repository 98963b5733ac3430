"""Streaming inference for causal TemporalModel (BASELINE configs[3]): S concurrent streams, one new frame per step.

The reference has no streaming mode; it re-runs the fully-convolutional model over the whole padded sequence
(run.py:697-711, generators.py:178-205). A causal model's output at frame t only needs, per layer, the layer's input
at t, t-d and t-2d (d = 1, 3, 9, 27, 81), so each layer keeps a ring of its last 2d+1 input frames and one step is ten
M = S GEMMs (34.7 GFLOP for S = 1024) instead of a 243-frame window per output.

Ring layout: ring[i] is [2*L_i slots][S streams][C] in the operand type with L_i = 2*d_i + 1; frame t lives in slot
(t mod L_i) AND in slot (t mod L_i) + L_i. With the mirror, the three taps of a step are always the arithmetic row
progression q-2d, q-d, q of one flat [slot*S + stream] view (q = (t mod L) + L), which is exactly what one
vp3d_conv_block_fwd launch with taps = 3, tap_row_step = d*S consumes -- no wrap-around case, no gather kernel.
The mirror copy is a device-to-device memcpy of S*C elements per layer per step.

Per-frame camera: `step_world()` takes world-space joints plus this frame's camera (quaternion, translation, intrinsics
with distortion) per stream and projects on the device (vp3d_project_points) before the stack.
"""
import torch

from . import native, ops
from .temporal import N_TILE, packed_for, resolve_dtype


class CausalStream:
    def __init__(self, model, n_streams, dtype=None):
        assert not model._strided, 'streaming uses the dilated TemporalModel (weights are interchangeable with the 1f model)'
        assert all(s > 0 for s in model.causal_shift[1:]) or len(model.filter_widths) == 1, \
            'streaming needs TemporalModel(causal=True)'
        assert not model.training, 'streaming runs the eval-mode (folded BatchNorm) path'
        self.model = model
        self.S = int(n_streams)
        self.dt = resolve_dtype(dtype or getattr(model, 'operand_dtype', None))
        if self.dt == native.TF32:
            raise RuntimeError('vp3d_b200 streaming runs fp16 / bf16 operands')
        self.pk = packed_for(model, self.dt)
        self.dev = model.expand_conv.weight.device
        fw = model.filter_widths
        self.taps = [fw[0]] + [model.layers_conv[2 * i].kernel_size[0] for i in range(len(fw) - 1)]
        self.dil = [1] + [model.layers_conv[2 * i].dilation[0] for i in range(len(fw) - 1)]
        td = ops.torch_dtype(self.dt)
        widths = [self.pk.c_in_pad] + [self.pk.c_pad] * (len(fw) - 1)
        # ring i holds the INPUT of convolution i (i = 0: packed 2-D keypoints; i >= 1: output of the previous block)
        self.L = [(self.taps[i] - 1) * self.dil[i] + 1 for i in range(len(fw))]
        self.rings = [torch.zeros((2 * self.L[i], self.S, widths[i]), dtype=td, device=self.dev) for i in range(len(fw))]
        self.t = 0
        # small M: 64-wide column tiles give 4x more CTAs than the 256-wide tile of the batch path
        self.block_n = 64 if (self.S + 127) // 128 * (self.pk.c_pad // N_TILE) < 148 else N_TILE

    def reset(self):
        for r in self.rings:
            r.zero_()
        self.t = 0

    def prime(self, x0):
        """Left edge padding of the reference's generator (generators.py:193-195 replicates the first frame over the
        receptive field): feed frame 0 RF-1 times so that the first real step sees that history."""
        for _ in range(self.model.receptive_field() - 1):
            self.step(x0)

    def _slot(self, i):
        return self.t % self.L[i] + self.L[i]

    def _conv(self, i, w, scale, shift, out, out_rows_view, k_pad, res=None, res_row=0):
        S, d, taps = self.S, self.dil[i], self.taps[i]
        ring = self.rings[i]
        q = self._slot(i)
        rows = ring.shape[0] * S
        kw = {}
        if res is not None:
            kw = dict(res=res, res_view=(res.shape[-1], 0, 1, res_row))
        ops.conv_block(self.dt, ring, (1, rows, k_pad, k_pad, rows * k_pad), w, taps, d * S, k_pad, S, out, out_rows_view,
                       block_n=self.block_n, a_row_off=(q - (taps - 1) * d) * S, scale=scale, shift=shift, relu=True, **kw)

    def step(self, x_t):
        """x_t: (S, J, F) fp32 CUDA, the frame every stream has just received -> (S, J_out, 3) fp32."""
        ops.require_cuda(x_t)
        m, pk, S, dt = self.model, self.pk, self.S, self.dt
        assert x_t.shape[0] == S and x_t.shape[1] * x_t.shape[2] == pk.c_in
        nb = len(m.filter_widths) - 1
        # frame t of the input ring (+ mirror)
        q0 = self._slot(0)
        xin = ops.pack_rows(dt, x_t.reshape(S, pk.c_in), pk.c_in_pad)
        self.rings[0][q0].copy_(xin)
        self.rings[0][q0 - self.L[0]].copy_(xin)
        td = ops.torch_dtype(dt)
        C = pk.c_pad
        h_last = None
        for i in range(nb + 1):
            last = i == nb
            # destination of this stage's output: slot of the next ring, or a plain buffer after the last block
            if not last:
                qn = self._slot(i + 1)
                dst = self.rings[i + 1][qn]
            else:
                dst = torch.empty((S, C), dtype=td, device=self.dev)
            if i == 0:
                self._conv(0, pk.w_expand, pk.bn_expand[0], pk.bn_expand[1], dst, (C, S * C), pk.c_in_pad)
            else:
                y1 = torch.empty((S, C), dtype=td, device=self.dev)
                bn3, bn1 = pk.bn_layers[2 * (i - 1)], pk.bn_layers[2 * (i - 1) + 1]
                self._conv(i, pk.w_layers[2 * (i - 1)], bn3[0], bn3[1], y1, (C, S * C), C)
                # 1x1 convolution + residual = this block's input at frame t (the causal slice keeps the newest frame)
                qi = self._slot(i)
                ops.conv_block(dt, y1, (1, S, C, C, S * C), pk.w_layers[2 * (i - 1) + 1], 1, 0, C, S, dst, (C, S * C),
                               block_n=self.block_n, scale=bn1[0], shift=bn1[1], relu=True, res=self.rings[i],
                               res_view=(C, 0, 1, qi * S))
            if not last:
                self.rings[i + 1][qn - self.L[i + 1]].copy_(dst)
            h_last = dst
        y = torch.empty((S, pk.n_out), dtype=torch.float32, device=self.dev)
        ops.conv_block(dt, h_last, (1, S, C, C, S * C), pk.w_shrink, 1, 0, C, S, y, (pk.n_out, S * pk.n_out), block_n=64,
                       scale=pk.shrink_scale, shift=pk.shrink_shift, relu=False, out_f32=True, n_valid=pk.n_out)
        self.t += 1
        return y.view(S, m.num_joints_out, 3)

    def step_world(self, X_world, q, t, camera_params):
        """X_world (S, J, 3) world-space joints of this frame, q (S, 4) / t (S, 3) this frame's camera pose per stream,
        camera_params (S, 9) intrinsics incl. distortion -> (S, J_out, 3). world -> camera -> image on the device."""
        J = X_world.shape[1]
        mode = native.PT_WORLD_TO_CAMERA | native.PT_PROJECT
        _, x2d = ops.project_points(X_world, q=q, t=t, cam=camera_params, pts_per_q=J, pts_per_cam=J, mode=mode, want2=True)
        return self.step(x2d)
