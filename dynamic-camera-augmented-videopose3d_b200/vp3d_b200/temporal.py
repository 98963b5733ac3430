"""Host-side driver of the temporal-convolution stack: turns a TemporalModel / TemporalModelOptimized1f module
(parameters under the reference's names) into a sequence of vp3d_conv_block_fwd launches.

Layout in HBM: activations are channels-last [sequence][frame][channel] in the operand type (fp16 / bf16 / fp32-as-tf32),
channels padded to the 256-wide output tile; weights are re-packed once per parameter version into K-major
[c_out_pad][tap][c_in_pad] operands; eval-mode BatchNorm is folded into per-channel scale/shift (fp32).
Nothing here does arithmetic on activations: every FLOP runs in the CUDA library.
"""
import os

import torch

from . import native, ops

K_ALIGN = 64     # input channels are padded to one 128-byte K block of 16-bit elements (two blocks of tf32)
N_TILE = 256     # output-channel tile of the wide layers
N_TILE_NARROW = 64
# eval-mode residual of the dilated model through the TMA side input of the pair kernel (A/B switch; measured in
# profiles/README.md)
RES_VIA_TMA = os.environ.get('VP3D_RES_TMA', '0') == '1'
# fp16 activations saturate at 65,504: a checkpoint whose residual stream exceeds that (huge BatchNorm scales) would yield
# inf / NaN without a word. The guard looks at the OUTPUT (an overflow anywhere in the stack propagates to it as inf / NaN)
# of an eval forward with fp16 operands and finite input -- one device-to-host read, so by default only on the first
# forward after the operands of a model were (re)packed, never inside a stream capture:
#   'first' (default) raise FloatingPointError    'always' check every forward    'bf16' switch the model to bf16 operands
#   (8 more exponent bits, 3 fewer mantissa bits) with a warning and run again    'off'
FP16_GUARD = os.environ.get('VP3D_FP16_GUARD', 'first')
NARROW_TILES = os.environ.get('VP3D_NARROW', '1') != '0'    # A/B switch of the 64-column tiles for launches of few tiles


def _round_up(v, m):
    return (v + m - 1) // m * m


def resolve_dtype(name=None):
    name = name or os.environ.get('VP3D_DTYPE', 'fp16')
    if name not in native.DTYPE_NAMES:
        raise ValueError('unknown operand dtype %r (choose fp16, bf16 or tf32)' % (name,))
    return native.DTYPE_NAMES[name]


class LayerPlan:
    """Geometry of one convolution of the stack (taps, dilation or stride, residual placement)."""
    __slots__ = ('taps', 'dilation', 'stride', 'res_mul', 'res_off')

    def __init__(self, taps, dilation=1, stride=1, res_mul=0, res_off=0):
        self.taps, self.dilation, self.stride, self.res_mul, self.res_off = taps, dilation, stride, res_mul, res_off


class PackedStack:
    """Operand-typed copies of a module's parameters, valid for one (dtype, parameter version) pair."""

    def __init__(self, model, dt):
        self.dt = dt
        ch = model.expand_conv.out_channels
        self.c_in = model.expand_conv.in_channels
        self.c_in_pad = _round_up(self.c_in, K_ALIGN)
        self.c_pad = _round_up(ch, N_TILE)
        self.n_out = model.shrink.out_channels
        self.n_out_pad = _round_up(self.n_out, N_TILE_NARROW)
        dev = model.expand_conv.weight.device
        # Eval-mode BatchNorm: y = conv(x) * scale + shift with scale = gamma / sqrt(var + eps). The scale is folded into
        # the packed weight rows (before they are rounded to the operand type), the epilogue only adds the shift:
        # entries are (None, shift) -- half the per-chunk constant loads and an add instead of an fma per element.
        def fold(conv, bn, k_pad):
            scale, shift = ops.bn_fold(bn, self.c_pad)
            return ops.pack_conv_weight_scaled(dt, conv.weight, scale, self.c_pad, k_pad), (None, shift)
        self.w_expand, self.bn_expand = fold(model.expand_conv, model.expand_bn, self.c_in_pad)
        folded = [fold(conv, bn, self.c_pad) for conv, bn in zip(model.layers_conv, model.layers_bn)]
        self.w_layers = [f[0] for f in folded]
        self.bn_layers = [f[1] for f in folded]
        self.w_shrink = ops.pack_conv_weight(dt, model.shrink.weight, self.n_out_pad, self.c_pad)
        self.shrink_scale = None
        self.shrink_shift = torch.zeros(self.n_out_pad, dtype=torch.float32, device=dev)
        self.shrink_shift[:self.n_out] = model.shrink.bias.detach().float()
        self.range_checked = False     # FP16_GUARD: the output of these operands has been looked at once


def _param_versions(model):
    return tuple((p.data_ptr(), p._version) for p in list(model.parameters()) + list(model.buffers()))


def packed_for(model, dt):
    key = (dt, _param_versions(model))
    cache = model.__dict__.setdefault('_vp3d_pack_cache', {})
    hit = cache.get('eval')
    if hit is None or hit[0] != key:
        hit = (key, PackedStack(model, dt))
        cache['eval'] = hit
    return hit[1]


def _run_layer(dt, x, n, t_in, c_in_pad, w, plan, scale, shift, relu, res=None, res_t=0, out_f32=False, n_valid=None,
               block_n=N_TILE, stats=None, drop=None, fin=None):
    """x: [n][t_in][c_in_pad] operand-typed, contiguous. Returns (y [n][t_out][cols], t_out).

    View selection. A stride==width convolution reads `taps` consecutive frames per output frame, i.e. it is a plain
    GEMM on the reshaped view [t_out][taps*c]; when t_in == taps*t_out the sequences concatenate seamlessly and the
    whole batch is one flat row range (no per-sequence tile padding -- essential for the 1f model whose last layers
    have 3 or 1 frames per sequence). Dilated convolutions tile per sequence; TMA zero-fills the ragged tails."""
    taps, d, s = plan.taps, plan.dilation, plan.stride
    t_out = (t_in - d * (taps - 1) - 1) // s + 1
    assert t_out >= 1, 'sequence shorter than the receptive field'
    n_pad = w.shape[0]
    final = out_f32                      # only the shrink layer asks for fp32 explicitly
    out_f32 = out_f32 or dt == native.TF32
    out_dtype = torch.float32 if out_f32 else ops.torch_dtype(dt)
    cols = n_valid if n_valid is not None else n_pad
    y = torch.empty((n, t_out, cols), dtype=out_dtype, device=x.device)
    res_c = 0 if res is None else res.shape[-1]

    if s > 1:
        assert s == taps and d == 1, 'strided layers of the 1f model have stride == width'
        g_taps, g_step, k_per_tap = 1, 0, taps * c_in_pad
        flat = (t_in == taps * t_out) and res is None
        seq_rows = t_out
    else:
        g_taps, g_step, k_per_tap = taps, d, c_in_pad
        flat = taps == 1 and (res is None or res_t == plan.res_mul * t_out)
        seq_rows = t_in
    if flat:
        a_view = (1, n * seq_rows, k_per_tap, k_per_tap, n * t_in * c_in_pad)
        rows_out = n * t_out
        out_view = (cols, n * t_out * cols)
        res_view = None if res is None else (res_c, n * res_t * res_c, plan.res_mul, plan.res_off)
    else:
        a_view = (n, seq_rows, k_per_tap, k_per_tap, t_in * c_in_pad)
        rows_out = t_out
        out_view = (cols, t_out * cols)
        res_view = None if res is None else (res_c, res_t * res_c, plan.res_mul, plan.res_off)
    if block_n == N_TILE and dt != native.TF32 and drop is None:   # (fused dropout lives in the 256-wide pair kernel)
        # few output tiles (the 1f model's last blocks: 3 or 1 frames per sample): 64-wide column tiles give 4x more
        # CTAs than SMs would otherwise be left idle by 128 x 256 tiles
        tiles = a_view[0] * ((rows_out + 127) // 128) * (n_pad // N_TILE)
        if NARROW_TILES and tiles * 4 <= native.sm_count(x.device):          # the narrow tiles still fit in one wave
            block_n = N_TILE_NARROW
    side_kw = {}
    if (RES_VIA_TMA and res is not None and plan.res_mul == 1 and block_n == N_TILE and not out_f32 and
            dt != native.TF32 and stats is None):
        # residual rows that map 1:1 onto output rows (dilated model): fetched by TMA into the epilogue's staging tile
        # (vp3d_conv_args.side_mode 1) instead of through registers
        side_kw = dict(side=res, side_view=(res_view[0], res_view[1], n * res_t if flat else res_t, plan.res_off),
                       side_mode=1)
        res, res_view = None, None
    ops.conv_block(dt, x, a_view, w, g_taps, g_step, k_per_tap, rows_out, y, out_view, block_n=block_n,
                   scale=scale, shift=shift, relu=relu, res=res, res_view=res_view, out_f32=out_f32, n_valid=cols,
                   **side_kw,
                   out_round_tf32=(dt == native.TF32 and not final),
                   stat_sum=None if stats is None else stats[0], stat_sqsum=None if stats is None else stats[1],
                   drop=drop, fin=fin)
    return y, t_out


def forward_eval(model, x, dt=None):
    """Eval-mode forward of TemporalModel._forward_blocks (TemporalModel.py:126-138) or
    TemporalModelOptimized1f._forward_blocks (:188-198): (N, T, J*F) fp32 -> (N, T', 3*J_out) fp32."""
    ops.require_cuda(x)
    dt = resolve_dtype(getattr(model, 'operand_dtype', None)) if dt is None else dt
    pk = packed_for(model, dt)
    n, t_in, c = x.shape
    assert c == pk.c_in
    strided = model._strided
    fw = model.filter_widths

    h = ops.pack_rows(dt, x.reshape(n * t_in, c), pk.c_in_pad).view(n, t_in, pk.c_in_pad)
    plan = LayerPlan(fw[0], 1, fw[0] if strided else 1)
    h, t = _run_layer(dt, h, n, t_in, pk.c_in_pad, pk.w_expand, plan, pk.bn_expand[0], pk.bn_expand[1], True)

    for i in range(len(fw) - 1):
        w3, w1 = pk.w_layers[2 * i], pk.w_layers[2 * i + 1]
        taps = model.layers_conv[2 * i].kernel_size[0]
        shift = model.causal_shift[i + 1]
        if strided:
            p3 = LayerPlan(taps, 1, taps)
            p1 = LayerPlan(1, 1, 1, res_mul=taps, res_off=shift + taps // 2)  # x[:, :, shift + fw//2 :: fw]
        else:
            d = model.layers_conv[2 * i].dilation[0]
            p3 = LayerPlan(taps, d, 1)
            p1 = LayerPlan(1, 1, 1, res_mul=1, res_off=model.pad[i + 1] + shift)  # x[:, :, pad+shift : T-pad+shift]
        res, res_t = h, t
        h, t = _run_layer(dt, h, n, t, pk.c_pad, w3, p3, pk.bn_layers[2 * i][0], pk.bn_layers[2 * i][1], True)
        h, t = _run_layer(dt, h, n, t, pk.c_pad, w1, p1, pk.bn_layers[2 * i + 1][0], pk.bn_layers[2 * i + 1][1], True,
                          res=res, res_t=res_t)

    y, t = _run_layer(dt, h, n, t, pk.c_pad, pk.w_shrink, LayerPlan(1), pk.shrink_scale, pk.shrink_shift, False,
                      out_f32=True, n_valid=pk.n_out, block_n=N_TILE_NARROW)
    if dt == native.F16 and FP16_GUARD != 'off' and (FP16_GUARD == 'always' or not pk.range_checked):
        if not torch.cuda.is_current_stream_capturing():
            pk.range_checked = True
            if not bool(torch.isfinite(y).all()) and bool(torch.isfinite(x).all()):
                if FP16_GUARD == 'bf16':
                    import warnings
                    warnings.warn('vp3d_b200: fp16 activations overflowed (|x| > 65504) in this model; switching it to '
                                  'bf16 operands (model.operand_dtype = "bf16")')
                    model.operand_dtype = 'bf16'
                    return forward_eval(model, x, native.BF16)
                raise FloatingPointError(
                    'vp3d_b200: the forward produced inf / NaN from finite input with fp16 operands -- an activation '
                    'exceeded the fp16 range (65504). Set model.operand_dtype = "bf16" (or "tf32"), or VP3D_FP16_GUARD=bf16 '
                    'to switch automatically; VP3D_FP16_GUARD=off disables this check.')
    return y
