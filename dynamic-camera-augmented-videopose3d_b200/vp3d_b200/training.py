"""Train-mode forward and backward of the temporal-convolution stack (TemporalModel.py:126-138 / :188-198 in train()
mode and the autograd backward run.py:485 triggers), as one torch.autograd.Function over the whole stack.

Every FLOP runs in libvp3d_b200.so:
  forward   conv GEMM (raw output z + per-channel sum / sum of squares in the epilogue)  vp3d_conv_block_fwd
            batch statistics -> scale/shift, running statistics                          vp3d_bn_finalize
            a = dropout(relu(z * scale + shift)) [+ residual rows]                       vp3d_bn_act_fwd
            (both in one launch: vp3d_bn_finalize_act_fwd, optional)
  backward  dz = BN/ReLU/dropout backward of the incoming gradient (two passes)           vp3d_bn_act_bwd_*
            dW = dz^T a_in (stream-K tcgen05 GEMM, MN-major operands)                    vp3d_wgrad (+ _finish)
            g_in = dz W (same kernel as the forward, transposed weights, residual fan-in) vp3d_conv_block_fwd
PyTorch only allocates the buffers and records the graph edge. Gradients travel in the 16-bit operand type multiplied
by a power-of-two scale chosen on the device from max|dL/dy| (no host synchronisation); parameter gradients come out
as unscaled fp32 in the nn.Conv1d / nn.BatchNorm1d layouts, so torch.optim and load/state_dict work unchanged.
"""
import os

import torch

from . import native, ops
from .temporal import K_ALIGN, N_TILE, NARROW_TILES, LayerPlan, _round_up, _run_layer, resolve_dtype

SHRINK_PAD = 128      # shrink-layer output channels are padded to one 128-row MMA tile for its weight gradient
grad_ready_hook = None   # set by vp3d_b200.ddp: called as hook(parameter, gradient) as soon as a gradient is issued
grad_finish_hook = None  # ... and once at the end of the backward (waits for the outstanding all-reduces); may return
                         # {id(parameter): gradient} replacements for the tensors the backward hands to autograd
grad_begin_hook = None   # ... and once at its start
grad_alloc = None        # set by vp3d_b200.ddp (peer exchange): grad_alloc(parameter) -> flat fp32 buffer the gradient of
                         # that parameter is to be written into (a slot of the exchange buffer), or None
# Optimiser update inside the backward (vp3d_b200.optim.FusedAdam.update_in_backward): param_update_hook([(parameter,
# gradient), ...]) applies the update of those parameters NOW, on the current stream. The backward calls it on a third
# stream the moment a gradient is final and nothing of this backward reads the parameter (or its packed operand) again, so
# the HBM-bound update runs beside the tensor-bound weight-gradient GEMMs of the earlier layers instead of after the
# backward. `update_stream` / `update_filter` are set by the data-parallel exchange (the update of an exchanged gradient
# must follow its all-reduce: same stream; gradients exchanged at the end of the backward are left to optimizer.step()).
param_update_hook = None
update_stream = None     # callable -> torch.cuda.Stream, or None (a private stream per device)
update_filter = None     # callable(parameter) -> bool, or None (every parameter)
_update_streams = {}
sync_bn_group = None     # process group over which train-mode BatchNorm statistics are summed (None: per replica)
# vp3d_bn_finalize_act_fwd (statistics -> scale/shift inside the apply pass, one launch less per layer) is available but
# off: same-box A/B at batch 1024 gave 2.01-2.03 ms per step with it against 1.97 without -- every block of the apply pass
# then starts with dependent loads of the double-precision sums before its first row, which costs more than the 5 us
# finalize launch it saves. VP3D_FUSE_BN=1 switches it on.
fuse_bn_finalize = os.environ.get('VP3D_FUSE_BN', '0') == '1'
overlap_wgrad = True     # weight-gradient GEMMs on a second stream, concurrent with the HBM-bound BN backward passes
_side_streams = {}
_ones = {}               # constant unit scale vectors of the shrink layer, per (device, padded width)
# Expand layer without its HBM-bound passes (csrc/expand.cu): BatchNorm statistics from the Gram matrix of the layer
# input, BatchNorm + ReLU + dropout applied by the GEMM epilogue, and a backward that needs neither the raw output nor a
# reduction / apply pass. VP3D_FUSED_EXPAND=0 restores bn_finalize + bn_act_fwd / bn_act_bwd for that layer (the path
# every other layer takes; also used under SyncBN, where the statistics must be exchanged between the two steps).
fused_expand = os.environ.get('VP3D_FUSED_EXPAND', '1') != '0'
# vp3d_bn_finalize in the tail of the producing GEMM (last CTA done; vp3d_conv_args.fin): 8 launches and their dependent-
# launch gaps less per training forward. VP3D_FIN_IN_GEMM=0 restores the stand-alone finalize launch.
finalize_in_gemm = os.environ.get('VP3D_FIN_IN_GEMM', '1') != '0'
# Issue order of a layer's two backward GEMMs. Both become ready when the layer's dz exists and cannot share an SM (each
# is a persistent kernel with ~200 KB of shared memory), so one runs after the other whatever the streams say. 'after':
# the weight-gradient GEMM is made to wait for the data-gradient GEMM -- it then runs on the side stream BESIDE the HBM-bound
# BatchNorm backward of the layer below (which needs the data gradient and is what the critical path continues with);
# 'before': both are released together and the hardware picks (VP3D_WGRAD_ORDER).
wgrad_after_dgrad = os.environ.get('VP3D_WGRAD_ORDER', 'after') != 'before'
# The forward's BatchNorm apply pass stores the keep decision (ReLU passed and not dropped) as one bit per element; the two
# backward passes read it instead of recomputing the Philox stream and the affine comparison and run at memory speed
# (vp3d_bn_act_fwd_mask / vp3d_bn_act_bwd_*_mask). Measured (ncu, serialised): backward passes 327 -> 271 us per step, the
# forward pass 119 -> 133 us (building and storing the bits); the backward passes run beside the weight-gradient GEMMs
# while the forward pass is on the critical path, so the STEP does not gain: 1.684-1.725 ms with the mask against
# 1.670-1.693 without (three alternating runs, same box). Off by default; VP3D_BN_MASK=1 turns it on.
store_keep_mask = os.environ.get('VP3D_BN_MASK', '0') == '1'
prefill_arena = os.environ.get('VP3D_PREFILL_ARENA', '1') != '0'   # zero arena of the backward filled during the forward
debug_keep_saved = False  # tests: keep the last forward's saved per-layer tensors in `debug_last_saved`
debug_last_saved = None


class _Layer:
    """Everything the backward needs about one convolution + BatchNorm + activation of the stack."""
    __slots__ = ('conv', 'bn', 'taps', 'dilation', 'stride', 't_in', 't_out', 'c_in', 'c_in_pad', 'a_in', 'z', 'scale',
                 'shift', 'mean', 'invstd', 'drop', 'res_of', 'res_mul', 'res_off', 'res_t', 'w_fwd', 'count', 'fused', 'mask')


def _conv_w(dt, conv, rows_pad, k_pad):
    """Forward-packed operand [rows_pad][taps * k_pad] of a convolution weight. The buffer is registered on the
    parameter (`_vp3d_packed`, keyed by operand type and padding, stamped with the parameter version): a weight that
    has not changed is not packed again, and vp3d_b200.optim.FusedAdam refreshes the registered buffer inside its
    update kernel, so with that optimiser the pack kernel disappears from the training step."""
    w = conv.weight
    reg = w.__dict__.setdefault('_vp3d_packed', {})
    key = (dt, rows_pad, k_pad)
    entry = reg.get(key)
    if entry is not None and entry[1] == w._version and entry[0].device == w.device:
        return entry[0]
    packed = ops.pack_conv_weight(dt, w, rows_pad, k_pad)
    reg[key] = [packed, w._version]
    return packed


def _step_counter(model, dev):
    """Per-model training-step counter on the device (int64), ticked once per training forward by a kernel so that a
    captured CUDA graph of the step draws new dropout masks on every replay."""
    c = model.__dict__.get('_vp3d_step_counter')
    if c is None or c.device != dev:
        c = torch.zeros(1, dtype=torch.int64, device=dev)
        model.__dict__['_vp3d_step_counter'] = c
    return c


def _dropout_for(model, layer_idx, counter):
    p = float(model.drop.p) if model.training else 0.0
    return ops.make_dropout(p, torch.initial_seed(), layer_idx, counter)


def _forward_stack(model, x, dt, keep_masks=False):
    """-> (y fp32 (N, T', 3*J_out), saved layers). x: (N, T, C_in) fp32 CUDA."""
    n, t_in, c_in = x.shape
    strided = model._strided
    fw = model.filter_widths
    ch = model.expand_conv.out_channels
    c_pad = _round_up(ch, N_TILE)
    c_in_pad = _round_up(c_in, K_ALIGN)
    dev = x.device
    counter = _step_counter(model, dev)
    ops.counter_add(counter, 1)
    step = counter.clone()   # this call's own copy: its backward sees the same value even if another forward runs first
    layers = []
    # BatchNorm statistics of all layers in one zero-filled arena (one fill launch per forward instead of one per layer);
    # each layer's row ends with the word in which its GEMM counts finished CTAs (in-GEMM finalize)
    stats_all = torch.zeros((2 * len(fw) - 1, 2 * c_pad + 2), dtype=torch.float64, device=dev)

    def conv_bn_act(idx, conv, bn, a_in, t, cin, cin_pad, plan, res=None, res_t=0, res_mul=1, res_off=0):
        L = _Layer()
        L.fused = None
        L.mask = None
        L.conv, L.bn = conv, bn
        L.taps, L.dilation, L.stride = plan.taps, plan.dilation, plan.stride
        L.c_in, L.c_in_pad, L.t_in, L.a_in = cin, cin_pad, t, a_in
        w = _conv_w(dt, conv, c_pad, cin_pad)
        L.w_fwd = w   # [c_out_pad][taps * c_in_pad]; the data-gradient GEMM reads it again as W^T (MN-major operand)
        stats = stats_all[idx, :2 * c_pad].view(2, c_pad)
        # per-channel sum / sum of squares of the stored z come out of the GEMM epilogue (the staged output tile is read
        # back column-wise from shared memory); vp3d_col_stats remains as a stand-alone entry point
        taps_, d_, s_ = plan.taps, plan.dilation, plan.stride
        t_out_pre = (t - d_ * (taps_ - 1) - 1) // s_ + 1
        fin = fin_out = None
        if finalize_in_gemm and sync_bn_group is None and not fuse_bn_finalize:
            # the last CTA of the GEMM finalizes the statistics (no vp3d_bn_finalize launch)
            fin, fin_out = ops.make_bn_fin(bn, n * t_out_pre, c_pad, stats_all[idx, 2 * c_pad:].view(torch.int32))
        z, t_out = _run_layer(dt, a_in, n, t, cin_pad, w, plan, None, None, False, stats=stats, fin=fin)
        L.z, L.t_out = z, t_out
        count = n * t_out
        if sync_bn_group is not None:
            import torch.distributed as dist
            dist.all_reduce(stats, group=sync_bn_group)
            count *= dist.get_world_size(sync_bn_group)
        L.count = count
        L.drop = _dropout_for(model, idx, step)
        L.res_of, L.res_t, L.res_mul, L.res_off = res, res_t, res_mul, res_off
        want_mask = store_keep_mask and keep_masks     # only a forward that will be differentiated stores them
        if fin is not None:
            L.scale, L.shift, L.mean, L.invstd = fin_out
            a = ops.bn_act_fwd(dt, z, L.scale, L.shift, n, t_out, L.drop, res=res, res_seq_rows=res_t,
                               res_row_mul=res_mul, res_row_off=res_off, want_mask=want_mask)
        elif fuse_bn_finalize and bn.momentum is not None:
            a, L.scale, L.shift, L.mean, L.invstd = ops.bn_finalize_act_fwd(
                dt, z, stats, count, bn, n, t_out, L.drop, res=res, res_seq_rows=res_t, res_row_mul=res_mul,
                res_row_off=res_off)
        else:
            L.scale, L.shift, L.mean, L.invstd = ops.bn_finalize(stats, count, bn, c_pad)
            a = ops.bn_act_fwd(dt, z, L.scale, L.shift, n, t_out, L.drop, res=res, res_seq_rows=res_t,
                               res_row_mul=res_mul, res_row_off=res_off, want_mask=want_mask)
        if isinstance(a, tuple):
            a, L.mask = a
        layers.append(L)
        return a, t_out

    plan = LayerPlan(fw[0], 1, fw[0] if strided else 1)
    k0 = fw[0] * c_in_pad
    if (fused_expand and sync_bn_group is None and len(fw) > 1 and c_in < c_in_pad and k0 <= 256 and
            model.expand_conv.dilation[0] == 1):
        h, t = _expand_fused(model, dt, x, n, t_in, c_in, c_in_pad, c_pad, plan, _dropout_for(model, 0, step), layers)
    else:
        h = ops.pack_rows(dt, x.reshape(n * t_in, c_in), c_in_pad).view(n, t_in, c_in_pad)
        h, t = conv_bn_act(0, model.expand_conv, model.expand_bn, h, t_in, c_in, c_in_pad, plan)
    for i in range(len(fw) - 1):
        conv3, conv1 = model.layers_conv[2 * i], model.layers_conv[2 * i + 1]
        taps = conv3.kernel_size[0]
        shift = model.causal_shift[i + 1]
        if strided:
            p3 = LayerPlan(taps, 1, taps)
            res_mul, res_off = taps, shift + taps // 2            # x[:, :, shift + fw//2 :: fw]   (:192)
        else:
            p3 = LayerPlan(taps, conv3.dilation[0], 1)
            res_mul, res_off = 1, model.pad[i + 1] + shift        # x[:, :, pad+shift : T-pad+shift] (:132)
        res, res_t = h, t
        h, t = conv_bn_act(2 * i + 1, conv3, model.layers_bn[2 * i], h, t, ch, c_pad, p3)
        h, t = conv_bn_act(2 * i + 2, conv1, model.layers_bn[2 * i + 1], h, t, ch, c_pad, LayerPlan(1), res=res,
                           res_t=res_t, res_mul=res_mul, res_off=res_off)

    n_out = model.shrink.out_channels
    n_out_pad = _round_up(n_out, 64)
    w_shrink = _conv_w(dt, model.shrink, n_out_pad, c_pad)
    bias = torch.nn.functional.pad(model.shrink.bias.detach().float(), (0, n_out_pad - n_out))   # one launch
    ones = _ones.get((dev, n_out_pad))
    if ones is None:
        ones = _ones[(dev, n_out_pad)] = torch.ones(n_out_pad, dtype=torch.float32, device=dev)
    y, t = _run_layer(dt, h, n, t, c_pad, w_shrink, LayerPlan(1), ones, bias, False, out_f32=True, n_valid=n_out,
                      block_n=64)
    return y, layers, h, t, c_pad, w_shrink


def _expand_fused(model, dt, x, n, t_in, c_in, c_in_pad, c_pad, plan, drop, layers):
    """expand_conv -> expand_bn -> relu -> drop (TemporalModel.py:127 / :189) as Gram GEMM + statistics kernel + ONE
    convolution GEMM whose epilogue applies BatchNorm, ReLU and dropout (csrc/expand.cu). Appends the saved layer."""
    conv, bn = model.expand_conv, model.expand_bn
    taps, s = plan.taps, plan.stride
    k_total = taps * c_in_pad
    ones_col = c_in                                      # first padding column of tap 0
    h = ops.pack_rows(dt, x.reshape(n * t_in, c_in), c_in_pad, ones_col=ones_col).view(n, t_in, c_in_pad)
    t_out = (t_in - taps) // s + 1
    # X = the [rows][k_total] view of the packed input the convolution contracts over: consecutive frames of a strided
    # (stride == width) layer, overlapping windows of a stride-1 layer (row stride = one frame)
    row_stride = k_total if s > 1 else c_in_pad
    if s > 1 and t_in == taps * t_out:
        xv, av = (1, n * t_out, row_stride, n * t_in * c_in_pad), (n * t_out, k_total, row_stride, n * t_in * c_in_pad)
    else:
        xv, av = (n, t_out, row_stride, t_in * c_in_pad), (t_out, k_total, row_stride, t_in * c_in_pad)
    # one 256 x 256 tile, 64 row slices: 27 us at batch 1024 (tools/gram_probe.py: 128 x 64 tiles 32 us, 148 slices 33 us,
    # 16 slices 58 us)
    gram = torch.zeros((1, 256, 256), dtype=torch.float32, device=x.device)
    ops.wgrad(dt, h, xv, h, av, 256, 256, 1, gram, block_n=256, dz_cols=k_total)
    w = _conv_w(dt, conv, c_pad, c_in_pad)
    L = _Layer()
    L.mask = None
    L.conv, L.bn, L.w_fwd = conv, bn, w
    L.taps, L.dilation, L.stride = plan.taps, plan.dilation, plan.stride
    L.c_in, L.c_in_pad, L.t_in, L.a_in = c_in, c_in_pad, t_in, h
    L.scale, L.shift, L.mean, L.invstd, wg = ops.expand_bn_stats(dt, gram, w, k_total, ones_col, bn, c_pad)
    L.drop = drop
    a, t_out2 = _run_layer(dt, h, n, t_in, c_in_pad, w, plan, L.scale, L.shift, True, drop=drop)
    assert t_out2 == t_out
    L.z, L.t_out, L.count = None, t_out, n * t_out
    L.res_of, L.res_t, L.res_mul, L.res_off = None, 0, 1, 0
    L.fused = dict(gram=gram, wg=wg, xv=xv, av=av, k_total=k_total, ones_col=ones_col, a=a,
                   keep_scale=ops.keep_scale(drop.p))
    layers.append(L)
    return a, t_out


def _views(L, n, c_pad):
    """(dz view, input view, flat?) of one layer for the weight-gradient GEMM (vp3d_wgrad)."""
    taps, d, s = L.taps, L.dilation, L.stride
    if s > 1:
        flat = L.t_in == taps * L.t_out
        k = taps * L.c_in_pad
        if flat:
            return ((1, n * L.t_out, c_pad, n * L.t_out * c_pad), (n * L.t_out, k, k, n * L.t_in * L.c_in_pad), 0,
                    L.c_in_pad)
        return ((n, L.t_out, c_pad, L.t_out * c_pad), (L.t_out, k, k, L.t_in * L.c_in_pad), 0, L.c_in_pad)
    if taps == 1:
        return ((1, n * L.t_out, c_pad, n * L.t_out * c_pad), (n * L.t_in, L.c_in_pad, L.c_in_pad,
                                                              n * L.t_in * L.c_in_pad), 0, 0)
    return ((n, L.t_out, c_pad, L.t_out * c_pad), (L.t_in, L.c_in_pad, L.c_in_pad, L.t_in * L.c_in_pad), d, 0)


def _packed_floats(L, c_pad):
    """fp32 elements of the zero-filled split-K accumulator of one layer's weight gradient."""
    if L.fused is not None or (L.stride > 1 and L.taps * L.c_in_pad <= 256):
        return c_pad * 256
    return L.taps * c_pad * L.c_in_pad


class _ZeroArena:
    """ONE zero-filled buffer per backward for every split-K accumulator and the BatchNorm-backward sums (one fill
    launch instead of one per layer: ~20 launches of 2-4 us each at batch 1024)."""

    def __init__(self, n_floats, device):
        self.buf = torch.zeros(n_floats, dtype=torch.float32, device=device)
        self.used = 0

    def take(self, shape, dtype=torch.float32):
        n = 1
        for d in shape:
            n *= d
        words = n * (2 if dtype == torch.float64 else 1)
        self.used = (self.used + 3) // 4 * 4                    # 16-byte aligned views
        out = self.buf[self.used:self.used + words]
        assert out.numel() == words, 'zero arena exhausted'
        self.used += words
        return out.view(dtype).view(shape)


def _slot(param):
    return grad_alloc(param) if grad_alloc is not None else None


def _weight_grad(dt, L, dz, n, c_pad, gscale, keep=None, arena=None):
    dzv, av, row_step, col_step = _views(L, n, c_pad)
    c_out = L.conv.out_channels
    out = _slot(L.conv.weight)
    zeros = (lambda shape: arena.take(shape)) if arena is not None else (
        lambda shape: torch.zeros(shape, dtype=torch.float32, device=dz.device))
    if L.stride > 1 and L.taps * L.c_in_pad <= 256:
        # narrow strided layer (expand: 3 x 64 input columns): all taps are adjacent columns of the reshaped view, so
        # they form ONE 256-wide tile (columns past taps * c_in_pad are zero-filled by TMA) and dz is read once
        packed = zeros((1, c_pad, 256))
        if keep is not None:
            keep.append(packed)
        ops.wgrad(dt, dz, dzv, L.a_in, av, c_pad, 256, 1, packed, block_n=256)
        return ops.wgrad_finish(packed, c_out, L.c_in, L.taps, c_pad, 256, gscale, tap_stride=L.c_in_pad, row_stride=256,
                                out=out)
    block_n = 256 if L.c_in_pad % 256 == 0 else 64
    packed = zeros((L.taps, c_pad, L.c_in_pad))
    if keep is not None:
        keep.append(packed)
    ops.wgrad(dt, dz, dzv, L.a_in, av, c_pad, L.c_in_pad, L.taps, packed, b_tap_row_step=row_step,
              b_tap_col_step=col_step, block_n=block_n)
    return ops.wgrad_finish(packed, c_out, L.c_in, L.taps, c_pad, L.c_in_pad, gscale, out=out)


def _data_grad(dt, L, dz, n, c_pad, fan_in=None, fan_rows=0, fan_off=0, fan_mul=1, gate=None):
    """Gradient wrt the layer input: g_in[s][t'][ci] = sum_{tap, co} dz[s][t' - tap * d][co] * W[co][tap][ci]. The
    forward-packed weights [co][tap * c_in_pad + ci] are the W^T operand as they are (MN-major), no transposed copy.
    `fan_in` is the block-output gradient that also reaches this input through the residual slice.
    `gate` = (a, keep_scale): the input of this layer is the activation `a` of a layer whose backward works on the
    GATED gradient (the fused expand layer): the epilogue multiplies by keep_scale where a > 0 and zeroes the rest."""
    taps, d, s = L.taps, L.dilation, L.stride
    cin_pad = L.c_in_pad
    g_in = torch.empty((n, L.t_in, cin_pad), dtype=dz.dtype, device=dz.device)
    block_n = 256 if cin_pad % 256 == 0 else 64
    if gate is not None:
        assert block_n == 256
    elif NARROW_TILES and block_n == 256 and ((n * L.t_out + 127) // 128) * (taps * cin_pad // 256) * 4 <= native.sm_count(
            dz.device):
        block_n = 64    # few tiles: narrower column tiles keep all SMs busy (while they still fit in one wave)
    if s > 1 or taps == 1:
        if s > 1 and L.t_in != taps * L.t_out:
            raise RuntimeError('vp3d_b200: training a strided (1f) block needs t_in == %d * t_out (got %d -> %d)'
                               % (taps, L.t_in, L.t_out))
        rows = n * L.t_out
        n_cols = taps * cin_pad          # stride == width: the [rows][taps * C] view of g_in is one plain GEMM
        kw = {}
        if fan_in is not None:
            # residual x[:, :, off::taps]: in that view it is the column block [off*C, (off+1)*C)
            kw = dict(res=fan_in, res_view=(c_pad, rows * c_pad, 1, 0), res_col_off=fan_off * cin_pad, res_cols=cin_pad)
        if gate is not None:
            kw.update(side=gate[0], side_view=(n_cols, rows * n_cols, rows, 0), side_mode=2, side_scale=gate[1])
        ops.conv_block(dt, dz, (1, rows, c_pad, c_pad, rows * c_pad), L.w_fwd, 1, 0, c_pad, rows, g_in,
                       (n_cols, rows * n_cols), block_n=block_n, w_mn_major=(n_cols, 0), **kw)
        return g_in
    # dilated: rows of dz outside [0, t_out) read as zero through TMA
    kw = {}
    if fan_in is not None:
        # residual x[:, :, off : off + fan_rows]: input row t' receives block-output row t' - off
        kw = dict(res=fan_in, res_view=(c_pad, fan_rows * c_pad, 1, -fan_off), res_rows=fan_rows)
    if gate is not None:
        kw.update(side=gate[0], side_view=(cin_pad, L.t_in * cin_pad, L.t_in, 0), side_mode=2, side_scale=gate[1])
    ops.conv_block(dt, dz, (n, L.t_out, c_pad, c_pad, L.t_out * c_pad), L.w_fwd, taps, -d, c_pad, L.t_in, g_in,
                   (cin_pad, L.t_in * cin_pad), block_n=block_n, w_mn_major=(cin_pad, cin_pad), **kw)
    return g_in


def _side_stream(device, main):
    key = (device.index, main.cuda_stream)
    side = _side_streams.get(key)
    if side is None:
        side = _side_streams[key] = torch.cuda.Stream(device)
    return side


def _arena_floats(model, c_in, n_layers):
    """Upper bound of the fp32 words the backward takes from its zero arena (_ZeroArena): BatchNorm-backward sums, the
    split-K accumulators of every weight gradient, slack for 16-byte alignment."""
    ch = model.expand_conv.out_channels
    c_pad = _round_up(ch, N_TILE)
    c_in_pad = _round_up(c_in, K_ALIGN)
    taps0 = model.expand_conv.kernel_size[0]
    total = 4 * n_layers * c_pad + SHRINK_PAD * c_pad + 8 * (n_layers + 2) + 256
    total += max(c_pad * 256, taps0 * c_pad * c_in_pad)
    for conv in model.layers_conv:
        total += conv.kernel_size[0] * c_pad * c_pad
    return total


class _StackTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, dt, x, *params):
        # The backward's zero-filled accumulators (68 MB at 1024 channels: one fill launch of ~11 us at HBM speed) depend on
        # nothing, so the fill is issued NOW on the side stream and runs beside the forward's first kernels instead of at
        # the head of the backward's critical path; forked from and joined to the current stream inside this call, so the
        # step stays capturable.
        arena = None
        if overlap_wgrad and prefill_arena:
            main = torch.cuda.current_stream(x.device)
            side = _side_stream(x.device, main)
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                arena = _ZeroArena(_arena_floats(model, x.shape[-1], 2 * len(model.filter_widths) - 1), x.device)
        y, layers, a_last, t_last, c_pad, w_shrink = _forward_stack(model, x, dt, keep_masks=True)
        if arena is not None:
            main.wait_stream(side)
        ctx.arena = arena
        ctx.w_shrink = w_shrink
        if debug_keep_saved:
            global debug_last_saved
            debug_last_saved = layers
        ctx.model, ctx.dt, ctx.layers, ctx.a_last, ctx.t_last, ctx.c_pad = model, dt, layers, a_last, t_last, c_pad
        ctx.n = x.shape[0]
        ctx.params = params
        # The backward reads the packed weight operands again (data gradient). They are plain buffers that
        # FusedAdam.step() rewrites in place, outside autograd's version tracking, so the parameter versions of this
        # forward are stamped here and checked in the backward.
        ctx.weight_versions = [(L.conv.weight, L.conv.weight._version) for L in layers]
        ctx.weight_versions.append((model.shrink.weight, model.shrink.weight._version))
        return y

    @staticmethod
    def backward(ctx, dy):
        model, dt, layers, n, c_pad = ctx.model, ctx.dt, ctx.layers, ctx.n, ctx.c_pad
        if layers is None:
            raise RuntimeError('vp3d_b200: the saved activations of this forward were released by its first backward '
                               '(a second backward / retain_graph=True is not supported on the fused training path)')
        for w, version in ctx.weight_versions:
            if w._version != version:
                raise RuntimeError('vp3d_b200: a convolution weight was modified (optimizer.step()?) between the forward '
                                   'and its backward; the data gradient would be taken with the NEW weights. Run '
                                   'backward() before step(), as run.py:485-487 does.')
        hook = grad_ready_hook
        grads = {}
        if grad_begin_hook is not None:
            grad_begin_hook()
        # The weight gradient of layer L (tensor-core bound) depends only on dz_L and the saved input; the critical path
        # continues with dgrad_L -> BN/ReLU backward of layer L-1 (HBM bound). Issuing the wgrad GEMMs (+ their layout
        # pass, + the gradient all-reduce hook) on a second stream lets the two kinds of work share the SMs: the wgrad
        # kernel leaves enough registers for one BN-backward block per SM. Every tensor the side stream touches is kept
        # alive in `keep` until the streams have joined (no reliance on record_stream, so the step stays capturable).
        main = torch.cuda.current_stream(dy.device)
        side = _side_stream(dy.device, main) if overlap_wgrad else None
        keep = []

        def done(param, g):
            grads[id(param)] = g
            if hook is not None:
                hook(param, g)

        upd = None
        if param_update_hook is not None:
            if update_stream is not None:
                upd = update_stream()
            else:
                upd = _update_streams.get(dy.device.index)
                if upd is None:
                    upd = _update_streams[dy.device.index] = torch.cuda.Stream(dy.device)
        early = []      # parameters whose gradient has been issued and which this backward does not read again
        upd_used = []   # something was issued on the update stream (only then is it joined: it may be outside a capture)

        def release(*params):
            if upd is not None:
                early.extend(q for q in params if update_filter is None or update_filter(q))

        def flush_early():
            """Optimiser update of the released parameters on the update stream, after everything issued so far on the
            main and the side stream (the gradient's producer; the last reader of the packed weights)."""
            if upd is None or not early:
                return
            pairs = [(q, grads[id(q)]) for q in early if id(q) in grads]
            early[:] = [q for q in early if id(q) not in grads]
            if not pairs:
                return
            evs = [torch.cuda.Event()]
            evs[0].record(main)
            if side is not None:
                evs.append(torch.cuda.Event())
                evs[1].record(side)
            with torch.cuda.stream(upd):
                for ev in evs:
                    upd.wait_event(ev)
                param_update_hook(pairs)
            upd_used.append(True)

        def on_side(fn, *tensors):
            """Runs fn() on the side stream after everything issued so far on the main stream."""
            if side is None:
                return fn()
            keep.extend(tensors)
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                return fn()

        # every zero-initialised accumulator of this backward in one fill
        arena = ctx.arena      # filled during the forward (side stream, joined there)
        if arena is None:
            arena = _ZeroArena(4 * len(layers) * c_pad + SHRINK_PAD * c_pad + sum(_packed_floats(L, c_pad) for L in layers) +
                               8 * (len(layers) + 2), dy.device)
        sums_all = arena.take((len(layers), 2, c_pad), torch.float64)

        # ---- shrink layer: y = a_last W^T + b
        n_out = model.shrink.out_channels
        dy2 = ops.f32c(dy).reshape(n * ctx.t_last, n_out)
        gscale = ops.grad_scale(dy2)
        dzs, dbias = ops.grad_pack_rows(dt, dy2, SHRINK_PAD, gscale, want_col_sum=True,
                                        col_sum_out=_slot(model.shrink.bias))
        done(model.shrink.bias, dbias)
        release(model.shrink.bias)
        rows = n * ctx.t_last

        def shrink_wgrad():
            packed = arena.take((1, SHRINK_PAD, c_pad))
            ops.wgrad(dt, dzs, (1, rows, SHRINK_PAD, rows * SHRINK_PAD), ctx.a_last, (rows, c_pad, c_pad, rows * c_pad),
                      SHRINK_PAD, c_pad, 1, packed)
            keep.append(packed)
            done(model.shrink.weight,
                 ops.wgrad_finish(packed, n_out, model.shrink.in_channels, 1, SHRINK_PAD, c_pad, gscale,
                                  out=_slot(model.shrink.weight)))
        if not wgrad_after_dgrad:
            on_side(shrink_wgrad, dzs, gscale)
        g = torch.empty((n, ctx.t_last, c_pad), dtype=dzs.dtype, device=dy.device)
        k_shrink = ctx.w_shrink.shape[0]      # forward-packed [n_out_pad][c_pad], read as W^T
        ops.conv_block(dt, dzs, (1, rows, SHRINK_PAD, SHRINK_PAD, rows * SHRINK_PAD), ctx.w_shrink, 1, 0, k_shrink, rows,
                       g, (c_pad, rows * c_pad), w_mn_major=(c_pad, 0))
        if wgrad_after_dgrad:
            on_side(shrink_wgrad, dzs, gscale)
        release(model.shrink.weight)

        # ---- blocks and the expand layer, last to first. `g` is the gradient wrt the current layer's output.
        for idx in range(len(layers) - 1, -1, -1):
            L = layers[idx]
            rows = n * L.t_out
            if L.fused is not None:
                # fused expand layer: `g` arrived gated by its ReLU / dropout mask (gm); everything else is P = gm^T X
                F = L.fused

                def expand_grads(L=L, F=F, gm=g):
                    p_packed = arena.take((1, c_pad, 256))
                    keep.append(p_packed)
                    seqs, rws = F['xv'][0], F['xv'][1]
                    ops.wgrad(dt, gm, (seqs, rws, c_pad, rws * c_pad), L.a_in, F['av'], c_pad, 256, 1, p_packed,
                              block_n=256)
                    dw, dgamma, dbeta = ops.expand_bwd_finish(dt, p_packed, F['wg'], F['gram'], L.w_fwd, F['k_total'],
                                                              F['ones_col'], L.scale, L.mean, L.invstd, gscale,
                                                              L.bn.num_features, c_pad, L.c_in, L.c_in_pad, L.taps,
                                                              out=(_slot(L.conv.weight), _slot(L.bn.weight),
                                                                   _slot(L.bn.bias)))
                    done(L.bn.weight, dgamma)
                    done(L.bn.bias, dbeta)
                    done(L.conv.weight, dw)
                on_side(expand_grads, g)
                release(L.bn.weight, L.bn.bias, L.conv.weight)
                break
            dz, dgamma, dbeta = ops.bn_act_bwd(dt, g, L.z, L.scale, L.shift, L.mean, L.invstd, rows, L.bn.num_features,
                                               L.drop, gscale, count=L.count, group=sync_bn_group, sums=sums_all[idx],
                                               out=(_slot(L.bn.weight), _slot(L.bn.bias)), mask=L.mask)
            done(L.bn.weight, dgamma)
            done(L.bn.bias, dbeta)
            release(L.bn.weight, L.bn.bias)
            wgrad = lambda L=L, dz=dz: done(L.conv.weight, _weight_grad(dt, L, dz, n, c_pad, gscale, keep, arena))
            if idx == 0 or not wgrad_after_dgrad:
                on_side(wgrad, dz)
            if idx == 0:
                release(L.conv.weight)
                break  # no gradient wrt the 2-D keypoints (the reference never asks for one, run.py:458-485)
            if L.res_of is not None:
                # second convolution of a block: its output gradient g also feeds the residual source; that fan-in is
                # added by the data-gradient epilogue of the block's first convolution (next iteration)
                fan = (g, L.t_out, L.res_off, L.res_mul)
                g = _data_grad(dt, L, dz, n, c_pad)
            else:
                g_block, fan_rows, fan_off, fan_mul = fan
                below = layers[idx - 1].fused
                gate = (below['a'], below['keep_scale']) if below is not None else None
                g = _data_grad(dt, L, dz, n, c_pad, fan_in=g_block, fan_rows=fan_rows, fan_off=fan_off, fan_mul=fan_mul,
                               gate=gate)
            if wgrad_after_dgrad:
                on_side(wgrad, dz)
            # the data gradient was the last reader of this layer's packed weights
            release(L.conv.weight)
            flush_early()
        flush_early()
        if side is not None:
            main.wait_stream(side)
        if upd is not None and upd_used:
            main.wait_stream(upd)
        if grad_finish_hook is not None:
            grads.update(grad_finish_hook() or {})
        keep.clear()
        out = [grads.get(id(p)) for p in ctx.params]
        ctx.layers = ctx.a_last = ctx.arena = None
        return (None, None, None) + tuple(out)


def stack_parameters(model):
    ps = [model.expand_conv.weight, model.expand_bn.weight, model.expand_bn.bias]
    for conv, bn in zip(model.layers_conv, model.layers_bn):
        ps += [conv.weight, bn.weight, bn.bias]
    ps += [model.shrink.weight, model.shrink.bias]
    return ps


def forward_train(model, x):
    """Train-mode _forward_blocks: (N, T, J*F) fp32 CUDA -> (N, T', 3*J_out) fp32 with a grad_fn."""
    ops.require_cuda(x)
    dt = resolve_dtype(getattr(model, 'operand_dtype', None))
    if dt == native.TF32:
        raise RuntimeError('vp3d_b200: the training path runs fp16 or bf16 operands (fp32 accumulation); '
                           'set model.operand_dtype / VP3D_DTYPE to fp16 or bf16')
    if x.requires_grad and torch.is_grad_enabled():
        raise RuntimeError('vp3d_b200: the training path does not produce a gradient wrt the 2-D input keypoints (the '
                           'reference never asks for one, run.py:458-485); detach() the input or keep the module that '
                           'produces it out of the autograd graph')
    x = ops.f32c(x)
    if not torch.is_grad_enabled():
        with torch.no_grad():
            return _forward_stack(model, x, dt)[0]
    return _StackTrainFn.apply(model, dt, x, *stack_parameters(model))
