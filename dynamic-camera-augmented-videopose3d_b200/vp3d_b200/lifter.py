"""StackedPoseLifter on the convolution-stack kernels (reference common/models/StackedPoseLifter.py:37-56,
called from run.py:474-478 on the outputs of a TemporalModel and a camera-aware model).

The model is an MLP, (Linear -> ReLU -> Dropout) x (1 + num_layers) -> Linear, on the concatenation of two (B, J * F)
pose vectors. A Linear is a 1-tap convolution, so every layer is one launch of the kernels the temporal stack uses:

  forward   a_i = drop(relu(a_{i-1} W_i^T + b_i))   K1 / K1p (vp3d_conv_block_fwd): bias through the epilogue's shift,
                                                    ReLU and the counter-based dropout mask in the epilogue (EPI = 1)
  backward  dz_i = (dz_{i+1} W_{i+1}) * [a_i > 0] / (1 - p)   data-gradient GEMM with the gated epilogue (side_mode 2:
                                                    where the stored activation is 0 the element was clipped or dropped)
            dW_i = dz_i^T a_{i-1}                   vp3d_wgrad (+ layout pass)
            db_i = column sums of dz_i              vp3d_col_stats
No activation is recomputed and no mask is stored. CUDA tensors only.
"""
import torch

from . import native, ops
from .temporal import K_ALIGN, N_TILE, N_TILE_NARROW, LayerPlan, _round_up, _run_layer, resolve_dtype
from .training import SHRINK_PAD, _step_counter


debug_keep_saved = False   # tests: keep the activations the last training forward saved in `debug_last_acts`
debug_last_acts = None


def _linears(model):
    return [m for m in model.mlp_layers if isinstance(m, torch.nn.Linear)]


def _lin_w(dt, lin, rows_pad, k_pad):
    """Packed operand [rows_pad][k_pad] of a Linear weight, cached on the parameter per (operand type, padding, version)."""
    w = lin.weight
    reg = w.__dict__.setdefault('_vp3d_packed_lin', {})
    key = (dt, rows_pad, k_pad)
    entry = reg.get(key)
    if entry is not None and entry[1] == w._version and entry[0].device == w.device:
        return entry[0]
    packed = ops.pack_conv_weight(dt, w.detach().unsqueeze(-1), rows_pad, k_pad)
    reg[key] = [packed, w._version]
    return packed


def _bias(lin, n_pad):
    return torch.nn.functional.pad(lin.bias.detach().float(), (0, n_pad - lin.out_features))


def _forward(model, x, dt, train):
    """x: (B, 2 * J * F) fp32 CUDA -> (y (B, J * F) fp32, saved activations [a_0 .. a_last], dropout descriptions)."""
    lins = _linears(model)
    b, k0 = x.shape
    assert k0 == lins[0].in_features, 'expected %d input features, got %d' % (lins[0].in_features, k0)
    k_pad = _round_up(k0, K_ALIGN)
    h = ops.pack_rows(dt, x, k_pad).view(1, b, k_pad)
    acts, drops = [h], []
    p = float(model.dropout.p) if train else 0.0
    counter = None
    if p > 0:
        counter = _step_counter(model, x.device)
        ops.counter_add(counter, 1)
        counter = counter.clone()
    for i, lin in enumerate(lins[:-1]):
        c_pad = _round_up(lin.out_features, N_TILE)
        w = _lin_w(dt, lin, c_pad, k_pad)
        drop = ops.make_dropout(p, torch.initial_seed(), 1000 + i, counter) if p > 0 else None
        h, _ = _run_layer(dt, h, 1, b, k_pad, w, LayerPlan(1), None, _bias(lin, c_pad), True, drop=drop)
        acts.append(h)
        drops.append(drop)
        k_pad = c_pad
    last = lins[-1]
    n_out_pad = _round_up(last.out_features, N_TILE_NARROW)
    w_last = _lin_w(dt, last, n_out_pad, k_pad)
    y, _ = _run_layer(dt, h, 1, b, k_pad, w_last, LayerPlan(1), None, _bias(last, n_out_pad), False, out_f32=True,
                      n_valid=last.out_features, block_n=N_TILE_NARROW)
    return y.view(b, last.out_features), acts, w_last, p


class _LifterFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, dt, x, *params):
        y, acts, w_last, p = _forward(model, x, dt, True)
        if debug_keep_saved:
            global debug_last_acts
            debug_last_acts = acts
        ctx.model, ctx.dt, ctx.acts, ctx.w_last, ctx.p = model, dt, acts, w_last, p
        ctx.versions = [(lin.weight, lin.weight._version) for lin in _linears(model)]
        ctx.params = params
        return y

    @staticmethod
    def backward(ctx, dy):
        model, dt, acts = ctx.model, ctx.dt, ctx.acts
        if acts is None:
            raise RuntimeError('vp3d_b200: the saved activations of this forward were released by its first backward')
        for w, version in ctx.versions:
            if w._version != version:
                raise RuntimeError('vp3d_b200: a StackedPoseLifter weight was modified between the forward and its '
                                   'backward; run backward() before optimizer.step(), as run.py:485-487 does')
        lins = _linears(model)
        dev = dy.device
        b = acts[0].shape[1]
        keep = ops.keep_scale(ctx.p)
        grads = {}
        zeros = lambda shape, dtype=torch.float32: torch.zeros(shape, dtype=dtype, device=dev)

        # ---- output layer: y = a_last W^T + b
        last = lins[-1]
        n_out = last.out_features
        dy2 = ops.f32c(dy).reshape(b, n_out)
        gscale = ops.grad_scale(dy2)
        dz, dbias = ops.grad_pack_rows(dt, dy2, SHRINK_PAD, gscale, want_col_sum=True)
        grads[id(last.bias)] = dbias
        a_in = acts[-1]
        c_pad = a_in.shape[-1]
        packed = zeros((1, SHRINK_PAD, c_pad))
        ops.wgrad(dt, dz, (1, b, SHRINK_PAD, b * SHRINK_PAD), a_in, (b, c_pad, c_pad, b * c_pad), SHRINK_PAD, c_pad, 1,
                  packed, block_n=256 if c_pad % 256 == 0 else 64)
        grads[id(last.weight)] = ops.wgrad_finish(packed, n_out, last.in_features, 1, SHRINK_PAD, c_pad,
                                                  gscale).view(n_out, last.in_features)
        # gradient wrt the last hidden PRE-activation: the data-gradient GEMM gates its result by a_last > 0
        g = torch.empty((1, b, c_pad), dtype=dz.dtype, device=dev)
        ops.conv_block(dt, dz, (1, b, SHRINK_PAD, SHRINK_PAD, b * SHRINK_PAD), ctx.w_last, 1, 0, ctx.w_last.shape[0], b, g,
                       (c_pad, b * c_pad), w_mn_major=(c_pad, 0), side=a_in, side_view=(c_pad, b * c_pad, b, 0),
                       side_mode=2, side_scale=keep)

        # ---- hidden layers, last to first; `g` = dz_i
        for i in range(len(lins) - 2, -1, -1):
            lin = lins[i]
            c_pad = g.shape[-1]
            a_in = acts[i]
            k_pad = a_in.shape[-1]
            stats = zeros((2, c_pad), torch.float64)
            ops.col_stats(dt, g.view(b, c_pad), stats)
            grads[id(lin.bias)] = (stats[0, :lin.out_features] * gscale[1]).float()
            packed = zeros((1, c_pad, k_pad))
            ops.wgrad(dt, g, (1, b, c_pad, b * c_pad), a_in, (b, k_pad, k_pad, b * k_pad), c_pad, k_pad, 1, packed,
                      block_n=256 if k_pad % 256 == 0 else 64)
            grads[id(lin.weight)] = ops.wgrad_finish(packed, lin.out_features, lin.in_features, 1, c_pad, k_pad,
                                                     gscale).view(lin.out_features, lin.in_features)
            if i == 0:
                break     # no gradient wrt the two input poses: run.py:474-477 computes them under torch.no_grad()
            g_in = torch.empty((1, b, k_pad), dtype=g.dtype, device=dev)
            ops.conv_block(dt, g, (1, b, c_pad, c_pad, b * c_pad), _lin_w(dt, lin, c_pad, k_pad), 1, 0, c_pad, b, g_in,
                           (k_pad, b * k_pad), w_mn_major=(k_pad, 0), side=a_in, side_view=(k_pad, b * k_pad, b, 0),
                           side_mode=2, side_scale=keep)
            g = g_in
        ctx.acts = None
        return (None, None, None) + tuple(grads.get(id(p)) for p in ctx.params)


def lifter_parameters(model):
    ps = []
    for lin in _linears(model):
        ps += [lin.weight, lin.bias]
    return ps


def forward(model, input_3d_transformer, input_3d_fcn):
    """StackedPoseLifter.forward (:37-56): (B, ..., J, F) x 2 -> (B, 1, J, F)."""
    ops.require_cuda(input_3d_transformer, input_3d_fcn)
    dt = resolve_dtype(getattr(model, 'operand_dtype', None))
    if dt == native.TF32:
        dt = native.F16 if model.training else dt
    a = input_3d_transformer.reshape(input_3d_transformer.size(0), -1)        # :47
    b = input_3d_fcn.reshape(input_3d_fcn.size(0), -1)                          # :48
    x = ops.f32c(torch.cat((a, b), dim=-1))                                     # :50
    if x.requires_grad and torch.is_grad_enabled():
        raise RuntimeError('vp3d_b200: StackedPoseLifter does not produce a gradient wrt its two input poses (run.py:474-477 '
                           'computes them under torch.no_grad()); detach() them')
    if model.training and torch.is_grad_enabled():
        y = _LifterFn.apply(model, dt, x, *lifter_parameters(model))
    else:
        with torch.no_grad():
            y = _forward(model, x, dt, model.training)[0]
    return y.view(y.size(0), 1, model.num_joints, model.features)               # :55
