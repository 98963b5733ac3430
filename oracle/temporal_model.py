"""CPU oracle for the TemporalModel / TemporalModelOptimized1f stack.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this module;
the product path (dynamic-camera-augmented-videopose3d_b200/) never does.

This is a *functional restatement* of /root/reference/common/models/TemporalModel.py: the model is a plain dict
of tensors with the reference's state_dict keys (TemporalModel.py:32-33,102,113-124 -> SURVEY 8a-0) and the forward
is spelled out with the same torch CPU primitives the reference's nn.Modules dispatch to (F.conv1d, F.batch_norm,
F.relu), in the same order. Backward comes from torch autograd over this functional forward, exactly as the
reference obtains it (run.py:485).

Parity pin: tests/golden/*.npz hold inputs / state_dicts / outputs produced by importing the real reference in the
build container (tests/golden/make_golden.py); tests/test_oracle_golden.py checks this restatement against them.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm1d default, TemporalModel.py:32,117,119


def make_plan(filter_widths, causal=False, dense=False, strided=False):
    """Per-layer geometry. Mirrors TemporalModel.__init__ (TemporalModel.py:85-124) and
    TemporalModelOptimized1f.__init__ (:152-186)."""
    for fw in filter_widths:
        assert fw % 2 != 0, 'Only odd filter widths are supported'  # TemporalModel.py:20-21
    pad = [filter_widths[0] // 2]
    causal_shift = [filter_widths[0] // 2 if causal else 0]
    blocks = []
    dilation = filter_widths[0]
    for fw in filter_widths[1:]:
        p = (fw - 1) * dilation // 2
        pad.append(p)
        if strided:
            causal_shift.append(fw // 2 if causal else 0)  # :177
            blocks.append(dict(taps=fw, dilation=1, stride=fw))
        else:
            causal_shift.append(fw // 2 * dilation if causal else 0)  # :111
            if dense:
                blocks.append(dict(taps=2 * p + 1, dilation=1, stride=1))  # :113-116
            else:
                blocks.append(dict(taps=fw, dilation=dilation, stride=1))
        dilation *= fw
    return dict(filter_widths=list(filter_widths), pad=pad, causal_shift=causal_shift, blocks=blocks,
                strided=strided, expand_stride=filter_widths[0] if strided else 1)


def receptive_field(plan):
    return 1 + 2 * sum(plan['pad'])  # TemporalModel.py:40-47


def init_state(num_joints_in, in_features, num_joints_out, filter_widths, channels=1024, dense=False, seed=0,
               randomize_bn=True):
    """Random parameters with the reference's shapes and nn.Conv1d default init distribution
    (kaiming-uniform, bound 1/sqrt(fan_in)); BN statistics randomised so that folding bugs cannot hide."""
    g = torch.Generator().manual_seed(seed)
    plan = make_plan(filter_widths, dense=dense)
    sd = {}

    def conv_w(cout, cin, k):
        bound = 1.0 / (cin * k) ** 0.5
        return (torch.rand(cout, cin, k, generator=g) * 2 - 1) * bound

    def bn(prefix):
        if randomize_bn:
            sd[prefix + '.weight'] = torch.rand(channels, generator=g) + 0.5
            sd[prefix + '.bias'] = torch.randn(channels, generator=g) * 0.1
            sd[prefix + '.running_mean'] = torch.randn(channels, generator=g) * 0.1
            sd[prefix + '.running_var'] = torch.rand(channels, generator=g) + 0.5
        else:
            sd[prefix + '.weight'] = torch.ones(channels)
            sd[prefix + '.bias'] = torch.zeros(channels)
            sd[prefix + '.running_mean'] = torch.zeros(channels)
            sd[prefix + '.running_var'] = torch.ones(channels)
        sd[prefix + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.int64)

    bn('expand_bn')
    sd['shrink.weight'] = conv_w(num_joints_out * 3, channels, 1)
    sd['shrink.bias'] = (torch.rand(num_joints_out * 3, generator=g) * 2 - 1) / channels ** 0.5
    sd['expand_conv.weight'] = conv_w(channels, num_joints_in * in_features, filter_widths[0])
    for i, blk in enumerate(plan['blocks']):
        sd['layers_conv.%d.weight' % (2 * i)] = conv_w(channels, channels, blk['taps'])
        sd['layers_conv.%d.weight' % (2 * i + 1)] = conv_w(channels, channels, 1)
    for i in range(2 * len(plan['blocks'])):
        bn('layers_bn.%d' % i)
    return sd


def _bn(x, sd, prefix, training, momentum, new_stats):
    w, b = sd[prefix + '.weight'], sd[prefix + '.bias']
    rm, rv = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
    if training:
        rm2, rv2 = rm.clone(), rv.clone()
        y = F.batch_norm(x, rm2, rv2, w, b, True, momentum, BN_EPS)
        if new_stats is not None:
            new_stats[prefix + '.running_mean'] = rm2
            new_stats[prefix + '.running_var'] = rv2
            new_stats[prefix + '.num_batches_tracked'] = sd[prefix + '.num_batches_tracked'] + 1
        return y
    return F.batch_norm(x, rm, rv, w, b, False, momentum, BN_EPS)


def forward(sd, x, filter_widths, causal=False, dense=False, strided=False, training=False, momentum=0.1,
            new_stats=None):
    """(N, T, J, F) -> (N, T', J_out, 3).  TemporalModelBase.forward (TemporalModel.py:62-76) around
    TemporalModel._forward_blocks (:126-138) or TemporalModelOptimized1f._forward_blocks (:188-198).
    Dropout is the identity here (p = 0 / eval); the reference's Philox stream cannot be reproduced."""
    assert x.dim() == 4
    plan = make_plan(filter_widths, causal, dense, strided)
    n, t = x.shape[0], x.shape[1]
    assert x.shape[2] * x.shape[3] == sd['expand_conv.weight'].shape[1]
    h = x.reshape(n, t, -1).permute(0, 2, 1)
    h = F.conv1d(h, sd['expand_conv.weight'], None, stride=plan['expand_stride'])
    h = F.relu(_bn(h, sd, 'expand_bn', training, momentum, new_stats))
    for i, blk in enumerate(plan['blocks']):
        pad, shift = plan['pad'][i + 1], plan['causal_shift'][i + 1]
        if strided:
            fw = plan['filter_widths'][i + 1]
            res = h[:, :, shift + fw // 2::fw]
        else:
            res = h[:, :, pad + shift: h.shape[2] - pad + shift]
        h = F.conv1d(h, sd['layers_conv.%d.weight' % (2 * i)], None, stride=blk['stride'], dilation=blk['dilation'])
        h = F.relu(_bn(h, sd, 'layers_bn.%d' % (2 * i), training, momentum, new_stats))
        h = F.conv1d(h, sd['layers_conv.%d.weight' % (2 * i + 1)], None)
        h = res + F.relu(_bn(h, sd, 'layers_bn.%d' % (2 * i + 1), training, momentum, new_stats))
    h = F.conv1d(h, sd['shrink.weight'], sd['shrink.bias'])
    j_out = sd['shrink.weight'].shape[0] // 3
    return h.permute(0, 2, 1).reshape(n, -1, j_out, 3)


def forward_lowp(sd, x, filter_widths, causal=False, dense=False, strided=False, dtype=torch.float16):
    """Eval forward with every GEMM operand (activations and weights) rounded to `dtype` and fp32 accumulation --
    a CPU emulation of the tensor-core path, used to derive the parity tolerances written in the GPU tests."""
    def rnd(v):
        return v.to(dtype).to(torch.float32)

    def fold(prefix):
        s = sd[prefix + '.weight'] / torch.sqrt(sd[prefix + '.running_var'] + BN_EPS)
        return s.view(1, -1, 1), (sd[prefix + '.bias'] - sd[prefix + '.running_mean'] * s).view(1, -1, 1)

    plan = make_plan(filter_widths, causal, dense, strided)
    n, t = x.shape[0], x.shape[1]
    h = rnd(x.reshape(n, t, -1).permute(0, 2, 1))
    s, b = fold('expand_bn')
    h = rnd(F.relu(F.conv1d(h, rnd(sd['expand_conv.weight']), None, stride=plan['expand_stride']) * s + b))
    for i, blk in enumerate(plan['blocks']):
        pad, shift = plan['pad'][i + 1], plan['causal_shift'][i + 1]
        if strided:
            fw = plan['filter_widths'][i + 1]
            res = h[:, :, shift + fw // 2::fw]
        else:
            res = h[:, :, pad + shift: h.shape[2] - pad + shift]
        s, b = fold('layers_bn.%d' % (2 * i))
        h = rnd(F.relu(F.conv1d(h, rnd(sd['layers_conv.%d.weight' % (2 * i)]), None, stride=blk['stride'],
                                dilation=blk['dilation']) * s + b))
        s, b = fold('layers_bn.%d' % (2 * i + 1))
        h = rnd(res + F.relu(F.conv1d(h, rnd(sd['layers_conv.%d.weight' % (2 * i + 1)]), None) * s + b))
    h = F.conv1d(h, rnd(sd['shrink.weight']), sd['shrink.bias'])
    j_out = sd['shrink.weight'].shape[0] // 3
    return h.permute(0, 2, 1).reshape(n, -1, j_out, 3)


def train_step_grads(sd, x, target, filter_widths, causal=False, strided=True, momentum=0.1):
    """One training forward+backward with dropout = 0: returns (loss, prediction, grads dict, new BN statistics).
    Restates run.py:473-485 (forward, mpjpe, backward) for the TemporalModel family."""
    from . import loss as oloss
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.dtype.is_floating_point and 'running_' not in k}
    full = dict(sd)
    full.update(params)
    new_stats = {}
    pred = forward(full, x, filter_widths, causal=causal, strided=strided, training=True, momentum=momentum,
                   new_stats=new_stats)
    loss = oloss.mpjpe(pred, target)
    loss.backward()
    grads = {k: v.grad.detach() for k, v in params.items()}
    return loss.detach(), pred.detach(), grads, new_stats


def _ste_round(v, dtype):
    """Value rounded to `dtype`, gradient passed straight through (the GPU backward treats operand rounding the same
    way: it differentiates the fp32 formulas and only rounds what it stores)."""
    return v + (v.to(dtype).to(torch.float32) - v).detach()


def forward_lowp_train(sd, x, filter_widths, causal=False, strided=True, dtype=torch.float16, masks=None, dense=False,
                       expand_unrounded=True):
    """Train-mode forward with the rounding points of the CUDA training path emulated on the CPU (see
    train_step_grads_lowp). Returns (prediction with grad_fn, dict of leaf parameters)."""
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.dtype.is_floating_point and 'running_' not in k}
    plan = make_plan(filter_widths, causal, dense, strided)
    rnd = lambda v: _ste_round(v, dtype)
    mask_iter = iter(masks) if masks is not None else None

    def bn_act(z, prefix, res=None, round_z=True):
        mean = z.mean(dim=(0, 2), keepdim=True)
        var = z.var(dim=(0, 2), unbiased=False, keepdim=True)
        scale = params[prefix + '.weight'].view(1, -1, 1) / torch.sqrt(var + BN_EPS)
        shift = params[prefix + '.bias'].view(1, -1, 1) - mean * scale
        pre = (rnd(z) if round_z else z) * scale + shift
        y = F.relu(pre) if mask_iter is None else pre * next(mask_iter).to(pre.dtype)
        return rnd(y if res is None else y + res)

    n, t = x.shape[0], x.shape[1]
    h = rnd(x.reshape(n, t, -1).permute(0, 2, 1))
    # the expand layer's GEMM epilogue applies BatchNorm to the fp32 accumulator (its raw output is never stored), every
    # other layer stores the raw output in the operand type first
    h = bn_act(F.conv1d(h, rnd(params['expand_conv.weight']), None, stride=plan['expand_stride']), 'expand_bn',
               round_z=not expand_unrounded)
    for i, blk in enumerate(plan['blocks']):
        pad, shift = plan['pad'][i + 1], plan['causal_shift'][i + 1]
        if strided:
            fw = plan['filter_widths'][i + 1]
            res = h[:, :, shift + fw // 2::fw]
        else:
            res = h[:, :, pad + shift: h.shape[2] - pad + shift]
        h = bn_act(F.conv1d(h, rnd(params['layers_conv.%d.weight' % (2 * i)]), None, stride=blk['stride'],
                            dilation=blk['dilation']), 'layers_bn.%d' % (2 * i))
        h = bn_act(F.conv1d(h, rnd(params['layers_conv.%d.weight' % (2 * i + 1)]), None), 'layers_bn.%d' % (2 * i + 1),
                   res=res)
    h = F.conv1d(h, rnd(params['shrink.weight']), params['shrink.bias'])
    j_out = sd['shrink.weight'].shape[0] // 3
    return h.permute(0, 2, 1).reshape(n, -1, j_out, 3), params


def train_step_grads_lowp(sd, x, target, filter_widths, causal=False, strided=True, dtype=torch.float16, masks=None):
    """train_step_grads with the *forward* rounding points of the CUDA training path emulated on the CPU: GEMM
    operands (input, weights) and every stored matrix (raw convolution output z, activation a) rounded to `dtype`,
    batch statistics taken from the unrounded fp32 accumulators, fp32 everywhere else. Gradients of a ReLU network
    are discontinuous in the activations -- a pre-activation that rounding moves across zero flips a mask bit -- so
    the fp32 oracle is only a loose bound for low-precision gradients. Even this emulation differs from the GPU in
    fp32 summation order, which moves a few stored values by one 16-bit ulp and flips a few more bits; `masks` (one
    bool tensor (N, C, T') per BatchNorm layer, in forward order, taken from the GPU's own saved pre-activations) pins
    the ReLU pattern so that the comparison isolates the backward arithmetic. Used by tests/ only."""
    from . import loss as oloss
    pred, params = forward_lowp_train(sd, x, filter_widths, causal, strided, dtype, masks)
    loss = oloss.mpjpe(pred, target)
    loss.backward()
    grads = {k: v.grad.detach() for k, v in params.items()}
    return loss.detach(), pred.detach(), grads
