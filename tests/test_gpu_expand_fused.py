"""The fused expand layer (csrc/expand.cu) and the fused epilogue of the CTA-pair kernel (conv_gemm2.cu, EPI = 1):
BatchNorm statistics from the Gram matrix of the layer input, BatchNorm + ReLU + dropout applied by the GEMM epilogue,
the gated data gradient (side_mode 2), the TMA residual path (side_mode 1), and the whole training step against the
unfused path (bn_finalize + bn_act_fwd / bn_act_bwd) that every other layer -- and the oracle comparison -- uses."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from common.loss import mpjpe  # noqa: E402
from common.models.TemporalModel import TemporalModel, TemporalModelOptimized1f  # noqa: E402
from oracle import temporal_model as otm  # noqa: E402
from vp3d_b200 import native, ops, training  # noqa: E402


def rel_err(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _bn(c, seed):
    g = torch.Generator().manual_seed(seed)
    bn = torch.nn.BatchNorm1d(c).cuda()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(c, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(c, generator=g) * 0.1)
    return bn


@pytest.mark.parametrize('strided', [True, False])
def test_gram_statistics_match_the_statistics_of_the_layer_output(strided):
    """scale / shift / mean / invstd / running statistics from X^T X against fp64 statistics of the actual convolution
    output (the rounded operands contracted exactly); the input has a large mean so that E[z^2] - E[z]^2 cancels."""
    dt, c_in, c_in_pad, c, taps = native.F16, 34, 64, 1024, 3
    n, t_in = 24, 81 if strided else 50
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(n, t_in, c_in, generator=g) * 2 - 1 + 0.7).cuda()
    w32 = (torch.randn(c, c_in, taps, generator=g) / (c_in * taps) ** 0.5).cuda()
    h = ops.pack_rows(dt, x.reshape(n * t_in, c_in), c_in_pad, ones_col=c_in).view(n, t_in, c_in_pad)
    assert torch.all(h[..., c_in] == 1) and torch.all(h[..., c_in + 1:] == 0)
    w = ops.pack_conv_weight(dt, w32, c, c_in_pad)
    k_total = taps * c_in_pad
    t_out = t_in // taps if strided else t_in - taps + 1
    if strided:
        xv, av = (1, n * t_out, k_total, n * t_in * c_in_pad), (n * t_out, k_total, k_total, n * t_in * c_in_pad)
    else:
        xv, av = (n, t_out, c_in_pad, t_in * c_in_pad), (t_out, k_total, c_in_pad, t_in * c_in_pad)
    gram = torch.zeros(1, 256, 256, device='cuda')
    ops.wgrad(dt, h, xv, h, av, 256, 256, 1, gram, block_n=64, dz_cols=k_total)
    # exact reference of the Gram matrix and of the layer output from the rounded operands
    hd = h.double().cpu()
    if strided:
        X = hd.reshape(n * t_out, k_total)
    else:
        X = torch.stack([hd[:, k:k + t_out] for k in range(taps)], dim=2).reshape(n * t_out, k_total)
    G_ref = X.T @ X
    assert rel_err(gram[0, :k_total, :k_total], G_ref) < 1e-5
    assert gram[0, c_in, c_in].item() == n * t_out
    assert gram[0, k_total:].abs().max().item() == 0 and gram[0, :, k_total:].abs().max().item() == 0
    z = X @ w.double().cpu().T                                       # (rows, c)
    bn = _bn(c, 4)
    scale, shift, mean, invstd, wg = ops.expand_bn_stats(dt, gram, w, k_total, c_in, bn, c)
    torch.cuda.synchronize()
    m_ref, v_ref = z.mean(0), z.var(0, unbiased=False)
    assert (mean.double().cpu() - m_ref).abs().max().item() < 1e-5
    assert rel_err(invstd, 1 / torch.sqrt(v_ref + bn.eps)) < 2e-5
    sc_ref = bn.weight.double().cpu() / torch.sqrt(v_ref + bn.eps)
    assert rel_err(scale, sc_ref) < 2e-5
    assert (shift.double().cpu() - (bn.bias.double().cpu() - m_ref * sc_ref)).abs().max().item() < 2e-5
    assert rel_err(wg[:, :k_total], w.double().cpu() @ G_ref) < 1e-5
    rows = n * t_out
    assert (bn.running_mean.double().cpu() - 0.1 * m_ref).abs().max().item() < 1e-5
    assert rel_err(bn.running_var, 0.9 + 0.1 * v_ref * rows / (rows - 1)) < 2e-5
    assert int(bn.num_batches_tracked.item()) == 1


@pytest.mark.parametrize('dtype', ['fp16', 'bf16'])
def test_epilogue_dropout_draws_the_mask_of_the_stand_alone_pass(dtype):
    """conv -> scale/shift -> ReLU -> dropout in the GEMM epilogue against the raw GEMM followed by vp3d_bn_act_fwd:
    the same Philox mask element for element, values equal up to the rounding of the stored raw output."""
    dt = native.DTYPE_NAMES[dtype]
    td = ops.torch_dtype(dt)
    g = torch.Generator().manual_seed(7)
    for seqs, rows, c, n in [(1, 777, 192, 1024), (3, 333, 64, 512)]:
        a = (torch.randn(seqs, rows, c, generator=g) * 0.5).to(td).cuda()
        w = (torch.randn(n, c, generator=g) / c ** 0.5).to(td).cuda()
        scale = (torch.rand(n, generator=g) + 0.5).cuda()
        shift = (torch.randn(n, generator=g) * 0.2).cuda()
        d = ops.make_dropout(0.25, 99, 3)
        fused = torch.full((seqs, rows, n), float('nan'), dtype=td, device='cuda')
        ops.conv_block(dt, a, (seqs, rows, c, c, rows * c), w, 1, 0, c, rows, fused, (n, rows * n), scale=scale,
                       shift=shift, relu=True, drop=d)
        z = torch.empty((seqs, rows, n), dtype=td, device='cuda')
        ops.conv_block(dt, a, (seqs, rows, c, c, rows * c), w, 1, 0, c, rows, z, (n, rows * n))
        ref = ops.bn_act_fwd(dt, z.view(seqs * rows, n), scale, shift, seqs, rows, d).view(seqs, rows, n)
        torch.cuda.synchronize()
        assert torch.isfinite(fused.float()).all()
        pre = z.float() * scale + shift
        sure = pre.abs() > 2e-2                      # away from the ReLU boundary the two paths must agree on the mask
        assert torch.equal((fused > 0) & sure, (ref > 0) & sure)
        kept = (fused > 0).float().sum() / (pre > 0).float().sum()
        assert abs(kept.item() - 0.75) < 0.01
        assert (fused.float() - ref.float())[sure].abs().max().item() < (2e-2 if dtype == 'fp16' else 1.2e-1)
        assert rel_err(fused.float()[sure], ref.float()[sure]) < (1e-3 if dtype == 'fp16' else 8e-3)


def test_side_input_add_equals_the_register_residual_path():
    """side_mode 1 (residual tile through TMA into the staging buffer) against the generic `res` path: same bits."""
    dt, td = native.F16, torch.float16
    g = torch.Generator().manual_seed(11)
    for seqs, rows, c, n, off in [(2, 700, 1024, 1024, 81), (1, 130, 64, 256, 0), (5, 64, 128, 512, 3)]:
        a = (torch.randn(seqs, rows, c, generator=g) * 0.5).to(td).cuda()
        w = (torch.randn(n, c, generator=g) / c ** 0.5).to(td).cuda()
        shift = (torch.randn(n, generator=g) * 0.2).cuda()
        res = (torch.randn(seqs, rows + 2 * off, n, generator=g) * 0.5).to(td).cuda()
        outs = []
        for side in (False, True):
            out = torch.full((seqs, rows, n), float('nan'), dtype=td, device='cuda')
            kw = (dict(side=res, side_view=(n, (rows + 2 * off) * n, rows + 2 * off, off), side_mode=1) if side else
                  dict(res=res, res_view=(n, (rows + 2 * off) * n, 1, off)))
            native.check(native.lib().vp3d_set_pair_mode(2), 'pair')
            try:
                ops.conv_block(dt, a, (seqs, rows, c, c, rows * c), w, 1, 0, c, rows, out, (n, rows * n), shift=shift,
                               relu=True, **kw)
            finally:
                native.check(native.lib().vp3d_set_pair_mode(1), 'pair')
            outs.append(out)
        torch.cuda.synchronize()
        assert torch.isfinite(outs[1].float()).all()
        assert torch.equal(outs[0], outs[1])


def test_gated_data_gradient():
    """side_mode 2: out = side > 0 ? x * side_scale : 0 on the data-gradient GEMM (MN-major weights), with and without
    a residual fan-in, against the ungated launch."""
    dt, td = native.F16, torch.float16
    g = torch.Generator().manual_seed(13)
    rows, co, n_cols = 900, 1024, 3072              # dz [rows][co] x W^T -> [rows][3 * 1024]
    dz = (torch.randn(1, rows, co, generator=g) * 0.5).to(td).cuda()
    w = (torch.randn(co, n_cols, generator=g) / co ** 0.5).to(td).cuda()      # forward-packed [c_out][taps * c_in]
    act = torch.relu(torch.randn(1, rows, n_cols, generator=g)).to(td).cuda()
    act[0, :, ::7] = 0
    fan = (torch.randn(1, rows, 1024, generator=g) * 0.3).to(td).cuda()
    ks = 256.0 / 192.0
    for with_fan in (False, True):
        kw = dict(res=fan, res_view=(1024, rows * 1024, 1, 0), res_col_off=1024, res_cols=1024) if with_fan else {}
        plain = torch.empty((1, rows, n_cols), dtype=td, device='cuda')
        gated = torch.full((1, rows, n_cols), float('nan'), dtype=td, device='cuda')
        native.check(native.lib().vp3d_set_pair_mode(2), 'pair')
        try:
            ops.conv_block(dt, dz, (1, rows, co, co, rows * co), w, 1, 0, co, rows, plain, (n_cols, rows * n_cols),
                           w_mn_major=(n_cols, 0), **kw)
        finally:
            native.check(native.lib().vp3d_set_pair_mode(1), 'pair')
        ops.conv_block(dt, dz, (1, rows, co, co, rows * co), w, 1, 0, co, rows, gated, (n_cols, rows * n_cols),
                       w_mn_major=(n_cols, 0), side=act, side_view=(n_cols, rows * n_cols, rows, 0), side_mode=2,
                       side_scale=ks, **kw)
        torch.cuda.synchronize()
        assert torch.isfinite(gated.float()).all()
        assert torch.all(gated[act == 0] == 0)
        on = act > 0
        # the ungated result went through one extra rounding to fp16 before the comparison multiplies it
        assert rel_err(gated.float()[on], plain.float()[on] * ks) < 6e-4


def _grads(cls, sd, fw, x, tgt, fused, dropout, dtype='fp16', **kw):
    training.fused_expand = fused
    try:
        torch.manual_seed(5)
        m = cls(17, 2, 17, fw, dropout=dropout, channels=sd['expand_conv.weight'].shape[0], **kw)
        m.load_state_dict(sd, strict=True)
        m = m.cuda().train()
        m.operand_dtype = dtype
        pred = m(x.cuda())
        loss = mpjpe(pred, tgt.cuda())
        loss.backward()
        torch.cuda.synchronize()
        bufs = {k: v.detach().clone() for k, v in m.named_buffers()}
        return pred.detach(), loss.item(), {k: p.grad.clone() for k, p in m.named_parameters()}, bufs
    finally:
        training.fused_expand = True


@pytest.mark.parametrize('dropout', [0.0, 0.25])
@pytest.mark.parametrize('name,cls,t_in,kw', [('1f', TemporalModelOptimized1f, 27, {}), ('full', TemporalModel, 40, {}),
                                               ('1f_causal', TemporalModelOptimized1f, 27, {'causal': True})])
def test_training_step_fused_expand_equals_unfused(name, cls, t_in, kw, dropout):
    """The whole step with the fused expand layer against the same step with bn_finalize / bn_act_fwd / bn_act_bwd on
    that layer: same dropout masks (counter based), so predictions, loss and the running statistics agree up to the
    rounding of the one stored matrix (z) the fused path never rounds. Gradients get the loose bound of a ReLU network
    whose forward was perturbed by one 16-bit ulp here and there (mask flips in EVERY later layer, sqrt(f) relative each
    -- the deeper layers' gradients do not depend on the expand layer's backward at all and move just as much); the
    tight check of the fused backward arithmetic is the mask-pinned emulation in test_gpu_training.py."""
    fw = [3, 3, 3]
    sd = otm.init_state(17, 2, 17, fw, channels=1024, seed=31)
    g = torch.Generator().manual_seed(32)
    n = 96
    x = torch.rand(n, t_in, 17, 2, generator=g) * 2 - 1 + 0.3
    tgt = torch.randn(n, t_in - 26, 17, 3, generator=g) * 0.3
    p0, l0, g0, b0 = _grads(cls, sd, fw, x, tgt, False, dropout, **kw)
    p1, l1, g1, b1 = _grads(cls, sd, fw, x, tgt, True, dropout, **kw)
    assert rel_err(p1, p0) < 2e-3
    assert abs(l1 - l0) < 1e-3 * abs(l0)
    worst = {k: rel_err(g1[k], g0[k]) for k in g0}
    print(name, dropout, 'fused vs unfused grad rel errs', {k: '%.2e' % v for k, v in worst.items()})
    assert max(worst.values()) < 8e-2, worst
    deeper = max(v for k, v in worst.items() if k.startswith('layers_'))
    for k in ('expand_conv.weight', 'expand_bn.weight', 'expand_bn.bias'):
        assert worst[k] < max(2 * deeper, 1e-2), (k, worst)       # the fused layer is no worse than the untouched ones
    for k in b0:
        if 'num_batches' in k:
            assert int(b0[k]) == int(b1[k]) == 1
        else:
            assert (b0[k] - b1[k]).abs().max().item() < 2e-3, k


def test_fused_expand_against_fp32_oracle_243():
    """243-frame 1f model, batch 64, dropout 0: expand-layer gradients and statistics against the fp32 CPU oracle."""
    fw = [3, 3, 3, 3, 3]
    sd = otm.init_state(17, 2, 17, fw, channels=1024, seed=41)
    g = torch.Generator().manual_seed(42)
    x = torch.rand(64, 243, 17, 2, generator=g) * 2 - 1
    tgt = torch.randn(64, 1, 17, 3, generator=g) * 0.3
    loss_o, pred_o, grads_o, stats_o = otm.train_step_grads(sd, x, tgt, fw, strided=True)
    pred, loss, grads, bufs = _grads(TemporalModelOptimized1f, sd, fw, x, tgt, True, 0.0)
    assert rel_err(pred, pred_o) < 2e-3
    assert abs(loss - loss_o.item()) < 1e-3 * loss_o.item()
    worst = {k: rel_err(grads[k], grads_o[k]) for k in grads}
    print('fused expand vs fp32 oracle', {k: '%.2e' % v for k, v in worst.items() if 'expand' in k})
    assert max(worst.values()) < 1.2e-1, worst
    for k in ('expand_bn.running_mean', 'expand_bn.running_var'):
        assert (bufs[k].cpu() - stats_o[k]).abs().max().item() < 1e-4, k     # analytic statistics: fp32-accurate


@pytest.mark.parametrize('pair', [False, True])
@pytest.mark.parametrize('momentum', [0.1, None])
def test_finalize_in_the_gemm_tail_is_bit_identical_to_the_stand_alone_launch(pair, momentum):
    """vp3d_conv_args.fin (the last CTA of the GEMM finalizes the BatchNorm statistics) against the GEMM followed by
    vp3d_bn_finalize: same sums, same arithmetic -> the same bits, running statistics and counter included; also
    nn.BatchNorm1d(momentum=None) (cumulative average) over two calls."""
    dt, td = native.F16, torch.float16
    g = torch.Generator().manual_seed(17)
    seqs, rows, c, n = (3, 700, 256, 1024) if pair else (1, 300, 128, 512)
    a = (torch.randn(seqs, rows, c, generator=g) * 0.5).to(td).cuda()
    w = (torch.randn(n, c, generator=g) / c ** 0.5).to(td).cuda()
    bns = [_bn(n - 3, 18) for _ in range(2)]
    for bn in bns:
        bn.momentum = momentum
    res = []
    native.check(native.lib().vp3d_set_pair_mode(2 if pair else 0), 'pair')
    try:
        for use_fin, bn in zip((False, True), bns):
            outs = []
            for _ in range(2):
                arena = torch.zeros(2 * n + 2, dtype=torch.float64, device='cuda')
                stats = arena[:2 * n].view(2, n)
                z = torch.empty((seqs, rows, n), dtype=td, device='cuda')
                fin, fin_out = (ops.make_bn_fin(bn, seqs * rows, n, arena[2 * n:].view(torch.int32)) if use_fin
                                else (None, None))
                ops.conv_block(dt, a, (seqs, rows, c, c, rows * c), w, 1, 0, c, rows, z, (n, rows * n),
                               stat_sum=stats[0], stat_sqsum=stats[1], fin=fin)
                if not use_fin:
                    fin_out = ops.bn_finalize(stats, seqs * rows, bn, n)
                torch.cuda.synchronize()
                outs.append([t.clone() for t in fin_out])
            res.append((outs, bn))
    finally:
        native.check(native.lib().vp3d_set_pair_mode(1), 'pair')
    (o0, bn0), (o1, bn1) = res
    for step in range(2):
        for t0, t1 in zip(o0[step], o1[step]):
            assert torch.equal(t0, t1)
    assert torch.equal(bn0.running_mean, bn1.running_mean) and torch.equal(bn0.running_var, bn1.running_var)
    assert int(bn0.num_batches_tracked) == int(bn1.num_batches_tracked) == 2
    if momentum is None:     # cumulative average of two identical batches = the batch statistics themselves
        assert (bn1.running_mean[:n - 3] - o1[1][2][:n - 3]).abs().max().item() < 1e-6


def test_dynamic_tile_schedule_gives_the_same_bits():
    """Cluster launch control (one cluster per tile, running clusters steal the tiles of clusters not yet launched)
    against the static persistent schedule: forward with residual, data gradient with MN-major weights + fan-in, and
    both side-input modes -- identical outputs; sizes from one tile to many waves per cluster."""
    dt, td = native.F16, torch.float16
    g = torch.Generator().manual_seed(23)
    lib = native.lib()

    def run(sched, seqs, rows, c, n, kind):
        a = (torch.randn(seqs, rows, c, generator=torch.Generator().manual_seed(seqs * rows + c)) * 0.5).to(td).cuda()
        gen = torch.Generator().manual_seed(n + c)
        out = torch.full((seqs, rows, n), float('nan'), dtype=td, device='cuda')
        kw = {}
        if kind == 'dgrad':
            w = (torch.randn(c, n, generator=gen) / c ** 0.5).to(td).cuda()
            kw['w_mn_major'] = (n, 0)
        else:
            w = (torch.randn(n, c, generator=gen) / c ** 0.5).to(td).cuda()
            kw['shift'] = (torch.randn(n, generator=gen) * 0.2).cuda()
            kw['relu'] = True
        extra = (torch.randn(seqs, rows, n, generator=gen)).to(td).cuda()
        if kind == 'res':
            kw.update(res=extra, res_view=(n, rows * n, 1, 0))
        elif kind == 'side_add':
            kw.update(side=extra, side_view=(n, rows * n, rows, 0), side_mode=1)
        elif kind == 'dgrad':
            kw.update(side=torch.relu(extra), side_view=(n, rows * n, rows, 0), side_mode=2, side_scale=4.0 / 3.0)
        native.check(lib.vp3d_set_pair_mode(2), 'pair')
        native.check(lib.vp3d_set_sched_mode(sched), 'sched')
        try:
            ops.conv_block(dt, a, (seqs, rows, c, c, rows * c), w, 1, 0, c, rows, out, (n, rows * n), **kw)
            torch.cuda.synchronize()
        finally:
            native.check(lib.vp3d_set_pair_mode(1), 'pair')
            native.check(lib.vp3d_set_sched_mode(0), 'sched')
        return out

    for seqs, rows, c, n in [(1, 100, 64, 256), (3, 700, 256, 1024), (1, 40000, 128, 1024), (7, 1111, 192, 512)]:
        for kind in ('res', 'side_add', 'dgrad'):
            ref = run(0, seqs, rows, c, n, kind)
            dyn = run(2, seqs, rows, c, n, kind)
            assert torch.isfinite(dyn.float()).all(), (seqs, rows, c, n, kind)
            assert torch.equal(ref, dyn), (seqs, rows, c, n, kind)


def test_weight_gradient_with_dynamic_items_matches_static():
    """vp3d_wgrad with the dynamic item schedule (one CTA per item, stolen through cluster launch control, finer row
    slices) against the static persistent schedule: the same sum in a different fp32 order."""
    dt, td = native.F16, torch.float16
    g = torch.Generator().manual_seed(29)
    lib = native.lib()
    for seqs, rows, co, ci, taps, step in [(1, 5000, 1024, 1024, 1, 0), (6, 700, 1024, 1024, 3, 9), (1, 300, 128, 1024, 1, 0)]:
        dz = (torch.randn(seqs, rows, co, generator=g) * 0.5).to(td).cuda()
        a = (torch.randn(seqs, rows + step * (taps - 1), ci, generator=g) * 0.5).to(td).cuda()
        outs = []
        for mode in (0, 2):
            native.check(lib.vp3d_set_sched_mode(mode), 'sched')
            try:
                packed = torch.zeros(taps, co, ci, device='cuda')
                ops.wgrad(dt, dz, (seqs, rows, co, rows * co), a, (a.shape[1], ci, ci, a.shape[1] * ci), co, ci, taps,
                          packed, b_tap_row_step=step, block_n=256)
                torch.cuda.synchronize()
            finally:
                native.check(lib.vp3d_set_sched_mode(0), 'sched')
            outs.append(packed)
        ref = torch.stack([torch.einsum('srk,src->kc', dz.double(), a[:, t * step:t * step + rows].double())
                           for t in range(taps)])
        assert rel_err(outs[0], ref) < 5e-6 and rel_err(outs[1], ref) < 5e-6     # fp32 split-K sums, order differs
        assert rel_err(outs[1], outs[0]) < 5e-6
