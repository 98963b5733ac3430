"""K1 unit parity: one vp3d_conv_block_fwd launch against an exact fp64 contraction of the same rounded operands.
Calls go through the C ABI (ctypes) exactly as the product path does."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from vp3d_b200 import native, ops  # noqa: E402

DT = {'fp16': native.F16, 'bf16': native.BF16, 'tf32': native.TF32}
TOL = {'fp16': 2e-3, 'bf16': 2e-2, 'tf32': 1e-4}   # output rounding of the operand type dominates (|y| ~ 1)


def _to_tf32(x):
    """Round fp32 to the TF32 grid (10 explicit mantissa bits), as the pack kernels / epilogue do."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1fff).view(torch.float32)


def _ref_conv(a, w, taps, step, rows_out, row_off=0):
    """a [s][rows][c] , w [n][taps*c] -> [s][rows_out][n] in fp64, rows outside the input read as zero."""
    s, rows, c = a.shape
    a64, w64 = a.double().cpu(), w.double().cpu()
    out = torch.zeros(s, rows_out, w.shape[0], dtype=torch.float64)
    for k in range(taps):
        lo = row_off + k * step
        src = torch.zeros(s, rows_out, c, dtype=torch.float64)
        r0, r1 = max(0, -lo), min(rows_out, rows - lo)
        if r1 > r0:
            src[:, r0:r1] = a64[:, r0 + lo:r1 + lo]
        out += src @ w64[:, k * c:(k + 1) * c].T
    return out


@pytest.mark.parametrize('dtype', ['fp16', 'bf16', 'tf32'])
@pytest.mark.parametrize('seqs,rows,c,n,taps,step', [
    (1, 128, 64, 256, 1, 0),        # single tile, single K block
    (1, 300, 256, 256, 1, 0),       # ragged M tail, several K blocks
    (3, 200, 128, 512, 3, 5),       # dilated 3-tap, per-sequence tiling, 2 N tiles
    (2, 700, 1024, 1024, 3, 81),    # the block-4 geometry of the 243-frame model
])
def test_conv_block_matches_fp64(dtype, seqs, rows, c, n, taps, step):
    dt = DT[dtype]
    td = ops.torch_dtype(dt)
    g = torch.Generator(device='cpu').manual_seed(seqs * 1000 + rows)
    a = (torch.randn(seqs, rows, c, generator=g) * 0.5).to(td)
    w = (torch.randn(n, taps * c, generator=g) / (taps * c) ** 0.5).to(td)
    if dt == native.TF32:   # the product path always hands the tf32 MMA pre-rounded operands
        a, w = _to_tf32(a), _to_tf32(w)
    a, w = a.cuda(), w.cuda()
    rows_out = rows - step * (taps - 1)
    scale = (torch.rand(n, generator=g) + 0.5).cuda()
    shift = (torch.randn(n, generator=g) * 0.1).cuda()
    res = (torch.randn(seqs, rows + 2, n, generator=g) * 0.5).to(td).cuda()
    out_f32 = dt == native.TF32
    out = torch.full((seqs, rows_out, n), float('nan'), dtype=torch.float32 if out_f32 else td, device='cuda')
    ops.conv_block(dt, a, (seqs, rows, c, c, rows * c), w, taps, step, c, rows_out, out, (n, rows_out * n),
                   scale=scale, shift=shift, relu=True, res=res, res_view=(n, (rows + 2) * n, 1, 2), out_f32=out_f32,
                   n_valid=n)
    torch.cuda.synchronize()
    a_r = a.float() if dt != native.TF32 else a
    ref = _ref_conv(a_r, w.float(), taps, step, rows_out)
    ref = torch.relu(ref * scale.double().cpu() + shift.double().cpu()) + res.double().cpu()[:, 2:2 + rows_out]
    got = out.double().cpu()
    assert torch.isfinite(got).all(), 'kernel left unwritten / non-finite outputs'
    err = (got - ref).abs().max().item()
    assert err < TOL[dtype], 'max abs err %.3e' % err


def test_narrow_fp32_output_with_bias_and_column_mask():
    """The shrink layer: N = 51 real columns in a 64-wide tile, fp32 output with an odd row stride."""
    dt = native.F16
    g = torch.Generator(device='cpu').manual_seed(5)
    rows, c, n, n_pad = 333, 1024, 51, 64
    a = (torch.randn(1, rows, c, generator=g) * 0.5).half().cuda()
    w = torch.zeros(n_pad, c)
    w[:n] = torch.randn(n, c, generator=g) / c ** 0.5
    w = w.half().cuda()
    bias = torch.zeros(n_pad)
    bias[:n] = torch.randn(n, generator=g)
    out = torch.full((1, rows, n), float('nan'), device='cuda')
    ops.conv_block(dt, a, (1, rows, c, c, rows * c), w, 1, 0, c, rows, out, (n, rows * n), block_n=64,
                   scale=torch.ones(n_pad, device='cuda'), shift=bias.cuda(), relu=False, out_f32=True, n_valid=n)
    torch.cuda.synchronize()
    ref = a.double().cpu()[0] @ w.double().cpu()[:n].T + bias[:n].double()
    assert torch.isfinite(out).all()
    assert (out.double().cpu()[0] - ref).abs().max().item() < 1e-3


def test_strided_view_is_a_plain_gemm_with_strided_residual():
    """stride == width convolution of the 1f model as taps=1 on the [t_out][3c] view; residual row = 3*row + 1."""
    dt = native.F16
    g = torch.Generator(device='cpu').manual_seed(9)
    n_seq, t_in, c, n = 5, 81, 256, 256
    t_out = t_in // 3
    x = (torch.randn(n_seq, t_in, c, generator=g) * 0.5).half().cuda()
    w = (torch.randn(n, 3 * c, generator=g) / (3 * c) ** 0.5).half().cuda()
    out = torch.full((n_seq * t_out, n), float('nan'), dtype=torch.float16, device='cuda')
    ops.conv_block(dt, x, (1, n_seq * t_out, 3 * c, 3 * c, n_seq * t_in * c), w, 1, 0, 3 * c, n_seq * t_out, out,
                   (n, n_seq * t_out * n), res=x, res_view=(c, n_seq * t_in * c, 3, 1))
    torch.cuda.synchronize()
    xr = x.double().cpu().reshape(n_seq * t_out, 3 * c)
    ref = xr @ w.double().cpu().T + x.double().cpu().reshape(n_seq * t_out, 3, c)[:, 1, :n]
    assert (out.double().cpu() - ref).abs().max().item() < 4e-3


def test_negative_row_offset_reads_zeros():
    """a_row_off < 0 (used by data-gradient GEMMs of dilated layers): TMA zero-fills rows before the sequence."""
    dt = native.F16
    g = torch.Generator(device='cpu').manual_seed(10)
    rows, c, n = 200, 64, 256
    a = (torch.randn(2, rows, c, generator=g)).half().cuda()
    w = (torch.randn(n, 2 * c, generator=g) / (2 * c) ** 0.5).half().cuda()
    out = torch.full((2, rows, n), float('nan'), dtype=torch.float16, device='cuda')
    ops.conv_block(dt, a, (2, rows, c, c, rows * c), w, 2, 7, c, rows, out, (n, rows * n), a_row_off=-7)
    torch.cuda.synchronize()
    ref = _ref_conv(a.float(), w.float(), 2, 7, rows, row_off=-7)
    assert (out.double().cpu() - ref).abs().max().item() < 4e-3


def test_argument_errors_are_reported_not_crashed():
    a = torch.zeros(1, 128, 64, dtype=torch.float16, device='cuda')
    w = torch.zeros(256, 64, dtype=torch.float16, device='cuda')
    out = torch.zeros(1, 128, 256, dtype=torch.float16, device='cuda')
    with pytest.raises(AssertionError):
        ops.conv_block(native.F16, a, (1, 128, 64, 64, 128 * 64), w, 1, 0, 48, 128, out, (256, 128 * 256))  # bad K
    with pytest.raises(AssertionError):
        ops.conv_block(native.F16, a, (1, 128, 64, 64, 128 * 64), w, 1, 0, 64, 128, out, (256, 128 * 256), block_n=96)


@pytest.fixture
def pair_mode():
    """Selects the K1 kernel explicitly (vp3d_set_pair_mode) and restores the default afterwards."""
    def set_mode(mode):
        native.check(native.lib().vp3d_set_pair_mode(mode), 'set_pair_mode')
    yield set_mode
    set_mode(1)


@pytest.mark.parametrize('dtype', ['fp16', 'bf16'])
@pytest.mark.parametrize('seqs,rows,c,n,taps,step', [
    (1, 128, 64, 256, 1, 0),        # one pair tile whose second CTA has no valid row at all
    (1, 300, 256, 256, 1, 0),       # odd number of row tiles: the last pair's partner tile is out of range
    (3, 200, 128, 512, 3, 5),       # dilated 3-tap, per-sequence pairs, 2 column tiles
    (2, 700, 1024, 1024, 3, 81),    # the block-4 geometry of the 243-frame model (4 column tiles, affine table)
    (1, 1500, 128, 1024, 1, 0),     # more pair tiles than clusters: several tiles per pair, column tile changes
])
def test_pair_kernel_matches_fp64_and_single_cta_kernel(pair_mode, dtype, seqs, rows, c, n, taps, step):
    """conv_gemm_pair_kernel (tcgen05.mma.cta_group::2) forced on small shapes: against the fp64 contraction and,
    bit for bit, against the single-CTA kernel (same MMA shapes per row, same epilogue arithmetic)."""
    dt = DT[dtype]
    td = ops.torch_dtype(dt)
    g = torch.Generator(device='cpu').manual_seed(seqs * 1000 + rows + 7)
    a = (torch.randn(seqs, rows, c, generator=g) * 0.5).to(td).cuda()
    w = (torch.randn(n, taps * c, generator=g) / (taps * c) ** 0.5).to(td).cuda()
    rows_out = rows - step * (taps - 1)
    scale = (torch.rand(n, generator=g) + 0.5).cuda()
    shift = (torch.randn(n, generator=g) * 0.1).cuda()
    res = (torch.randn(seqs, rows + 2, n, generator=g) * 0.5).to(td).cuda()
    outs = []
    for mode in (0, 2):
        pair_mode(mode)
        out = torch.full((seqs, rows_out, n), float('nan'), dtype=td, device='cuda')
        ops.conv_block(dt, a, (seqs, rows, c, c, rows * c), w, taps, step, c, rows_out, out, (n, rows_out * n),
                       scale=scale, shift=shift, relu=True, res=res, res_view=(n, (rows + 2) * n, 1, 2))
        torch.cuda.synchronize()
        outs.append(out)
    ref = _ref_conv(a.float(), w.float(), taps, step, rows_out)
    ref = torch.relu(ref * scale.double().cpu() + shift.double().cpu()) + res.double().cpu()[:, 2:2 + rows_out]
    got = outs[1].double().cpu()
    assert torch.isfinite(got).all(), 'pair kernel left unwritten / non-finite outputs'
    assert (got - ref).abs().max().item() < TOL[dtype]
    assert torch.equal(outs[0], outs[1])


def test_sm_limit_changes_the_grid_not_the_result(pair_mode):
    """vp3d_set_sm_limit (SMs left to NCCL under data-parallel training) only re-sizes the persistent grids."""
    g = torch.Generator(device='cpu').manual_seed(77)
    rows, c, n = 5000, 256, 512
    a = (torch.randn(1, rows, c, generator=g) * 0.5).half().cuda()
    w = (torch.randn(n, c, generator=g) / c ** 0.5).half().cuda()
    outs = []
    try:
        for limit, mode in ((0, 1), (100, 1), (37, 2)):
            native.check(native.lib().vp3d_set_sm_limit(limit), 'set_sm_limit')
            pair_mode(mode)
            out = torch.empty((1, rows, n), dtype=torch.float16, device='cuda')
            ops.conv_block(native.F16, a, (1, rows, c, c, rows * c), w, 1, 0, c, rows, out, (n, rows * n), relu=True)
            torch.cuda.synchronize()
            outs.append(out)
    finally:
        native.check(native.lib().vp3d_set_sm_limit(0), 'set_sm_limit')
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
