"""First-light diagnostics for the tcgen05 conv GEMM: structured operands whose product reveals layout mistakes
(swizzle, descriptor strides, TMEM lane/column mapping). Writes gpurun_out/diag.npz for offline inspection."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dynamic-camera-augmented-videopose3d_b200'))
from vp3d_b200 import native, ops  # noqa: E402

os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
print('device', torch.cuda.get_device_name(0), native.device_info())
dump = {}


def run(name, dt, a, w, taps=1, step=0, block_n=256, out_f32=False):
    seqs, rows, c = a.shape
    n = w.shape[0]
    rows_out = rows - step * (taps - 1)
    td = torch.float32 if out_f32 else ops.torch_dtype(dt)
    out = torch.full((seqs, rows_out, n), float('nan'), dtype=td, device='cuda')
    ops.conv_block(dt, a, (seqs, rows, c, c, rows * c), w, taps, step, c, rows_out, out, (n, rows_out * n),
                   block_n=block_n, out_f32=out_f32, n_valid=n)
    torch.cuda.synchronize()
    a64, w64 = a.double().cpu(), w.double().cpu()
    ref = torch.zeros(seqs, rows_out, n, dtype=torch.float64)
    for k in range(taps):
        ref += a64[:, k * step:k * step + rows_out] @ w64[:, k * c:(k + 1) * c].T
    got = out.double().cpu()
    err = (got - ref).abs()
    bad = (err > 1e-2 * (1 + ref.abs())) | ~torch.isfinite(got)
    print('%-28s max err %.3e  bad %d / %d  nan %d' % (name, err[torch.isfinite(err)].max().item() if torch.isfinite(err).any() else float('nan'),
                                                       int(bad.sum()), bad.numel(), int((~torch.isfinite(got)).sum())))
    if bad.any():
        idx = bad.nonzero()[:8]
        for i in idx:
            s, r, c_ = [int(v) for v in i]
            print('   [%d,%d,%d] got %.4f ref %.4f' % (s, r, c_, got[s, r, c_], ref[s, r, c_]))
        dump[name + '/got'] = got.numpy().astype(np.float32)
        dump[name + '/ref'] = ref.numpy().astype(np.float32)
    return not bad.any()


ok = True
# 1. one-hot A (row r selects k = r % 64), W[n, k] = n + k / 64  -> out[r, n] = n + (r % 64) / 64
a = torch.zeros(1, 128, 64)
a[0, torch.arange(128), torch.arange(128) % 64] = 1
w = torch.arange(256).float()[:, None] + torch.arange(64).float()[None, :] / 64
ok &= run('onehot fp16 128x64x256', native.F16, a.half().cuda(), w.half().cuda())
ok &= run('onehot bf16 128x64x256', native.BF16, a.bfloat16().cuda(), w.bfloat16().cuda())
a32 = torch.zeros(1, 128, 32)
a32[0, torch.arange(128), torch.arange(128) % 32] = 1
w32 = torch.arange(256).float()[:, None] + torch.arange(32).float()[None, :] / 32
ok &= run('onehot tf32 128x32x256', native.TF32, a32.cuda(), w32.cuda(), out_f32=True)
g = torch.Generator().manual_seed(0)
for dt, name in ((native.F16, 'fp16'), (native.BF16, 'bf16'), (native.TF32, 'tf32')):
    td = ops.torch_dtype(dt)
    f32 = dt == native.TF32
    A = lambda *s: (torch.randn(*s, generator=g) * 0.5).to(td).cuda()
    ok &= run('rand %s 1x128x64 n256' % name, dt, A(1, 128, 64), A(256, 64), out_f32=f32)
    ok &= run('rand %s 1x128x512 n256' % name, dt, A(1, 128, 512), A(256, 512), out_f32=f32)
    ok &= run('rand %s 1x1000x256 n1024' % name, dt, A(1, 1000, 256), A(1024, 256), out_f32=f32)
    ok &= run('rand %s 3x300x128 n512 3tap d7' % name, dt, A(3, 300, 128), A(512, 384), taps=3, step=7, out_f32=f32)
    ok &= run('rand %s 1x500x1024 n64 narrow' % name, dt, A(1, 500, 1024), A(64, 1024), block_n=64, out_f32=f32)
    ok &= run('rand %s 40x4000x128 many tiles' % name, dt, A(40, 4000, 128), A(256, 128), out_f32=f32)
if dump:
    np.savez_compressed(os.path.join(ROOT, 'gpurun_out', 'diag.npz'), **dump)
print('ALL OK' if ok else 'FAILURES')
sys.exit(0 if ok else 1)
