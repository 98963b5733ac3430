"""Micro-benchmarks of the HBM-bound kernels (projection, loss, BN/activation passes) with rotating buffers larger than
L2, CUDA-event timed: achieved GB/s against MEASURED_PEAKS.json. Usage: python tools/kernel_bench.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dynamic-camera-augmented-videopose3d_b200'))
from common.camera import world_to_image  # noqa: E402
from common.loss import mpjpe  # noqa: E402
from vp3d_b200 import native, ops  # noqa: E402

peak = 6565.5
p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
if os.path.exists(p):
    peak = json.load(open(p))['hbm_gbs']
dev = torch.device('cuda')


def timeit(fn, n_sets, reps=5):
    for i in range(n_sets):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for i in range(n_sets):
            fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * n_sets)


res = {}
B, T, J = 4096, 243, 17
sets = []
for i in range(2):
    X = torch.randn(B, T, J, 3, device=dev) * 0.3
    X[..., 2] += 4
    q = torch.randn(B, T, 4, device=dev)
    q = q / q.norm(dim=-1, keepdim=True)
    t = torch.randn(B, T, 3, device=dev) * 0.1
    cam = torch.tensor([2.29, 2.2876, 0.0251, 0.0289, -0.2071, 0.2478, -0.0031, -0.00098, -0.0014], device=dev).repeat(B, 1)
    sets.append((X, q, t, cam))
ms = timeit(lambda i: world_to_image(*sets[i], return_camera_space=False), 2)
gb = 404 * B * T / 1e9
res['project world_to_image (B=4096, 243 f, J=17)'] = (ms, gb / (ms * 1e-3), gb / (ms * 1e-3) / peak)

preds = [torch.randn(256, 4096, 17, 3, device=dev) for _ in range(2)]
tg = [torch.randn(256, 4096, 17, 3, device=dev) for _ in range(2)]
ms = timeit(lambda i: mpjpe(preds[i], tg[i]), 2)
gb = 24 * 256 * 4096 * 17 / 1e9
res['mpjpe fwd (256 x 4096 x 17 joints)'] = (ms, gb / (ms * 1e-3), gb / (ms * 1e-3) / peak)

rows, C = 82944, 1024
zs = [torch.randn(rows, C, device=dev).half() for _ in range(3)]
gs = [torch.randn(rows, C, device=dev).half() for _ in range(3)]
one = torch.rand(C, device=dev) + 0.5
zero = torch.randn(C, device=dev) * 0.1
for pdrop in (0.0, 0.25):
    d = ops.make_dropout(pdrop, 1, 2)
    ms = timeit(lambda i: ops.bn_act_fwd(native.F16, zs[i], one, zero, 1, rows, d), 3)
    gb = 4 * rows * C / 1e9
    res['bn_act_fwd p=%.2f (82944 x 1024)' % pdrop] = (ms, gb / (ms * 1e-3), gb / (ms * 1e-3) / peak)
    gsb = torch.tensor([1.0, 1.0, 0, 0], device=dev)
    ms = timeit(lambda i: ops.bn_act_bwd(native.F16, gs[i], zs[i], one, zero, zero, one, rows, C, d, gsb), 3)
    gb = 10 * rows * C / 1e9
    res['bn_act_bwd reduce+apply p=%.2f' % pdrop] = (ms, gb / (ms * 1e-3), gb / (ms * 1e-3) / peak)
st = torch.zeros(2, C, dtype=torch.float64, device=dev)
ms = timeit(lambda i: ops.col_stats(native.F16, zs[i], st), 3)
gb = 2 * rows * C / 1e9
res['col_stats'] = (ms, gb / (ms * 1e-3), gb / (ms * 1e-3) / peak)
for k, (ms, gbs, frac) in res.items():
    print('%-48s %8.3f ms  %8.1f GB/s  %.3f of measured HBM peak (%.1f)' % (k, ms, gbs, frac, peak))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump({k: dict(ms=v[0], gbs=v[1], frac=v[2]) for k, v in res.items()}, open(os.path.join(ROOT, 'gpurun_out', 'kernel_bench.json'), 'w'), indent=1)
