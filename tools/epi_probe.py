"""Epilogue ablation probe: a 1x1-layer-shaped GEMM (M = 82944, K = 1024, N = 1024) with shift + ReLU + residual."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dynamic-camera-augmented-videopose3d_b200'))
from vp3d_b200 import native, ops
dev = torch.device('cuda')
dt = native.F16
M, K, N = 82944 * 3, 1024, 1024
a = [torch.randn(M, K, device=dev).half() for _ in range(2)]
res = [torch.randn(M, N, device=dev).half() for _ in range(2)]
w = (torch.randn(N, K, device=dev) / K ** 0.5).half()
out = [torch.empty(M, N, dtype=torch.float16, device=dev) for _ in range(2)]
shift = torch.randn(N, device=dev) * 0.1
def run(i, with_res):
    kw = dict(res=res[i], res_view=(N, M * N, 1, 0)) if with_res else {}
    ops.conv_block(dt, a[i], (1, M, K, K, M * K), w, 1, 0, K, M, out[i], (N, M * N), scale=None, shift=shift, relu=True, **kw)
for with_res in (True, False):
    for i in range(4):
        run(i % 2, with_res)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        run(i % 2, with_res)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print('%s residual=%s: %.3f ms  %.0f TFLOP/s' % (os.path.basename(os.environ.get('VP3D_LIB_PATH', 'default')), with_res, ms, 2.0 * M * K * N / ms / 1e9))
