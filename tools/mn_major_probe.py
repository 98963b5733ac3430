"""Does an MN-major weight operand cost tensor-core rate? Same GEMM (M x K x N) through K1 with K-major weights [N][K]
and with MN-major weights [K][N]; CUDA-event timing over rotating buffers."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dynamic-camera-augmented-videopose3d_b200'))
from vp3d_b200 import native, ops
dev = torch.device('cuda')
dt = native.F16
for (M, K, N) in ((27648, 1024, 3072), (27648, 3072, 1024), (82944, 1024, 1024)):
    a = [torch.randn(M, K, device=dev).half() for _ in range(2)]
    wk = (torch.randn(N, K, device=dev) / K ** 0.5).half()
    wm = wk.t().contiguous()
    out = [torch.empty(M, N, dtype=torch.float16, device=dev) for _ in range(2)]
    def run_k(i):
        ops.conv_block(dt, a[i], (1, M, K, K, M * K), wk, 1, 0, K, M, out[i], (N, M * N))
    def run_m(i):
        ops.conv_block(dt, a[i], (1, M, K, K, M * K), wm, 1, 0, K, M, out[i], (N, M * N), w_mn_major=(N, 0))
    res = {}
    for name, fn in (('K-major', run_k), ('MN-major', run_m)):
        for i in range(4):
            fn(i % 2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            fn(i % 2)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        res[name] = (ms, 2.0 * M * K * N / ms / 1e9, out[0].float().clone())
    err = (res['K-major'][2] - res['MN-major'][2]).abs().max().item()
    print('M=%d K=%d N=%d  K-major %.3f ms %.0f TFLOP/s | MN-major %.3f ms %.0f TFLOP/s | max diff %.2e'
          % (M, K, N, res['K-major'][0], res['K-major'][1], res['MN-major'][0], res['MN-major'][1], err))
