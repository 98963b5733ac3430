"""Times the Gram GEMM of the fused expand layer (vp3d_wgrad with both operands = the packed input view) for the tile
shapes / split depths on offer, as CUDA-graph replays."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'dynamic-camera-augmented-videopose3d_b200'), ROOT]
import torch
from vp3d_b200 import native, ops

dt = native.F16
n, t_in, c_in, c_in_pad = 1024, 243, 34, 64
x = torch.rand(n * t_in, c_in, device='cuda') * 2 - 1
h = ops.pack_rows(dt, x, c_in_pad, ones_col=c_in)
rows, k = n * 81, 192
xv, av = (1, rows, k, rows * k), (rows, k, k, rows * k)
for block_n, smax in [(64, 0), (64, 148), (256, 0), (256, 148), (256, 32), (256, 16)]:
    gram = torch.zeros(1, 256, 256, device='cuda')
    ops.wgrad(dt, h, xv, h, av, 256, 256, 1, gram, block_n=block_n, dz_cols=k, max_slices=smax)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            ops.wgrad(dt, h, xv, h, av, 256, 256, 1, gram, block_n=block_n, dz_cols=k, max_slices=smax)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    print('block_n %3d max_slices %3d: %.1f us per launch' % (block_n, smax, e0.elapsed_time(e1) / 20 * 1e3))
