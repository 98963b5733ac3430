"""One launch of each HBM-bound training kernel on expand-layer-sized matrices (for ncu)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dynamic-camera-augmented-videopose3d_b200'))
from vp3d_b200 import native, ops
dev = torch.device('cuda')
rows, C = (int(sys.argv[1]) if len(sys.argv) > 1 else 82944), 1024
z = torch.randn(rows, C, device=dev).half()
g = torch.randn(rows, C, device=dev).half()
one = torch.rand(C, device=dev) + 0.5
zero = torch.randn(C, device=dev) * 0.1
d = ops.make_dropout(0.25, 1, 2)
gsb = torch.tensor([1.0, 1.0, 0, 0], device=dev)
st = torch.zeros(2, C, dtype=torch.float64, device=dev)
for _ in range(2):
    ops.col_stats(native.F16, z, st)
    ops.bn_act_fwd(native.F16, z, one, zero, 1, rows, d)
    ops.bn_act_bwd(native.F16, g, z, one, zero, zero, one, rows, C, d, gsb)
torch.cuda.synchronize()
print('done')
