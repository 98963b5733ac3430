"""Small end-to-end pass over every kernel family (for compute-sanitizer --tool memcheck): eval forward, training step
with dropout, fused Adam, feeder, streaming, projection backward, losses. Shapes are tiny but hit ragged tails."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dynamic-camera-augmented-videopose3d_b200'))
from common.camera import project_to_2d, world_to_image
from common.loss import mpjpe, n_mpjpe, weighted_mpjpe
from common.models.TemporalModel import TemporalModel, TemporalModelOptimized1f
from vp3d_b200.feeder import DeviceWindowFeeder
from vp3d_b200.optim import FusedAdam
from vp3d_b200.streaming import CausalStream

torch.manual_seed(0)
fw = [3, 3, 3]
m = TemporalModel(17, 2, 17, fw, causal=True, dropout=0.25, channels=128).cuda().eval()
x = torch.rand(3, 27 + 37, 17, 2).cuda() * 2 - 1
with torch.no_grad():
    y = m(x)
st = CausalStream(m, 3)
for t in range(6):
    st.step(x[:, t])
mt = TemporalModelOptimized1f(17, 2, 17, fw, dropout=0.25, channels=128).cuda().train()
opt = FusedAdam(mt.parameters(), lr=1e-3, amsgrad=True)
rng = np.random.default_rng(0)
X = [rng.normal(0, 0.3, (n, 17, 3)).astype(np.float32) + np.array([0, 0, 4], np.float32) for n in (40, 55)]
Q = [np.tile(np.array([1, 0, 0, 0], np.float32), (n, 1)) for n in (40, 55)]
T = [np.zeros((n, 3), np.float32) for n in (40, 55)]
cam = np.tile(np.array([2.29, 2.2876, 0.0251, 0.0289, -0.2071, 0.2478, -0.0031, -0.00098, -0.0014], np.float32), (2, 1))
fd = DeviceWindowFeeder(X, Q, T, cam, batch_size=37, pad=13, want_cameras=True)
for i, (cams, b3d, b2d) in enumerate(fd.next_epoch()):
    opt.zero_grad()
    pred = mt(b2d)
    loss = mpjpe(pred, b3d) + weighted_mpjpe(pred, b3d, torch.rand(b3d.shape[0], 1, 1, device='cuda')) + \
        mpjpe(project_to_2d(pred + torch.tensor([0.0, 0.0, 5.0], device='cuda'), torch.from_numpy(cam[:1]).cuda().repeat(b3d.shape[0], 1)), b2d[:, 13:14])
    loss.backward()
    opt.step()
    if i >= 1:
        break
mf = TemporalModel(17, 2, 17, fw, dropout=0.1, channels=128).cuda().train()
xx = torch.rand(2, 27 + 5, 17, 2).cuda()
out = mf(xx)
mpjpe(out, torch.zeros_like(out)).backward()
print('n_mpjpe', n_mpjpe(out.detach(), torch.ones_like(out)).item())
torch.cuda.synchronize()
print('sanity pass done, loss', loss.item())
