"""Diagnostic: after the 30-step run of tests/test_gpu_training.py::test_training_run_follows_the_oracle_loss_trajectory,
per-tensor relative differences between the CUDA-trained module and the oracle-trained state (parameters, running
statistics) and of the eval outputs with the statistics swapped."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dynamic-camera-augmented-videopose3d_b200')); sys.path.insert(0, ROOT)
import torch
from common.loss import mpjpe
from common.models.TemporalModel import TemporalModelOptimized1f
from oracle import temporal_model as otm
from vp3d_b200.optim import FusedAdam

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

fw = [3, 3, 3]
steps, batch, lr = int(os.environ.get('STEPS', 30)), 128, 1e-3
sd0 = otm.init_state(17, 2, 17, fw, channels=1024, seed=41)
g = torch.Generator().manual_seed(42)
proj = torch.randn(34, 51, generator=g) * 0.3
xs = torch.rand(steps, batch, 27, 17, 2, generator=g) * 2 - 1
tg = (xs[:, :, 13].reshape(steps, batch, 34) @ proj).reshape(steps, batch, 1, 17, 3)
sd = {k: v.clone() for k, v in sd0.items()}
names = [k for k, v in sd.items() if v.dtype.is_floating_point and 'running_' not in k]
plist = [sd[k] for k in names]
opt_o = torch.optim.Adam(plist, lr=lr, amsgrad=True)
m = TemporalModelOptimized1f(17, 2, 17, fw, dropout=0.0, channels=1024)
m.load_state_dict(sd0)
m = m.cuda().train()
opt = FusedAdam(m.parameters(), lr=lr, amsgrad=True) if os.environ.get('OPT', 'fused') == 'fused' else torch.optim.Adam(m.parameters(), lr=lr, amsgrad=True)
for i in range(steps):
    loss, _, grads, new_stats = otm.train_step_grads(sd, xs[i], tg[i], fw, strided=True)
    for k, p in zip(names, plist):
        p.grad = grads[k]
    opt_o.step(); sd.update(new_stats)
    opt.zero_grad()
    lg = mpjpe(m(xs[i].cuda()), tg[i].cuda()); lg.backward(); opt.step()
    if i % 5 == 0 or i == steps - 1:
        msd = m.state_dict()
        worst = max(((rel(msd[k], sd[k]), k) for k in sd if sd[k].dtype.is_floating_point), key=lambda t: t[0])
        print(i, 'loss', round(lg.item(), 5), round(float(loss), 5), 'worst tensor', worst)
msd = m.state_dict()
for k in sd:
    if sd[k].dtype.is_floating_point:
        print('%-36s %.3e' % (k, rel(msd[k], sd[k])))
    else:
        print(k, int(msd[k]), int(sd[k]))
xe = torch.rand(64, 27, 17, 2, generator=g) * 2 - 1
m.eval()
with torch.no_grad():
    ye = m(xe.cuda()).cpu()
    ref = otm.forward(sd, xe, fw, strided=True)
    cross = otm.forward({k: v.cpu() for k, v in msd.items()}, xe, fw, strided=True)   # oracle eval of the CUDA-trained state
print('eval: cuda vs oracle', rel(ye, ref), ' oracle(cuda state) vs oracle', rel(cross, ref), ' cuda vs oracle(cuda state)', rel(ye, cross))
