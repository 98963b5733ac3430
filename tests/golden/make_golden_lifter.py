"""Golden fixture for the sibling models' shared pieces, produced by importing the REAL reference from
/root/reference: StackedPoseLifter eval forward + the gradients of one mpjpe step (dropout 0), and the window layout of
CamLSTMBase.sliding_window. Run once in the build container:  python tests/golden/make_golden_lifter.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, '/root/reference')
sys.path.insert(1, ROOT)

from common.models.StackedPoseLifter import StackedPoseLifter  # noqa: E402  (reference)
from common.models.CamLSTM import CamLSTMBase  # noqa: E402  (reference)
from common.loss import mpjpe  # noqa: E402  (reference)

from oracle import lifter as ol  # noqa: E402

torch.set_num_threads(8)
J, F, LAYERS, SIZE, SEED = 17, 3, 3, 256, 21       # common/arguments.py:64-65 defaults


class _Probe(CamLSTMBase):
    """Stand-in for a camera-aware model: returns a checksum of each window, so the fixture pins WHICH frames
    sliding_window hands to the model and in what order."""

    def forward(self, win_2d, win_cam):
        w = torch.arange(1, win_2d.shape[1] + 1, dtype=win_2d.dtype).view(1, -1, 1, 1)
        s2 = (win_2d * w).sum(dim=(1, 3))                                   # (n, J)
        sc = (win_cam * w).sum(dim=(1, 2, 3)).view(-1, 1)                   # (n, 1)
        return torch.stack([s2, s2 + sc, s2 - sc], dim=-1)                  # (n, J, 3)


def main():
    out = {}
    sd = ol.init_state(J, F, LAYERS, SIZE, seed=SEED)
    m = StackedPoseLifter(J, F, LAYERS, SIZE, dropout=0.0)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    out['state_keys'] = np.array(sorted(m.state_dict().keys()))
    g = torch.Generator().manual_seed(22)
    a = torch.randn(160, 1, J, F, generator=g) * 0.4
    b = a + torch.randn(160, 1, J, F, generator=g) * 0.05
    tgt = a + torch.randn(160, 1, J, F, generator=g) * 0.02
    out['a'], out['b'], out['tgt'] = a.numpy(), b.numpy(), tgt.numpy()
    m.eval()
    with torch.no_grad():
        out['y'] = m(a, b).numpy()
        out['y_squeezed'] = m(a.squeeze(), b.squeeze()).numpy()            # run.py:521-522 passes (T, J, F)
    m.train()
    loss = mpjpe(m(a, b), tgt)
    loss.backward()
    out['loss'] = loss.detach().numpy()
    for k, p in m.named_parameters():
        # whole tensors for the first / last layer and every bias; a strided sample + the norm for the 256 x 256 ones
        if p.dim() == 1 or k.startswith('mlp_layers.0.') or p.shape[0] == J * F:
            out['grad/' + k] = p.grad.numpy()
        else:
            out['grad_sample/' + k] = p.grad.numpy()[::8, ::8].copy()
            out['grad_norm/' + k] = np.float64(p.grad.double().norm())
    # sliding window (CamLSTM.py:33-44)
    probe = _Probe(J, 2, J, 3, 8, 1, [8])
    x2 = torch.randn(1, 40, J, 2, generator=g)
    cam = torch.randn(1, 40, 3, 4, generator=g)
    out['sw_x2'], out['sw_cam'] = x2.numpy(), cam.numpy()
    out['sw_out_w9'] = probe.sliding_window(x2, cam, 9).numpy()
    out['sw_out_w40'] = probe.sliding_window(x2[:, :40], cam[:, :40], 39).numpy()
    np.savez_compressed(os.path.join(HERE, 'lifter.npz'), **out)
    print('lifter.npz', os.path.getsize(os.path.join(HERE, 'lifter.npz')))


if __name__ == '__main__':
    main()
