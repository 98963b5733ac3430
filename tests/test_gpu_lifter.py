"""Sibling models' shared pieces on the GPU: the StackedPoseLifter drop-in
(common/models/StackedPoseLifter.py:37-56) against the fixture the real reference produced
(tests/golden/make_golden_lifter.py) and against a mask-pinned fp32 emulation when dropout is on, and the
sliding-window evaluator (CamLSTM.py:33-44) against the reference's window layout."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import load_golden  # noqa: E402
from common.loss import mpjpe  # noqa: E402
from common.models.StackedPoseLifter import StackedPoseLifter  # noqa: E402
from oracle import lifter as ol  # noqa: E402
from vp3d_b200 import lifter as engine  # noqa: E402
from vp3d_b200.evaluation import sliding_window  # noqa: E402

J, F, LAYERS, SIZE, SEED = 17, 3, 3, 256, 21
# parameter gradients against an fp32 oracle: 16-bit operands move pre-activations by ~5e-4 relative, which flips the ReLU
# decision of the few elements that sit at zero; measured 1.5e-3 (last layer) to 3.4e-2 (first layer, four layers of flips
# below it) -- the same effect and bound as the temporal stack's GRAD_TOL_FP32
GRAD_TOL = 8e-2


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _model(dropout, dtype='fp16'):
    m = StackedPoseLifter(J, F, LAYERS, SIZE, dropout=dropout)
    res = m.load_state_dict(ol.init_state(J, F, LAYERS, SIZE, seed=SEED), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    m.operand_dtype = dtype
    return m.cuda()


@pytest.mark.parametrize('dtype,tol', [('fp16', 1e-3), ('bf16', 8e-3), ('tf32', 1e-3)])
def test_eval_forward_matches_the_reference(dtype, tol):
    z = load_golden('lifter.npz')
    m = _model(0.25, dtype).eval()
    a, b = torch.from_numpy(z['a']).cuda(), torch.from_numpy(z['b']).cuda()
    with torch.no_grad():
        y = m(a, b)
        ys = m(a.squeeze(), b.squeeze())                # run.py:521-522
    assert y.shape == (a.shape[0], 1, J, F) and ys.shape == y.shape
    assert rel(y, z['y']) < tol and rel(ys, z['y_squeezed']) < tol
    assert sorted(m.state_dict().keys()) == list(z['state_keys'])


def test_train_step_without_dropout_matches_the_reference_gradients():
    z = load_golden('lifter.npz')
    m = _model(0.0).train()
    a, b, tgt = (torch.from_numpy(z[k]).cuda() for k in ('a', 'b', 'tgt'))
    loss = mpjpe(m(a, b), tgt)
    loss.backward()
    assert abs(loss.item() - float(z['loss'])) < 1e-3 * abs(float(z['loss']))
    for k, p in m.named_parameters():
        if 'grad/' + k in z.files:
            assert rel(p.grad, z['grad/' + k]) < GRAD_TOL, k
        else:
            assert rel(p.grad[::8, ::8], z['grad_sample/' + k]) < GRAD_TOL, k
            assert abs(p.grad.double().norm().item() - float(z['grad_norm/' + k])) < 2e-2 * float(z['grad_norm/' + k]), k


def test_train_step_with_dropout_against_mask_pinned_emulation():
    """Dropout 0.25: the masks are whatever the counter-based generator drew; they are recovered from the saved
    activations (a > 0: kept AND not clipped) and an fp32 torch emulation with exactly those masks provides loss and
    gradients."""
    z = load_golden('lifter.npz')
    m = _model(0.25).train()
    a, b, tgt = (torch.from_numpy(z[k]).cuda() for k in ('a', 'b', 'tgt'))
    engine.debug_keep_saved = True
    try:
        loss = mpjpe(m(a, b), tgt)
        acts = engine.debug_last_acts
    finally:
        engine.debug_keep_saved = False
    loss.backward()
    lins = [l for l in m.mlp_layers if isinstance(l, torch.nn.Linear)]
    masks = [(acts[i + 1][0, :, :lins[i].out_features] > 0) for i in range(len(lins) - 1)]
    # about a quarter of the positive pre-activations were dropped, and two training forwards draw different masks
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    x = torch.cat((a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)), dim=-1)
    keep = 1.0 / 0.75
    h = x
    pos_frac = []
    for i, lin in enumerate(lins[:-1]):
        pre = torch.nn.functional.linear(h, sd['mlp_layers.%d.weight' % (3 * i)], sd['mlp_layers.%d.bias' % (3 * i)])
        pos_frac.append((masks[i].float().sum() / (pre > 0).float().sum().clamp_min(1)).item())
        h = torch.relu(pre) * masks[i] * keep
    y = torch.nn.functional.linear(h, sd['mlp_layers.%d.weight' % (3 * (len(lins) - 1))],
                                   sd['mlp_layers.%d.bias' % (3 * (len(lins) - 1))])
    loss_e = torch.mean(torch.linalg.norm(y.view(-1, 1, J, F) - tgt, dim=3))
    loss_e.backward()
    assert all(0.70 < f < 0.80 for f in pos_frac), pos_frac
    assert abs(loss.item() - loss_e.item()) < 2e-3 * abs(loss_e.item())
    for k, p in m.named_parameters():
        assert rel(p.grad, sd[k].grad) < GRAD_TOL, k
    with torch.no_grad():
        engine.debug_keep_saved = True
        try:
            m(a, b)
        finally:
            engine.debug_keep_saved = False


def test_lifter_trains_with_fused_adam():
    from vp3d_b200.optim import FusedAdam
    z = load_golden('lifter.npz')
    m = _model(0.25).train()
    a, b, tgt = (torch.from_numpy(z[k]).cuda() for k in ('a', 'b', 'tgt'))
    opt = FusedAdam(m.parameters(), lr=1e-3, amsgrad=True)
    losses = []
    for _ in range(80):
        opt.zero_grad()
        loss = mpjpe(m(a, b), tgt)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < 0.8 * np.mean(losses[:3]), losses


def test_sliding_window_hands_the_reference_windows_to_the_model():
    z = load_golden('lifter.npz')
    x2, cam = torch.from_numpy(z['sw_x2']).cuda(), torch.from_numpy(z['sw_cam']).cuda()

    def probe(win_2d, win_cam):      # the checksum model of make_golden_lifter.py
        w = torch.arange(1, win_2d.shape[1] + 1, dtype=win_2d.dtype, device=win_2d.device).view(1, -1, 1, 1)
        s2 = (win_2d * w).sum(dim=(1, 3))
        sc = (win_cam * w).sum(dim=(1, 2, 3)).view(-1, 1)
        return torch.stack([s2, s2 + sc, s2 - sc], dim=-1)

    for w, key in ((9, 'sw_out_w9'), (39, 'sw_out_w40')):
        for chunk in (None, 7):
            out = sliding_window(probe, x2, cam, w, max_windows=chunk)
            assert out.shape == z[key].shape
            assert rel(out, z[key]) < 1e-5
    with pytest.raises(ValueError):
        sliding_window(probe, x2, cam, 41)
