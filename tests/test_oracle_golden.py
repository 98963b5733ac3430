"""The CPU oracle (oracle/) against golden outputs of the real reference (tests/golden/make_golden.py).
This pins the oracle; the GPU tests then compare the CUDA path with the oracle."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import load_golden, state_from_npz
from oracle import camera as ocam
from oracle import loss as oloss
from oracle import temporal_model as otm


def _checksum(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def test_small_models_eval():
    z = load_golden('temporal_small.npz')
    sd = state_from_npz(z, 'sd/')
    x = torch.from_numpy(z['x'])
    fw = [3, 3, 3]
    with torch.no_grad():
        np.testing.assert_allclose(otm.forward(sd, x, fw).numpy(), z['y_full'], rtol=0, atol=1e-6)
        np.testing.assert_allclose(otm.forward(sd, x, fw, causal=True).numpy(), z['y_causal'], rtol=0, atol=1e-6)
        xw = x[:, :27].contiguous()
        np.testing.assert_allclose(otm.forward(sd, xw, fw, strided=True).numpy(), z['y_1f'], rtol=0, atol=1e-6)
        np.testing.assert_allclose(otm.forward(sd, xw, fw, strided=True, causal=True).numpy(), z['y_1f_causal'],
                                   rtol=0, atol=1e-6)
        sdd = state_from_npz(z, 'sdd/')
        np.testing.assert_allclose(otm.forward(sdd, x, fw, dense=True).numpy(), z['y_dense'], rtol=0, atol=1e-6)


@pytest.mark.parametrize('name,strided', [('1f', True), ('full', False)])
def test_small_models_train_step(name, strided):
    z = load_golden('temporal_small.npz')
    sd = state_from_npz(z, 'sd/')
    x = torch.from_numpy(z['x'])
    if strided:
        x = x[:, :27].contiguous()
    tgt = torch.from_numpy(z['train_%s/target' % name])
    loss, pred, grads, stats = otm.train_step_grads(sd, x, tgt, [3, 3, 3], strided=strided)
    np.testing.assert_allclose(pred.numpy(), z['train_%s/pred' % name], atol=2e-6)
    np.testing.assert_allclose(loss.numpy(), z['train_%s/loss' % name], atol=1e-6)
    for k, g in grads.items():
        np.testing.assert_allclose(g.numpy(), z['train_%s/grad/%s' % (name, k)], atol=1e-6, err_msg=k)
    for k, v in stats.items():
        np.testing.assert_allclose(v.numpy(), z['train_%s/buf/%s' % (name, k)], atol=1e-6, err_msg=k)


def test_api_helpers():
    z = load_golden('temporal_small.npz')
    assert otm.receptive_field(otm.make_plan([3, 3, 3])) == int(z['api/receptive_field']) == 27
    assert otm.receptive_field(otm.make_plan([3, 3, 3, 3, 3])) == 243


@pytest.mark.parametrize('fname', ['temporal_27f_1024.npz', 'temporal_243f_1024.npz',
                                   'temporal_243f_1024_causal.npz', 'temporal_243f_j31.npz'])
def test_seeded_1024_channel_models(fname):
    z = load_golden(fname)
    fw = [int(v) for v in z['filter_widths']]
    sd = otm.init_state(int(z['j_in']), 2, int(z['j_out']), fw, channels=1024, seed=int(z['seed']))
    assert _checksum(sd) == str(z['checksum']), 'seeded state_dict drifted from the one the golden was made with'
    x = torch.from_numpy(z['x'])
    causal = bool(z['causal'])
    with torch.no_grad():
        y = otm.forward(sd, x, fw, causal=causal).numpy()
        rf = otm.receptive_field(otm.make_plan(fw))
        y1f = otm.forward(sd, x[:, :rf].contiguous(), fw, causal=causal, strided=True).numpy()
    np.testing.assert_allclose(y, z['y'], atol=2e-6)
    np.testing.assert_allclose(y1f, z['y_1f'], atol=2e-6)
    # reference property (TemporalModel.py:147-149): the 1f model equals the full model on an RF-long window
    np.testing.assert_allclose(y1f, z['y'][:, :1], atol=2e-6)


def test_camera_oracle():
    z = load_golden('camera.npz')
    X, q, t = z['X'], z['q'], z['t']
    np.testing.assert_allclose(ocam.world_to_camera(X, q, t), z['w2c'], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(ocam.camera_to_world(X, q, t), z['c2w'], rtol=2e-6, atol=1e-6)
    qf, tf = z['qf'], z['tf']
    qb = np.ascontiguousarray(np.broadcast_to(qf[:, None, :], (*X.shape[:-1], 4)))
    np.testing.assert_allclose(ocam.qrot(qb, X), z['qrot_f'], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(ocam.world_to_camera(X, qf, tf), z['w2c_f'], rtol=2e-6, atol=1e-6)
    np.testing.assert_array_equal(ocam.qinverse(qf), z['qinv_f'])
    np.testing.assert_allclose(ocam.project_to_2d(z['Xc'], z['cams']), z['proj'], rtol=2e-6, atol=1e-6, equal_nan=True)
    np.testing.assert_allclose(ocam.project_to_2d_linear(z['Xc'], z['cams']), z['proj_linear'], rtol=2e-6, atol=1e-6,
                               equal_nan=True)
    assert np.isnan(z['proj'][0, 0, 1]).all(), 'golden must contain the 0/0 -> NaN case'
    ns = ocam.normalize_screen_coordinates(z['px'], w=1000, h=1002)
    assert ns.dtype == z['norm_sc'].dtype == np.float64  # NumPy-2 promotion quirk of camera.py:18
    np.testing.assert_allclose(ns, z['norm_sc'], atol=0)
    np.testing.assert_allclose(ocam.image_coordinates(ns, w=1000, h=1002), z['img_sc'], atol=0)
    # round trip (camera.py:28-34) and the 3x4-extrinsic einsum form (prepare_data_cmu_camera.py:61-66)
    np.testing.assert_allclose(ocam.camera_to_world(ocam.world_to_camera(X, q, t), q, t), X, atol=5e-6)
    Rm = ocam.quat_to_matrix(ocam.qinverse(qf).astype(np.float64))
    E = np.concatenate((Rm, -(Rm @ tf[..., None].astype(np.float64))), axis=-1)
    np.testing.assert_allclose(ocam.extrinsic_einsum(X.astype(np.float64), E), z['w2c_f'], atol=1e-5)


def test_loss_oracle():
    z = load_golden('loss.npz')
    pred, tgt = torch.from_numpy(z['pred']), torch.from_numpy(z['tgt'])
    p = pred.clone().requires_grad_(True)
    l = oloss.mpjpe(p, tgt)
    l.backward()
    np.testing.assert_allclose(l.detach().numpy(), z['mpjpe'], atol=1e-7)
    np.testing.assert_allclose(p.grad.numpy(), z['mpjpe_grad'], atol=1e-8)
    assert np.all(z['mpjpe_grad'][0, 0, 0] == 0)
    for wname in ('w_n', 'w_nt1', 'w_ntj'):
        w = torch.from_numpy(z[wname])
        p = pred.clone().requires_grad_(True)
        l = oloss.weighted_mpjpe(p, tgt, w)
        l.backward()
        np.testing.assert_allclose(l.detach().numpy(), z['wmpjpe_' + wname], atol=1e-7)
        np.testing.assert_allclose(p.grad.numpy(), z['wmpjpe_grad_' + wname], atol=1e-8)
    np.testing.assert_allclose(oloss.n_mpjpe(pred, tgt).numpy(), z['n_mpjpe'], atol=1e-7)
    J = pred.shape[2]
    P, T = z['pred'].reshape(-1, J, 3), z['tgt'].reshape(-1, J, 3)
    np.testing.assert_allclose(oloss.p_mpjpe(P.copy(), T.copy()), float(z['p_mpjpe']), rtol=1e-6)
    np.testing.assert_allclose(oloss.mean_velocity_error(P[:, 0], T[:, 0]), float(z['mve']), rtol=1e-6)


def test_projection_gradient_oracle():
    """Closed-form d project_to_2d / dX (oracle) against autograd of the reference (tests/golden/camera_grad.npz)."""
    from oracle import camera as ocam
    z = load_golden('camera_grad.npz')
    for key, lin in (('grad', False), ('grad_linear', True)):
        g = ocam.project_to_2d_grad(z['X'], z['cams'], z['W'], linear=lin)
        np.testing.assert_allclose(g, z[key], atol=2e-6)
    assert (z['grad'][0, 0, 0] == 0).all()                      # both ratios clamped
    assert z['grad'][0, 0, 2, 0] == 0 and z['grad'][0, 0, 2, 1] != 0


def test_unchunked_generator_composition_against_reference_generator():
    """The host restatement the GPU sequence-feeder tests use -- np.pad(..., 'edge') of the oracle's per-frame
    world_to_camera / project_to_2d, K @ [R | -R c] -- against the batches the reference's own UnchunkedGenerator yields
    (tests/golden/generator_unchunked.npz, generators.py:178-205)."""
    from oracle import camera as ocam
    z = load_golden('generator_unchunked.npz')
    intr = z['intrinsics']
    for tag in ('a', 'b'):
        pad, shift = [int(v) for v in z['params_' + tag]]
        for i in range(len(z['lens'])):
            X, q, t = z['world_%d' % i], z['q_%d' % i], z['t_%d' % i]
            xc = ocam.world_to_camera(X, q, t)
            p2 = ocam.project_to_2d(xc[None], intr[None])[0]
            want2 = np.pad(p2, ((pad + shift, pad - shift), (0, 0), (0, 0)), 'edge')
            np.testing.assert_allclose(want2[None], z['b2d_%s_%d' % (tag, i)], atol=2e-6)
            np.testing.assert_allclose((xc - xc[:, :1])[None], z['b3d_%s_%d' % (tag, i)], atol=2e-6)
            assert z['cam_%s_%d' % (tag, i)].shape == (1, X.shape[0] + 2 * pad, 3, 4)


def test_lifter_oracle_matches_the_reference_fixture():
    """oracle/lifter.py (StackedPoseLifter.py:37-56, CamLSTM.py:33-44) against tests/golden/lifter.npz, which
    tests/golden/make_golden_lifter.py produced by importing the real reference."""
    from oracle import lifter as ol
    z = load_golden('lifter.npz')
    J, F = 17, 3
    sd = {k: v.clone().requires_grad_(True) for k, v in ol.init_state(J, F, 3, 256, seed=21).items()}
    assert sorted(sd) == list(z['state_keys'])
    a, b, tgt = (torch.from_numpy(z[k]) for k in ('a', 'b', 'tgt'))
    y = ol.forward(sd, a, b, J, F)
    assert torch.equal(y.detach(), torch.from_numpy(z['y']))
    assert torch.equal(ol.forward(sd, a.squeeze(), b.squeeze(), J, F).detach(), torch.from_numpy(z['y_squeezed']))
    loss = torch.mean(torch.linalg.norm(y - tgt, dim=3))
    loss.backward()
    assert abs(loss.item() - float(z['loss'])) < 1e-6
    for k, p in sd.items():
        if 'grad/' + k in z.files:
            np.testing.assert_allclose(p.grad.numpy(), z['grad/' + k], rtol=1e-4, atol=1e-7)
        else:
            np.testing.assert_allclose(p.grad.numpy()[::8, ::8], z['grad_sample/' + k], rtol=1e-4, atol=1e-7)

    def probe(win_2d, win_cam):
        w = torch.arange(1, win_2d.shape[1] + 1, dtype=win_2d.dtype).view(1, -1, 1, 1)
        s2 = (win_2d * w).sum(dim=(1, 3))
        sc = (win_cam * w).sum(dim=(1, 2, 3)).view(-1, 1)
        return torch.stack([s2, s2 + sc, s2 - sc], dim=-1)
    x2, cam = torch.from_numpy(z['sw_x2']), torch.from_numpy(z['sw_cam'])
    np.testing.assert_allclose(ol.sliding_window(probe, x2, cam, 9, J, 3).numpy(), z['sw_out_w9'], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(ol.sliding_window(probe, x2, cam, 39, J, 3).numpy(), z['sw_out_w40'], rtol=1e-5, atol=1e-5)
