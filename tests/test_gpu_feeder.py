"""GPU window feeder (vp3d_project_windows) against the reference's ChunkedGenerator semantics restated in NumPy
(edge padding, pad / causal_shift window placement, RandomState(1234) epoch order) composed with the CPU oracle's
per-frame camera functions."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import camera as ocam  # noqa: E402
from vp3d_b200.feeder import DeviceWindowFeeder  # noqa: E402


def _data(n_seq=5, J=17, seed=0):
    rng = np.random.default_rng(seed)
    lens = rng.integers(40, 120, n_seq)
    X = [rng.normal(0, 0.3, (n, J, 3)).astype(np.float32) + np.array([0, 0, 4], np.float32) for n in lens]
    Q = []
    for n in lens:
        q = np.array([1, 0, 0, 0], np.float32) + rng.normal(0, 0.1, (n, 4)).astype(np.float32)
        Q.append(q / np.linalg.norm(q, axis=-1, keepdims=True))
    T = [rng.normal(0, 0.2, (n, 3)).astype(np.float32) for n in lens]
    cam = np.tile(np.array([2.29, 2.2876, 0.0251, 0.0289, -0.2071, 0.2478, -0.0031, -0.00098, -0.0014], np.float32),
                  (n_seq, 1))
    cam[:, 0] += rng.normal(0, 0.05, n_seq).astype(np.float32)
    return X, Q, T, cam


def _pad_chunk(a, start, end):     # generators.py:92-100
    lo, hi = max(start, 0), min(end, a.shape[0])
    pl, pr = lo - start, end - hi
    if pl or pr:
        return np.pad(a[lo:hi], ((pl, pr),) + ((0, 0),) * (a.ndim - 1), 'edge')
    return a[lo:hi]


@pytest.mark.parametrize('pad,shift,chunk', [(13, 0, 1), (13, 13, 1), (4, 0, 3)])
def test_feeder_matches_generator_semantics(pad, shift, chunk):
    X, Q, T, cam = _data()
    fd = DeviceWindowFeeder(X, Q, T, cam, batch_size=64, chunk_length=chunk, pad=pad, causal_shift=shift,
                            shuffle=True, random_seed=1234, want_cameras=True)
    rs = np.random.RandomState(1234)
    pairs = rs.permutation(fd.pairs)              # the reference's epoch order
    n_seen = 0
    for b_i, (cams, b3d, b2d) in enumerate(fd.next_epoch()):
        chunks = pairs[b_i * 64:(b_i + 1) * 64]
        assert b2d.shape == (len(chunks), chunk + 2 * pad, 17, 2) and b3d.shape == (len(chunks), chunk, 17, 3)
        for i, (s, a3, e3) in enumerate(chunks):
            a2, e2 = a3 - pad - shift, e3 + pad - shift
            xw, qw, tw = _pad_chunk(X[s], a2, e2), _pad_chunk(Q[s], a2, e2), _pad_chunk(T[s], a2, e2)
            xc = ocam.world_to_camera(xw, qw, tw)
            want2 = ocam.project_to_2d(xc[None], cam[s:s + 1])[0]
            np.testing.assert_allclose(b2d[i].cpu().numpy(), want2, atol=1e-5)
            xt = ocam.world_to_camera(_pad_chunk(X[s], a3, e3), _pad_chunk(Q[s], a3, e3), _pad_chunk(T[s], a3, e3))
            np.testing.assert_allclose(b3d[i].cpu().numpy(), xt - xt[:, :1], atol=2e-6)
            # K @ [R | -R c] reproduces the projection of the linear camera model on a world point
            if i == 0:
                P = cams[i].cpu().numpy().astype(np.float64)              # (window, 3, 4)
                xh = np.concatenate([xw, np.ones(xw.shape[:-1] + (1,), np.float32)], -1).astype(np.float64)
                proj = np.einsum('tij,tnj->tni', P, xh)
                uv = proj[..., :2] / proj[..., 2:]
                lin = ocam.project_to_2d_linear(xc[None], cam[s:s + 1])[0]
                ok = np.abs(xc[..., :2] / xc[..., 2:]).max(-1) < 1       # clamp not active
                np.testing.assert_allclose(uv[ok], lin[ok], atol=1e-4)
        n_seen += len(chunks)
        if b_i >= 2:
            break
    assert n_seen > 0
    assert fd.num_frames() == fd.num_batches * 64


def test_feeder_drives_a_training_step():
    from common.loss import mpjpe
    from common.models.TemporalModel import TemporalModelOptimized1f
    X, Q, T, cam = _data(n_seq=8, seed=3)
    torch.manual_seed(0)
    m = TemporalModelOptimized1f(17, 2, 17, [3, 3, 3], dropout=0.25, channels=1024).cuda().train()
    fd = DeviceWindowFeeder(X, Q, T, cam, batch_size=128, pad=13, shuffle=True)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, amsgrad=True)
    losses = []
    for _ in range(3):
        for _, b3d, b2d in fd.next_epoch():
            opt.zero_grad()
            loss = mpjpe(m(b2d), b3d)
            loss.backward()
            opt.step()
            losses.append(loss.item())
    assert np.isfinite(losses).all() and losses[-1] < losses[0]


@pytest.mark.parametrize('pad,shift', [(13, 0), (13, 13), (121, 0)])
def test_sequence_feeder_matches_unchunked_generator_semantics(pad, shift):
    """UnchunkedGenerator.next_epoch (generators.py:178-205): whole sequences, np.pad(..., 'edge') by
    (pad + shift, pad - shift); fed to the eval-mode model the output has one pose per un-padded frame."""
    from vp3d_b200.feeder import DeviceSequenceFeeder
    X, Q, T, cam = _data(n_seq=3, seed=5)
    info = [{'cam_velocity': float(i)} for i in range(3)]
    fd = DeviceSequenceFeeder(X, Q, T, cam, pad=pad, causal_shift=shift, want_cameras=True, seq_info=info)
    assert fd.num_frames() == sum(x.shape[0] for x in X)
    n = 0
    for s, (cams, b3d, b2d, meta) in enumerate(fd.next_epoch()):
        L = X[s].shape[0]
        assert b2d.shape == (1, L + 2 * pad, 17, 2) and b3d.shape == (1, L, 17, 3) and cams.shape == (1, L + 2 * pad, 3, 4)
        assert meta is info[s]
        xc = ocam.world_to_camera(X[s], Q[s], T[s])
        p2 = ocam.project_to_2d(xc[None], cam[s:s + 1])[0]
        want2 = np.pad(p2, ((pad + shift, pad - shift), (0, 0), (0, 0)), 'edge')
        np.testing.assert_allclose(b2d[0].cpu().numpy(), want2, atol=1e-5)
        np.testing.assert_allclose(b3d[0].cpu().numpy(), xc - xc[:, :1], atol=2e-6)
        c = cams[0].cpu().numpy()
        np.testing.assert_array_equal(c[:pad + shift + 1], np.broadcast_to(c[pad + shift], (pad + shift + 1, 3, 4)))
        n += 1
    assert n == 3


def test_sequence_feeder_drives_evaluation():
    from common.loss import mpjpe
    from common.models.TemporalModel import TemporalModel
    from vp3d_b200.feeder import DeviceSequenceFeeder
    X, Q, T, cam = _data(n_seq=2, seed=7)
    torch.manual_seed(0)
    m = TemporalModel(17, 2, 17, [3, 3, 3], channels=1024).cuda().eval()
    fd = DeviceSequenceFeeder(X, Q, T, cam, pad=(m.receptive_field() - 1) // 2)
    with torch.no_grad():
        for _, b3d, b2d, _ in fd.next_epoch():
            pred = m(b2d)
            assert pred.shape == b3d.shape        # run.py:711-734: one pose per frame of the sequence
            assert np.isfinite(mpjpe(pred, b3d).item())


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_sequence_feeder_against_the_reference_unchunked_generator(tag):
    """DeviceSequenceFeeder against what the reference's own UnchunkedGenerator yields for the same sequences
    (tests/golden/generator_unchunked.npz, made by importing common/generators.py and the reference camera functions):
    padded 2-D input, root-relative 3-D target, padded K @ [R|t] matrices."""
    from conftest import load_golden
    from vp3d_b200.feeder import DeviceSequenceFeeder
    z = load_golden('generator_unchunked.npz')
    n = len(z['lens'])
    pad, shift = [int(v) for v in z['params_' + tag]]
    X, Q, T = [z['world_%d' % i] for i in range(n)], [z['q_%d' % i] for i in range(n)], [z['t_%d' % i] for i in range(n)]
    cam = np.tile(z['intrinsics'], (n, 1))
    fd = DeviceSequenceFeeder(X, Q, T, cam, pad=pad, causal_shift=shift, want_cameras=True)
    for i, (cams, b3d, b2d, _info) in enumerate(fd.next_epoch()):
        np.testing.assert_allclose(b2d.cpu().numpy(), z['b2d_%s_%d' % (tag, i)], atol=1e-5)
        np.testing.assert_allclose(b3d.cpu().numpy(), z['b3d_%s_%d' % (tag, i)], atol=2e-6)
        np.testing.assert_allclose(cams.cpu().numpy(), z['cam_%s_%d' % (tag, i)], atol=1e-5)
