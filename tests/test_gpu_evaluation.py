"""vp3d_b200.evaluation (run.py:677-774 evaluate + :946-983 PMCC table on the device) against a host restatement that
follows the reference line by line with the CPU oracle's metrics."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from common.models.TemporalModel import TemporalModel  # noqa: E402
from oracle import camera as ocam  # noqa: E402
from oracle import loss as oloss  # noqa: E402
from oracle import temporal_model as otm  # noqa: E402
from vp3d_b200 import evaluation  # noqa: E402
from vp3d_b200.feeder import DeviceSequenceFeeder  # noqa: E402


def _sequences(n_seq, J, seed):
    rng = np.random.default_rng(seed)
    lens = rng.integers(60, 140, n_seq)
    X = [np.cumsum(rng.normal(0, 0.02, (n, J, 3)), axis=0).astype(np.float32) + np.array([0, 0, 4], np.float32) for n in lens]
    Q = []
    for n in lens:
        q = np.array([1, 0, 0, 0], np.float32) + np.cumsum(rng.normal(0, 0.004, (n, 4)), axis=0).astype(np.float32)
        Q.append((q / np.linalg.norm(q, axis=-1, keepdims=True)).astype(np.float32))
    T = [np.cumsum(rng.normal(0, 0.01, (n, 3)), axis=0).astype(np.float32) for n in lens]
    cam = np.tile(np.array([1.5625, 1.5625, 0, 0, 0, 0, 0, 0, 0], np.float32), (n_seq, 1))    # CMU intrinsics
    info = [{k: rng.normal(0, 1, 3) for k in evaluation.CAM_KEYS} for _ in lens]
    return X, Q, T, cam, info


def test_device_evaluation_matches_the_reference_loop():
    fw = [3, 3, 3]
    J = 17
    sd = otm.init_state(J, 2, J, fw, channels=256, seed=81)
    m = TemporalModel(J, 2, J, fw, channels=256)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    X, Q, T, cam, info = _sequences(6, J, 82)
    pad = (m.receptive_field() - 1) // 2
    fd = DeviceSequenceFeeder(X, Q, T, cam, pad=pad, seq_info=info)
    got = evaluation.evaluate(m, fd)

    # host restatement of run.py:697-767 on the same sequences, metrics from the CPU oracle, model from the CPU oracle
    tot = np.zeros(4)
    N = 0
    e1_seq, motion_seq = [], []
    for s in range(len(X)):
        xc = ocam.world_to_camera(X[s], Q[s], T[s])
        p2 = ocam.project_to_2d(xc[None], cam[s:s + 1])[0]
        b2d = np.pad(p2, ((pad, pad), (0, 0), (0, 0)), 'edge')[None]
        b3d = (xc - xc[:, :1])[None]
        with torch.no_grad():
            pred = otm.forward(sd, torch.from_numpy(b2d), fw)
        tgt = torch.from_numpy(b3d)
        n = b3d.shape[0] * b3d.shape[1]
        e1 = oloss.mpjpe(pred, tgt).item()
        e3 = oloss.n_mpjpe(pred, tgt).item()
        pn, tn = pred.numpy().reshape(-1, J, 3), b3d.reshape(-1, J, 3)
        e2 = oloss.p_mpjpe(pn, tn)
        ev = oloss.mean_velocity_error(pn, tn)
        tot += n * np.array([e1, e2, e3, ev])
        N += n
        e1_seq.append(e1)
        motion_seq.append(np.mean(np.linalg.norm(np.diff(b3d, axis=1), axis=-1).squeeze(), axis=(0, 1)))
    want = tot / N * 1000
    assert got['frames'] == N
    for k, w in zip(('e1', 'e2', 'e3', 'ev'), want):
        assert abs(got[k] - w) < 2e-3 * abs(w) + 1e-2, (k, got[k], w)      # fp16 operands vs fp32 oracle; values in mm
    np.testing.assert_allclose(got['e1_per_seq'].cpu().numpy(), e1_seq, rtol=2e-3)
    np.testing.assert_allclose(got['pose_motion_per_seq'].cpu().numpy(), motion_seq, rtol=1e-5)

    # PMCC table: column-wise Pearson coefficients (what run.py:979-983 prints claims to be) and the reference's own
    # row-wise quirk (np.corrcoef(corr_data), run.py:966)
    table = np.stack([np.array(e1_seq)] + [np.linalg.norm(np.array([i[k] for i in info]), axis=1) for k in evaluation.CAM_KEYS]
                     + [np.array(motion_seq)], axis=1)
    pm = evaluation.camera_motion_pmcc(np.array(e1_seq), info, np.array(motion_seq))
    np.testing.assert_allclose(list(pm.values()), np.corrcoef(table, rowvar=False)[0, 1:], atol=1e-9)
    pq = evaluation.camera_motion_pmcc(np.array(e1_seq), info, np.array(motion_seq), reference_quirk=True)
    np.testing.assert_allclose(list(pq.values()), np.corrcoef(table)[0, 1:6], atol=1e-9)
    r = evaluation.run_evaluation(m, {'walk': fd})
    assert abs(r['e1'] - got['e1']) < 1e-6 and set(r['pmcc']) == set(pm)
