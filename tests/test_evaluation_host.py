"""Host-side logic of vp3d_b200.evaluation: the camera-motion correlation table of run.py:946-983 (no GPU needed)."""
import numpy as np

from vp3d_b200 import evaluation


def _table(e1, info, motion):
    return np.stack([e1] + [np.linalg.norm(np.array([i[k] for i in info]), axis=1) for k in evaluation.CAM_KEYS]
                    + [motion], axis=1)


def test_pmcc_table_column_wise_and_reference_quirk():
    rng = np.random.default_rng(0)
    n = 9
    info = [{k: rng.normal(0, 1, 3) for k in evaluation.CAM_KEYS} for _ in range(n)]
    e1, motion = rng.random(n), rng.random(n)
    t = _table(e1, info, motion)
    got = evaluation.camera_motion_pmcc(e1, info, motion)
    assert list(got) == ['cam_velocity', 'cam_acceleration', 'cam_angular_velocity', 'cam_angular_acceleration',
                         'pose_motion']
    np.testing.assert_allclose(list(got.values()), np.corrcoef(t, rowvar=False)[0, 1:], atol=1e-12)
    # the reference's np.corrcoef(corr_data) correlates rows (sequences), run.py:966-972
    quirk = evaluation.camera_motion_pmcc(e1, info, motion, reference_quirk=True)
    np.testing.assert_allclose(list(quirk.values()), np.corrcoef(t)[0, 1:6], atol=1e-12)


def test_pmcc_detects_a_planted_correlation():
    rng = np.random.default_rng(1)
    n = 40
    speed = rng.random(n) + 0.1
    info = [{k: rng.normal(0, 1, 3) for k in evaluation.CAM_KEYS} for _ in range(n)]
    for i in range(n):
        d = rng.normal(0, 1, 3)
        info[i]['cam_velocity'] = d / np.linalg.norm(d) * speed[i]
    e1 = 0.05 + 0.1 * speed + rng.normal(0, 1e-3, n)      # error grows with camera speed
    got = evaluation.camera_motion_pmcc(e1, info, rng.random(n))
    assert got['cam_velocity'] > 0.99 and abs(got['cam_acceleration']) < 0.5
