"""Drop-in for the reference's common/camera.py, backed by the fused sm_100a projection kernel (K5).

Same names, argument meaning and AssertionError behaviour as the reference (camera.py:14-90). NumPy callers
(run.py:78,117,123,845,853; data/prepare_data_h36m.py:160-162) keep NumPy in / NumPy out through `wrap`; CUDA tensors
are processed in place on the device. Additions beyond the reference API:
  * world_to_camera / camera_to_world also accept per-frame R (T,4) and t (T,3) -- the dynamic-camera case the
    reference's np.tile cannot express;
  * world_to_image(): world -> camera -> image plane in ONE kernel launch.
"""
import numpy as np
import torch

from common.quaternion import qinverse, qrot  # noqa: F401  (re-exported like the reference module does)
from common.utils import wrap
from vp3d_b200 import native, ops


def normalize_screen_coordinates(X, w, h):
    assert X.shape[-1] == 2

    # Normalize so that [0, w] is mapped to [-1, 1], while preserving the aspect ratio
    return X / w * 2 - [1, h / w]


def image_coordinates(X, w, h):
    assert X.shape[-1] == 2

    # Reverse camera frame normalization
    return (X + [1, h / w]) * w / 2


def _rigid(X, R, t, mode):
    """X (..., J, 3); R (4,) with t (3,) for a static camera, or R (..., 4) / t (..., 3) matching X.shape[:-2]
    for one camera pose per frame."""
    assert X.shape[-1] == 3
    assert R.shape[-1] == 4 and t.shape[-1] == 3
    n_pts = X.numel() // 3
    if R.dim() == 1 and t.dim() > 1:
        # static orientation, one translation per frame: the reference computes X - t by broadcasting, so t is
        # (..., 1, 3) there; give every frame its own copy of R and take the per-frame path
        tt = t.squeeze(-2) if (t.dim() == X.dim() and t.shape[-2] == 1) else t
        assert tuple(tt.shape[:-1]) == tuple(X.shape[:-2]), \
            'a per-frame translation needs one row per frame of X (shape X.shape[:-2] + (3,) or (..., 1, 3))'
        R, t = R.expand(tuple(X.shape[:-2]) + (4,)).contiguous(), tt.contiguous()
    if R.dim() == 1:
        assert t.dim() == 1, 'a static camera is one quaternion (4,) and one translation (3,)'
        pts_per_q = max(n_pts, 1)
    else:
        assert tuple(R.shape[:-1]) == tuple(X.shape[:-2]) and tuple(t.shape[:-1]) == tuple(X.shape[:-2]), \
            'per-frame cameras need one quaternion and one translation per frame of X'
        pts_per_q = X.shape[-2]
    out, _ = ops.project_points(X, q=R, t=t, pts_per_q=pts_per_q, mode=mode, want3=True)
    return out.view(X.shape)


def _world_to_camera_t(X, R, t):
    return _rigid(X, R, t, native.PT_WORLD_TO_CAMERA)


def _camera_to_world_t(X, R, t):
    return _rigid(X, R, t, native.PT_CAMERA_TO_WORLD)


def _dispatch_rigid(fn, X, R, t):
    if isinstance(X, torch.Tensor):
        R = torch.as_tensor(R, dtype=torch.float32, device=X.device)
        t = torch.as_tensor(t, dtype=torch.float32, device=X.device)
        return fn(X, R, t)
    return wrap(fn, np.asarray(X), np.asarray(R, dtype=np.float32), np.asarray(t, dtype=np.float32))


def world_to_camera(X, R, t):
    # qrot(qinverse(R), X - t): rotate by the conjugate after removing the camera position
    return _dispatch_rigid(_world_to_camera_t, X, R, t)


def camera_to_world(X, R, t):
    # qrot(R, X) + t
    return _dispatch_rigid(_camera_to_world_t, X, R, t)


def _check_projection_args(X, camera_params):
    assert X.shape[-1] == 3
    assert len(camera_params.shape) == 2
    assert camera_params.shape[-1] == 9
    assert X.shape[0] == camera_params.shape[0]


def project_to_2d(X, camera_params):
    """
    Project 3D points to 2D using the Human3.6M camera projection function.
    This is a differentiable and batched reimplementation of the original MATLAB script.

    Arguments:
    X -- 3D points in *camera space* to transform (N, *, 3)
    camera_params -- intrinsic parameteres (N, 2+2+3+2=9)
    """
    _check_projection_args(X, camera_params)
    per_cam = max(X.numel() // 3 // max(X.shape[0], 1), 1)
    return ops.project_2d(X, camera_params, per_cam, linear=False).to(X.dtype)   # differentiable wrt X


def project_to_2d_linear(X, camera_params):
    """
    Project 3D points to 2D using only linear parameters (focal length and principal point).

    Arguments:
    X -- 3D points in *camera space* to transform (N, *, 3)
    camera_params -- intrinsic parameteres (N, 2+2+3+2=9)
    """
    _check_projection_args(X, camera_params)
    per_cam = max(X.numel() // 3 // max(X.shape[0], 1), 1)
    return ops.project_2d(X, camera_params, per_cam, linear=True).to(X.dtype)    # differentiable wrt X


def world_to_image(X, R, t, camera_params, linear=False, return_camera_space=True, exact=False):
    """Fused dynamic-camera projection: X (N, T, J, 3) world-space joints, R (N, T, 4) / t (N, T, 3) one camera pose
    per frame, camera_params (N, 9) per sequence or (N, T, 9) per frame.
    Equals project_to_2d(world_to_camera(X, R, t), camera_params) evaluated frame by frame, in one kernel launch.
    Returns (X_camera or None, x_2d). `exact=True` evaluates operation by operation like the reference functions above
    (bit-faithful, instruction bound); the default fuses multiply-adds (<= 1e-6 relative difference, ~2x faster)."""
    assert X.dim() == 4 and X.shape[-1] == 3
    assert tuple(R.shape) == tuple(X.shape[:2]) + (4,) and tuple(t.shape) == tuple(X.shape[:2]) + (3,)
    assert camera_params.shape[-1] == 9 and camera_params.shape[0] == X.shape[0]
    J = X.shape[2]
    if camera_params.dim() == 2:
        per_cam = X.shape[1] * J
    else:
        assert tuple(camera_params.shape[:2]) == tuple(X.shape[:2])
        per_cam = J
    mode = native.PT_WORLD_TO_CAMERA | native.PT_PROJECT | (native.PT_LINEAR if linear else 0) | \
        (0 if exact else native.PT_FAST)
    return ops.project_points(X, q=R, t=t, cam=camera_params, pts_per_q=J, pts_per_cam=per_cam, mode=mode,
                              want3=return_camera_space, want2=True)
