"""Drop-in for the reference's common/loss.py.

mpjpe / weighted_mpjpe / n_mpjpe run as fused sm_100a reductions (K6) on CUDA tensors and are differentiable
(run.py:480-485). p_mpjpe and mean_velocity_error are evaluation metrics that the reference computes in NumPy on the
host (run.py:749-756): NumPy arrays keep exactly those semantics here, CUDA tensors are reduced on the device
(vp3d_p_mpjpe_fwd: per-pose 3x3 SVD Procrustes; vp3d_velocity_error) and come back as 0-dim tensors.
"""
import numpy as np
import torch

from vp3d_b200 import ops


def mpjpe(predicted, target):
    """
    Mean per-joint position error (i.e. mean Euclidean distance),
    often referred to as "Protocol #1" in many papers.
    """
    assert predicted.shape == target.shape
    return ops.mpjpe(predicted, target)


def weighted_mpjpe(predicted, target, w):
    """
    Weighted mean per-joint position error (i.e. mean Euclidean distance)
    """
    assert predicted.shape == target.shape
    assert w.shape[0] == predicted.shape[0]
    return ops.mpjpe(predicted, target, w)


def p_mpjpe(predicted, target):
    """
    Pose error: MPJPE after rigid alignment (scale, rotation, and translation),
    often referred to as "Protocol #2" in many papers.
    """
    assert predicted.shape == target.shape
    if isinstance(predicted, torch.Tensor):
        return ops.p_mpjpe(predicted, target)       # CUDA tensors stay on the device (vp3d_p_mpjpe_fwd)

    # centre both point sets and bring them to unit Frobenius norm
    tgt_mean = target.mean(axis=1, keepdims=True)
    prd_mean = predicted.mean(axis=1, keepdims=True)
    tgt_c = target - tgt_mean
    prd_c = predicted - prd_mean
    tgt_scale = np.sqrt((tgt_c ** 2).sum(axis=(1, 2), keepdims=True))
    prd_scale = np.sqrt((prd_c ** 2).sum(axis=(1, 2), keepdims=True))
    tgt_c = tgt_c / tgt_scale
    prd_c = prd_c / prd_scale

    # orthogonal Procrustes: rotation from the SVD of the 3x3 cross-covariance, reflections removed
    U, sing, Vt = np.linalg.svd(np.matmul(tgt_c.transpose(0, 2, 1), prd_c))
    V = Vt.transpose(0, 2, 1)
    Ut = U.transpose(0, 2, 1)
    flip = np.sign(np.linalg.det(np.matmul(V, Ut)))
    V[:, :, -1] *= flip[:, None]
    sing[:, -1] *= flip
    rot = np.matmul(V, Ut)

    gain = sing.sum(axis=1)[:, None, None] * tgt_scale / prd_scale
    offset = tgt_mean - gain * np.matmul(prd_mean, rot)
    aligned = gain * np.matmul(predicted, rot) + offset
    return np.mean(np.linalg.norm(aligned - target, axis=len(target.shape) - 1))


def n_mpjpe(predicted, target):
    """
    Normalized MPJPE (scale only), adapted from:
    https://github.com/hrhodin/UnsupervisedGeometryAwareRepresentationLearning/blob/master/losses/poses.py
    """
    assert predicted.shape == target.shape
    return ops.n_mpjpe(predicted, target)


def mean_velocity_error(predicted, target):
    """
    Mean per-joint velocity error (i.e. mean Euclidean distance of the 1st derivative)
    """
    assert predicted.shape == target.shape
    if isinstance(predicted, torch.Tensor):
        return ops.mean_velocity_error(predicted, target)   # CUDA tensors stay on the device (vp3d_velocity_error)
    dv = np.diff(predicted, axis=0) - np.diff(target, axis=0)
    return np.mean(np.linalg.norm(dv, axis=len(target.shape) - 1))


def reprojection_mpjpe(predicted_3d, camera_params, target_2d, trajectory=None, linear=False):
    """Addition beyond the reference API: mpjpe(project_to_2d(predicted_3d + trajectory, camera_params), target_2d)
    -- the reprojection term of upstream VideoPose3D's semi-supervised step, built from this fork's camera.py:37-67 and
    loss.py:11-17 -- as ONE fused kernel per direction (vp3d_reproj_mpjpe_fwd / _bwd): the 2-D projection never reaches
    memory. predicted_3d (N, ..., J, 3) CUDA, trajectory (N, ..., 1, 3) or None, camera_params (N, 9),
    target_2d (N, ..., J, 2). Differentiable wrt predicted_3d and trajectory."""
    return ops.reproj_mpjpe(predicted_3d, camera_params, target_2d, traj=trajectory, linear=linear)
