"""CPU oracle for common/loss.py.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

torch-CPU restatement for the differentiable losses (so autograd provides the gradient oracle, as in the
reference) and NumPy for the evaluation-only protocol-2 / velocity metrics.
"""
import numpy as np
import torch


def mpjpe(predicted, target):
    """loss.py:11-17"""
    assert predicted.shape == target.shape
    return torch.mean(torch.linalg.norm(predicted - target, dim=target.dim() - 1))


def weighted_mpjpe(predicted, target, w):
    """loss.py:21-27"""
    assert predicted.shape == target.shape
    assert w.shape[0] == predicted.shape[0]
    return torch.mean(w * torch.linalg.norm(predicted - target, dim=target.dim() - 1))


def n_mpjpe(predicted, target):
    """loss.py:70-80"""
    assert predicted.shape == target.shape
    norm_predicted = torch.mean(torch.sum(predicted ** 2, dim=3, keepdim=True), dim=2, keepdim=True)
    norm_target = torch.mean(torch.sum(target * predicted, dim=3, keepdim=True), dim=2, keepdim=True)
    scale = norm_target / norm_predicted
    return mpjpe(scale * predicted, target)


def p_mpjpe(predicted, target):
    """loss.py:29-68: MPJPE after the optimal similarity transform (Procrustes via batched SVD)."""
    assert predicted.shape == target.shape
    mu_t = np.mean(target, axis=1, keepdims=True)
    mu_p = np.mean(predicted, axis=1, keepdims=True)
    t0 = target - mu_t
    p0 = predicted - mu_p
    n_t = np.sqrt(np.sum(t0 ** 2, axis=(1, 2), keepdims=True))
    n_p = np.sqrt(np.sum(p0 ** 2, axis=(1, 2), keepdims=True))
    t0 = t0 / n_t
    p0 = p0 / n_p
    H = np.matmul(t0.transpose(0, 2, 1), p0)
    U, s, Vt = np.linalg.svd(H)
    V = Vt.transpose(0, 2, 1)
    R = np.matmul(V, U.transpose(0, 2, 1))
    sign = np.sign(np.expand_dims(np.linalg.det(R), axis=1))
    V[:, :, -1] *= sign
    s[:, -1] *= sign.flatten()
    R = np.matmul(V, U.transpose(0, 2, 1))
    tr = np.expand_dims(np.sum(s, axis=1, keepdims=True), axis=2)
    a = tr * n_t / n_p
    t = mu_t - a * np.matmul(mu_p, R)
    aligned = a * np.matmul(predicted, R) + t
    return np.mean(np.linalg.norm(aligned - target, axis=target.ndim - 1))


def mean_velocity_error(predicted, target):
    """loss.py:82-91"""
    assert predicted.shape == target.shape
    vp = np.diff(predicted, axis=0)
    vt = np.diff(target, axis=0)
    return np.mean(np.linalg.norm(vp - vt, axis=target.ndim - 1))
