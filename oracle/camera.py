"""CPU oracle for common/camera.py and common/quaternion.py.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

NumPy float32 restatement, one NumPy operation per torch operation of the reference, so intermediate rounding is
the same as the reference's unfused ATen evaluation.
"""
import numpy as np


def _cross(a, b):
    # torch.cross(a, b, dim=-1): quaternion.py:22-23
    return np.stack((a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]), axis=-1)


def qrot(q, v):
    """quaternion.py:10-24"""
    assert q.shape[-1] == 4
    assert v.shape[-1] == 3
    assert q.shape[:-1] == v.shape[:-1]
    qvec = q[..., 1:]
    uv = _cross(qvec, v)
    uuv = _cross(qvec, uv)
    return v + 2 * (q[..., :1] * uv + uuv)


def qinverse(q):
    """quaternion.py:27-35 (non-inplace branch)"""
    return np.concatenate((q[..., :1], -q[..., 1:]), axis=-1)


def world_to_camera(X, R, t):
    """camera.py:28-30, extended to per-frame R (T,4) / t (T,3) by explicit broadcasting over joints
    (the reference's np.tile only accepts one static quaternion)."""
    Rt = qinverse(np.asarray(R, dtype=X.dtype))
    t = np.asarray(t, dtype=X.dtype)
    if Rt.ndim == 1:
        Rt_b = np.tile(Rt, (*X.shape[:-1], 1))
        return qrot(Rt_b, X - t)
    Rt_b = np.broadcast_to(Rt[..., None, :], (*X.shape[:-1], 4))
    return qrot(np.ascontiguousarray(Rt_b), X - t[..., None, :])


def camera_to_world(X, R, t):
    """camera.py:33-34 with the same per-frame extension."""
    R = np.asarray(R, dtype=X.dtype)
    t = np.asarray(t, dtype=X.dtype)
    if R.ndim == 1:
        return qrot(np.tile(R, (*X.shape[:-1], 1)), X) + t
    R_b = np.broadcast_to(R[..., None, :], (*X.shape[:-1], 4))
    return qrot(np.ascontiguousarray(R_b), X) + t[..., None, :]


def _clamp(x):
    # torch.clamp keeps NaN; np.clip does too
    return np.clip(x, -1, 1)


def project_to_2d(X, camera_params):
    """camera.py:37-67.  X (N, *, 3), camera_params (N, 9)."""
    assert X.shape[-1] == 3
    assert camera_params.ndim == 2
    assert camera_params.shape[-1] == 9
    assert X.shape[0] == camera_params.shape[0]
    cp = camera_params
    while cp.ndim < X.ndim:
        cp = cp[:, None]
    f, c, k, p = cp[..., :2], cp[..., 2:4], cp[..., 4:7], cp[..., 7:]
    with np.errstate(divide='ignore', invalid='ignore'):
        XX = _clamp(X[..., :2] / X[..., 2:])
    r2 = np.sum(XX[..., :2] ** 2, axis=-1, keepdims=True)
    radial = 1 + np.sum(k * np.concatenate((r2, r2 ** 2, r2 ** 3), axis=-1), axis=-1, keepdims=True)
    tan = np.sum(p * XX, axis=-1, keepdims=True)
    XXX = XX * (radial + tan) + p * r2
    return f * XXX + c


def project_to_2d_linear(X, camera_params):
    """camera.py:69-90"""
    assert X.shape[-1] == 3
    assert camera_params.ndim == 2
    assert camera_params.shape[-1] == 9
    assert X.shape[0] == camera_params.shape[0]
    cp = camera_params
    while cp.ndim < X.ndim:
        cp = cp[:, None]
    f, c = cp[..., :2], cp[..., 2:4]
    with np.errstate(divide='ignore', invalid='ignore'):
        XX = _clamp(X[..., :2] / X[..., 2:])
    return f * XX + c


def normalize_screen_coordinates(X, w, h):
    """camera.py:14-18"""
    assert X.shape[-1] == 2
    return X / w * 2 - [1, h / w]


def image_coordinates(X, w, h):
    """camera.py:21-25"""
    assert X.shape[-1] == 2
    return (X + [1, h / w]) * w / 2


def extrinsic_einsum(X, E):
    """Dynamic-camera form used by data/prepare_data_cmu_camera.py:61-66: per-frame 3x4 [R|t] applied to homogeneous
    world points (the y-flip of that script is a dataset convention and is not part of this oracle)."""
    Xh = np.concatenate((X, np.ones((*X.shape[:-1], 1), dtype=X.dtype)), axis=-1)
    return np.einsum('tij,tnj->tni', E, Xh)


def quat_to_matrix(q):
    """Rotation matrix of a unit quaternion (w,x,y,z); used only to cross-check qrot against the 3x4 einsum form."""
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    return np.stack((np.stack((1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)), -1),
                     np.stack((2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)), -1),
                     np.stack((2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)), -1)), -2)


def project_to_2d_grad(X, camera_params, grad_out, linear=False):
    """d sum(grad_out * project_to_2d(X)) / dX in closed form -- what torch autograd derives for camera.py:54-67
    (:85-90 when linear). torch.clamp passes the gradient where -1 <= ratio <= 1 (bounds included)."""
    cp = camera_params
    while cp.ndim < X.ndim:
        cp = cp[:, None]
    fx, fy = cp[..., 0], cp[..., 1]
    k1, k2, k3, p1, p2 = (cp[..., i] for i in range(4, 9))
    x, y, z = X[..., 0], X[..., 1], X[..., 2]
    with np.errstate(divide='ignore', invalid='ignore'):
        rx, ry = x / z, y / z
    xx, yy = np.clip(rx, -1, 1), np.clip(ry, -1, 1)
    a, b = fx * grad_out[..., 0], fy * grad_out[..., 1]
    if linear:
        gxx, gyy = a, b
    else:
        r2 = xx * xx + yy * yy
        s = 1 + k1 * r2 + k2 * r2 ** 2 + k3 * r2 ** 3 + p1 * xx + p2 * yy
        gs = a * xx + b * yy
        gr2 = a * p1 + b * p2 + gs * (k1 + 2 * k2 * r2 + 3 * k3 * r2 ** 2)
        gxx = a * s + gs * p1 + gr2 * 2 * xx
        gyy = b * s + gs * p2 + gr2 * 2 * yy
    gxx = np.where((rx >= -1) & (rx <= 1), gxx, 0)
    gyy = np.where((ry >= -1) & (ry <= 1), gyy, 0)
    return np.stack([gxx / z, gyy / z, -(gxx * x + gyy * y) / (z * z)], axis=-1).astype(X.dtype)
