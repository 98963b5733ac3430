"""Drop-in for the reference's common/quaternion.py (qrot :10-24, qinverse :27-35) on CUDA tensors."""
import torch

from vp3d_b200 import native, ops


def qrot(q, v):
    """
    Rotate vector(s) v about the rotation described by quaternion(s) q.
    Expects a tensor of shape (*, 4) for q and a tensor of shape (*, 3) for v,
    where * denotes any number of dimensions.
    Returns a tensor of shape (*, 3).
    """
    assert q.shape[-1] == 4
    assert v.shape[-1] == 3
    assert q.shape[:-1] == v.shape[:-1]
    out, _ = ops.project_points(v, q=q, pts_per_q=1, mode=native.PT_ROTATE, want3=True)
    return out.view(v.shape).to(v.dtype)


def qinverse(q, inplace=False):
    # We assume the quaternion to be normalized: the inverse is the conjugate (a sign flip, no arithmetic)
    if inplace:
        q[..., 1:] *= -1
        return q
    out = q.clone()
    out[..., 1:].neg_()
    return out
