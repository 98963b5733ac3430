"""Drop-in for the reference's common/utils.py (wrap :12-42, deterministic_random :44-47).

`wrap` keeps its NumPy-in / NumPy-out contract, but the wrapped function runs on the GPU: arrays are staged to the
current CUDA device, the sm_100a kernel runs there, and results come back as NumPy arrays of the input's dtype.
"""
import hashlib
import warnings

import numpy as np
import torch

_warned_f64 = False


def wrap(func, *args, unsqueeze=False):
    """
    Wrap a torch function so it can be called with NumPy arrays.
    Input and return types are seamlessly converted.
    """
    if not torch.cuda.is_available():
        raise RuntimeError('vp3d_b200: wrap() stages NumPy arrays to a CUDA device; no GPU is visible and there is '
                           'no CPU fallback')
    global _warned_f64
    np_dtype = None
    converted = []
    for arg in args:
        if type(arg) == np.ndarray:
            if np_dtype is None and arg.dtype.kind == 'f':
                np_dtype = arg.dtype
                if np_dtype == np.float64 and not _warned_f64:
                    # documented deviation ("wrap and float64"): the reference's torch-CPU path would
                    # compute float64 arrays in float64; the sm_100a kernels compute in fp32 and the result is cast back
                    _warned_f64 = True
                    warnings.warn('vp3d_b200.wrap: float64 arrays are computed in float32 on the GPU and returned as '
                                  'float64 (about 1e-7 relative precision); pass float32 arrays to silence this',
                                  RuntimeWarning, stacklevel=2)
            t = torch.from_numpy(np.ascontiguousarray(arg)).cuda()
            converted.append(t.unsqueeze(0) if unsqueeze else t)
        else:
            converted.append(arg)

    result = func(*converted)

    def back(res):
        if type(res) != torch.Tensor:
            return res
        if unsqueeze:
            res = res.squeeze(0)
        out = res.detach().cpu().numpy()
        if np_dtype is not None and out.dtype.kind == 'f' and out.dtype != np_dtype:
            out = out.astype(np_dtype)  # kernels compute in fp32; hand back the caller's float type
        return out

    if isinstance(result, tuple):
        return tuple(back(r) for r in result)
    return back(result)


def deterministic_random(min_value, max_value, data):
    digest = hashlib.sha256(data.encode()).digest()
    raw_value = int.from_bytes(digest[:4], byteorder='little', signed=False)
    return int(raw_value / (2**32 - 1) * (max_value - min_value)) + min_value
