"""ctypes binding of lib/libvp3d_b200.so (the C ABI declared in include/vp3d_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os
import threading

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get('VP3D_LIB_PATH') or os.path.join(_PKG_DIR, 'lib', 'libvp3d_b200.so')   # override: A/B builds

F16, BF16, TF32 = 0, 1, 2
DTYPE_NAMES = {'fp16': F16, 'float16': F16, 'half': F16, 'bf16': BF16, 'bfloat16': BF16, 'tf32': TF32}

PT_WORLD_TO_CAMERA, PT_CAMERA_TO_WORLD, PT_ROTATE, PT_CONJ, PT_PROJECT, PT_LINEAR, PT_FAST = 1, 2, 4, 8, 16, 32, 64


class ConvArgs(C.Structure):
    """struct vp3d_conv_args (include/vp3d_b200.h)"""
    _fields_ = [
        ('dtype', C.c_int), ('block_n', C.c_int),
        ('a', C.c_void_p), ('a_seqs', C.c_longlong), ('a_rows', C.c_longlong), ('a_kdim', C.c_longlong),
        ('a_row_stride', C.c_longlong), ('a_seq_stride', C.c_longlong), ('a_row_off', C.c_longlong),
        ('w', C.c_void_p), ('n_pad', C.c_longlong), ('k_total', C.c_longlong), ('taps', C.c_int),
        ('tap_row_step', C.c_int), ('k_per_tap', C.c_longlong),
        ('w_mn_major', C.c_int), ('w_row_stride', C.c_longlong), ('w_tap_col_step', C.c_longlong),
        ('rows_out', C.c_longlong), ('out', C.c_void_p), ('out_f32', C.c_int), ('out_row_stride', C.c_longlong),
        ('out_seq_stride', C.c_longlong), ('n_valid', C.c_longlong), ('out_round_tf32', C.c_int),
        ('scale', C.c_void_p), ('shift', C.c_void_p), ('relu', C.c_int),
        ('res', C.c_void_p), ('res_row_stride', C.c_longlong), ('res_seq_stride', C.c_longlong),
        ('res_row_mul', C.c_int), ('res_row_off', C.c_int),
        ('res_rows', C.c_longlong), ('res_col_off', C.c_longlong), ('res_cols', C.c_longlong),
        ('dyn_offsets', C.c_void_p), ('out_rows_total', C.c_longlong),
        ('stat_sum', C.c_void_p), ('stat_sqsum', C.c_void_p),
        ('fin', C.c_void_p),
        ('drop', C.c_void_p), ('side', C.c_void_p), ('side_mode', C.c_int), ('side_row_off', C.c_int),
        ('side_row_stride', C.c_longlong), ('side_seq_stride', C.c_longlong), ('side_rows', C.c_longlong),
        ('side_scale', C.c_float),
    ]


class BnFin(C.Structure):
    """struct vp3d_bn_fin (include/vp3d_b200.h)"""
    _fields_ = [('count', C.c_longlong), ('gamma', C.c_void_p), ('beta', C.c_void_p), ('eps', C.c_float),
                ('momentum', C.c_float), ('running_mean', C.c_void_p), ('running_var', C.c_void_p),
                ('num_batches_tracked', C.c_void_p), ('scale', C.c_void_p), ('shift', C.c_void_p), ('mean', C.c_void_p),
                ('invstd', C.c_void_p), ('c', C.c_int), ('done_counter', C.c_void_p)]


class WgradArgs(C.Structure):
    """struct vp3d_wgrad_args (include/vp3d_b200.h)"""
    _fields_ = [
        ('dtype', C.c_int), ('block_n', C.c_int),
        ('dz', C.c_void_p), ('dz_seqs', C.c_longlong), ('dz_rows', C.c_longlong), ('dz_row_stride', C.c_longlong),
        ('dz_seq_stride', C.c_longlong), ('co_pad', C.c_longlong),
        ('a', C.c_void_p), ('a_rows', C.c_longlong), ('a_cols', C.c_longlong), ('a_row_stride', C.c_longlong),
        ('a_seq_stride', C.c_longlong), ('ci_pad', C.c_longlong),
        ('taps', C.c_int), ('b_row_off', C.c_longlong), ('b_tap_row_step', C.c_int), ('b_tap_col_step', C.c_longlong),
        ('dw_packed', C.c_void_p), ('dz_cols', C.c_longlong), ('max_slices', C.c_int),
    ]


class WindowArgs(C.Structure):
    """struct vp3d_window_args (include/vp3d_b200.h)"""
    _fields_ = [('x_world', C.c_void_p), ('q', C.c_void_p), ('t', C.c_void_p), ('cam', C.c_void_p),
                ('seq_start', C.c_void_p), ('seq_len', C.c_void_p), ('sample_seq', C.c_void_p),
                ('sample_start', C.c_void_p),
                ('batch', C.c_int), ('joints', C.c_int), ('chunk_length', C.c_int), ('pad', C.c_int),
                ('causal_shift', C.c_int), ('root_relative', C.c_int), ('linear', C.c_int),
                ('out2', C.c_void_p), ('target3', C.c_void_p), ('cam3x4', C.c_void_p)]


class AdamArgs(C.Structure):
    """struct vp3d_adam_args (include/vp3d_b200.h)"""
    _fields_ = [('p', C.c_void_p), ('g', C.c_void_p), ('m', C.c_void_p), ('v', C.c_void_p), ('vmax', C.c_void_p),
                ('n', C.c_longlong),
                ('lr', C.c_float), ('beta1', C.c_float), ('beta2', C.c_float), ('eps', C.c_float),
                ('weight_decay', C.c_float),
                ('step', C.c_void_p), ('lr_dev', C.c_void_p), ('maximize', C.c_int),
                ('packed', C.c_void_p), ('dtype', C.c_int), ('c_in', C.c_int), ('taps', C.c_int), ('k_pad', C.c_int)]


class AllReduceArgs(C.Structure):
    """struct vp3d_allreduce_args (include/vp3d_b200.h)"""
    _fields_ = [('multicast', C.c_void_p), ('peers', C.POINTER(C.c_void_p)), ('flags', C.POINTER(C.c_void_p)),
                ('rank', C.c_int), ('world', C.c_int), ('offset', C.c_longlong), ('count', C.c_longlong),
                ('scale', C.c_float), ('ctas', C.c_int), ('timeout_s', C.c_double)]


class StreamLayer(C.Structure):
    """struct vp3d_stream_layer (include/vp3d_b200.h)"""
    _fields_ = [('a', C.c_void_p), ('w', C.c_void_p), ('shift', C.c_void_p), ('res', C.c_void_p), ('out', C.c_void_p),
                ('a_ring', C.c_int), ('res_ring', C.c_int), ('out_ring', C.c_int),
                ('k_per_tap', C.c_int), ('taps', C.c_int), ('tap_row_step', C.c_int),
                ('n', C.c_int), ('n_valid', C.c_int), ('relu', C.c_int), ('out_f32', C.c_int),
                ('res_row_stride', C.c_int), ('out_row_stride', C.c_int)]


class Dropout(C.Structure):
    """struct vp3d_dropout (include/vp3d_b200.h)"""
    _fields_ = [('p', C.c_float), ('seed', C.c_ulonglong), ('stream', C.c_ulonglong), ('step_counter', C.c_void_p)]


_SIGNATURES = {
    'vp3d_version': (C.c_int, []),
    'vp3d_last_error': (C.c_char_p, []),
    'vp3d_device_info': (C.c_int, [C.POINTER(C.c_int)] * 3),
    'vp3d_set_sm_limit': (C.c_int, [C.c_int]),
    'vp3d_set_pair_mode': (C.c_int, [C.c_int]),
    'vp3d_set_pdl': (C.c_int, [C.c_int]),
    'vp3d_set_sched_mode': (C.c_int, [C.c_int]),
    'vp3d_conv_block_fwd': (C.c_int, [C.POINTER(ConvArgs), C.c_void_p]),
    'vp3d_pack_rows': (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    'vp3d_pack_rows_ones': (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p]),
    'vp3d_expand_bn_stats': (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_float, C.c_float] + [C.c_void_p] * 8 + [C.c_int, C.c_int, C.c_void_p]),
    'vp3d_expand_bwd_finish': (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int] +
                               [C.c_void_p] * 4 + [C.c_int] * 5 + [C.c_void_p] * 4),
    'vp3d_pack_conv_weight': (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_void_p]),
    'vp3d_pack_conv_weight_scaled': (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                               C.c_int, C.c_int, C.c_void_p]),
    'vp3d_bn_fold': (C.c_int, [C.c_void_p] * 4 + [C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    'vp3d_project_points': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_void_p]),
    'vp3d_project_bwd': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_void_p,
                                   C.c_void_p]),
    'vp3d_project_windows': (C.c_int, [C.POINTER(WindowArgs), C.c_void_p]),
    'vp3d_loss_workspace_bytes': (C.c_longlong, []),
    'vp3d_mpjpe_fwd': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p] + [C.c_longlong] * 5 +
                       [C.c_void_p, C.c_void_p, C.c_void_p]),
    'vp3d_mpjpe_bwd': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p] + [C.c_longlong] * 5 +
                       [C.c_void_p, C.c_void_p]),
    'vp3d_mpjpe_nd_fwd': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p] + [C.c_longlong] * 5 +
                          [C.c_void_p, C.c_void_p, C.c_void_p]),
    'vp3d_mpjpe_nd_bwd': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p] +
                          [C.c_longlong] * 5 + [C.c_void_p, C.c_void_p]),
    'vp3d_n_mpjpe_fwd': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    'vp3d_p_mpjpe_fwd': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    'vp3d_velocity_error': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    'vp3d_n_mpjpe_bwd': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]),
    'vp3d_reproj_mpjpe_fwd': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_longlong,
                                        C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'vp3d_reproj_mpjpe_bwd': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_longlong,
                                        C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'vp3d_wgrad': (C.c_int, [C.POINTER(WgradArgs), C.c_void_p]),
    'vp3d_wgrad_finish': (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 3 + [C.c_longlong] * 2 + [C.c_void_p, C.c_void_p]),
    'vp3d_bn_finalize': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_float, C.c_float] +
                         [C.c_void_p] * 7 + [C.c_int, C.c_int, C.c_void_p]),
    'vp3d_bn_act_fwd': (C.c_int, [C.c_int] + [C.c_void_p] * 4 + [C.c_longlong] * 3 + [C.c_int] * 3 +
                        [C.POINTER(Dropout), C.c_void_p, C.c_void_p]),
    'vp3d_bn_act_fwd_mask': (C.c_int, [C.c_int] + [C.c_void_p] * 4 + [C.c_longlong] * 3 + [C.c_int] * 3 +
                             [C.POINTER(Dropout), C.c_void_p, C.c_void_p, C.c_void_p]),
    'vp3d_bn_act_bwd_reduce_mask': (C.c_int, [C.c_int] + [C.c_void_p] * 5 + [C.c_float, C.c_longlong, C.c_int] +
                                    [C.c_void_p] * 3),
    'vp3d_bn_act_bwd_apply_mask': (C.c_int, [C.c_int] + [C.c_void_p] * 6 + [C.c_float, C.c_longlong, C.c_longlong, C.c_int,
                                                                           C.c_int] + [C.c_void_p] * 7),
    'vp3d_bn_finalize_act_fwd': (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p,
                                           C.c_float, C.c_float] + [C.c_void_p] * 7 + [C.c_int, C.c_void_p] +
                                 [C.c_longlong] * 3 + [C.c_int] * 3 + [C.POINTER(Dropout), C.c_void_p, C.c_void_p]),
    'vp3d_col_stats': (C.c_int, [C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    'vp3d_bn_act_bwd_reduce': (C.c_int, [C.c_int] + [C.c_void_p] * 6 + [C.c_longlong, C.c_int, C.POINTER(Dropout),
                                                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    'vp3d_bn_act_bwd_apply': (C.c_int, [C.c_int] + [C.c_void_p] * 6 + [C.c_longlong, C.c_longlong, C.c_int, C.c_int,
                                                                      C.POINTER(Dropout)] + [C.c_void_p] * 7),
    'vp3d_stream_advance': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    'vp3d_stream_step_fused': (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                         C.POINTER(StreamLayer), C.c_int, C.c_void_p]),
    'vp3d_ring_write': (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int,
                                  C.c_void_p]),
    'vp3d_adam_step': (C.c_int, [C.POINTER(AdamArgs), C.c_void_p]),
    'vp3d_adam_step_multi': (C.c_int, [C.POINTER(AdamArgs), C.c_int, C.c_void_p]),
    'vp3d_peer_allreduce_f32': (C.c_int, [C.POINTER(AllReduceArgs), C.c_void_p]),
    'vp3d_counter_add': (C.c_int, [C.c_void_p, C.c_ulonglong, C.c_void_p]),
    'vp3d_grad_scale': (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]),
    'vp3d_grad_pack_rows': (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
}

_lib = None
_lock = threading.Lock()


def declared_symbols():
    return sorted(_SIGNATURES)


def lib():
    """The loaded library. Raises RuntimeError (loudly, no fallback) when it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        'vp3d_b200: %s not found -- build it with `python -c "import __graft_entry__ as g; g.build()"` '
                        'or `make -C dynamic-camera-augmented-videopose3d_b200/csrc`. There is no CPU fallback.'
                        % LIB_PATH)
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in _SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def check(rc, what=''):
    if rc != 0:
        msg = lib().vp3d_last_error().decode('utf-8', 'replace')
        if rc == 1:
            raise AssertionError('vp3d_b200 %s: %s' % (what, msg))  # the reference signals bad shapes with assert
        raise RuntimeError('vp3d_b200 %s failed (status %d): %s' % (what, rc, msg))


def device_info():
    sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
    check(lib().vp3d_device_info(C.byref(sm), C.byref(maj), C.byref(mnr)), 'device_info')
    return sm.value, maj.value, mnr.value


_sm_counts = {}


def sm_count(device=None):
    """SMs of `device` (a torch device / index; default: the current CUDA device), asked from the library once per
    device. Host-side tile heuristics size against this, never against a literal."""
    import torch
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    n = _sm_counts.get(idx)
    if n is None:
        with torch.cuda.device(idx):
            n = _sm_counts[idx] = device_info()[0]
    return n
