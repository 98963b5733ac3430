"""torch-facing wrappers over the C ABI: tensors in, tensors out, raw pointers + current stream underneath.

PyTorch is plumbing here (device memory from the caching allocator, the current CUDA stream, autograd graph
bookkeeping); every arithmetic operation on the hot path runs in libvp3d_b200.so.
"""
import ctypes as C
import math

import torch

from . import native
from .native import BnFin, ConvArgs, Dropout, WgradArgs, check, lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('vp3d_b200 runs on CUDA (sm_100a) tensors only; there is no CPU fallback '
                               '(got a %s tensor)' % t.device)


def f32c(t):
    """fp32 contiguous view/copy (layout plumbing only)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def torch_dtype(dt):
    return {native.F16: torch.float16, native.BF16: torch.bfloat16, native.TF32: torch.float32}[dt]


# ------------------------------------------------------------------------------------------------ K5 geometry
def project_points(x, q=None, t=None, cam=None, pts_per_q=1, pts_per_cam=1, mode=0, want3=False, want2=False):
    """x (..., 3) fp32 CUDA -> (out3 or None, out2 or None). See vp3d_project_points in include/vp3d_b200.h."""
    require_cuda(x, q, t, cam)
    x = f32c(x)
    n_pts = x.numel() // 3
    out3 = torch.empty_like(x) if want3 else None
    out2 = torch.empty(x.shape[:-1] + (2,), dtype=torch.float32, device=x.device) if want2 else None
    q = None if q is None else f32c(q)
    t = None if t is None else f32c(t)
    cam = None if cam is None else f32c(cam)
    with torch.cuda.device(x.device):
        check(lib().vp3d_project_points(_ptr(x), _ptr(out3), _ptr(out2), n_pts, _ptr(q), _ptr(t), _ptr(cam),
                                        int(pts_per_q), int(pts_per_cam), int(mode), _stream()), 'project_points')
    return out3, out2


class _ProjectFn(torch.autograd.Function):
    """project_to_2d / project_to_2d_linear with the gradient wrt the camera-space points (vp3d_project_bwd)."""

    @staticmethod
    def forward(ctx, x, cam, per_cam, linear):
        x, cam = f32c(x), f32c(cam)
        mode = native.PT_PROJECT | (native.PT_LINEAR if linear else 0)
        _, out = project_points(x.detach(), cam=cam.detach(), pts_per_cam=per_cam, mode=mode, want2=True)
        ctx.save_for_backward(x, cam)
        ctx.meta = (per_cam, linear)
        return out

    @staticmethod
    def backward(ctx, g):
        x, cam = ctx.saved_tensors
        per_cam, linear = ctx.meta
        if not ctx.needs_input_grad[0]:
            return None, None, None, None
        g = f32c(g)
        gx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            check(lib().vp3d_project_bwd(_ptr(x), _ptr(cam), _ptr(g), x.numel() // 3, int(per_cam), 1 if linear else 0,
                                         _ptr(gx), _stream()), 'project_bwd')
        return gx, None, None, None


def project_2d(x, cam, per_cam, linear=False):
    require_cuda(x, cam)
    return _ProjectFn.apply(x, cam, per_cam, linear)


# ------------------------------------------------------------------------------------------------ K6 losses
_ws_cache = {}


def _loss_workspace(device):
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None:
        ws = torch.empty(int(lib().vp3d_loss_workspace_bytes()), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def _weight_view(w, lead):
    """w broadcast over the joint grid `lead` = pred.shape[:-1] (loss.py:27 `w * norm`), described to the kernel as
    element strides over an (N, T, J) grid; broadcast dimensions get stride 0, nothing is materialised for rank 3."""
    wb = torch.broadcast_to(f32c(w), lead)
    if len(lead) == 3:
        s = wb.stride()
        return wb, lead[1], lead[2], s[0], s[1], s[2]
    wb = wb.contiguous().reshape(-1)
    return wb, 1, 1, 1, 0, 0


class _MpjpeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, tgt, w):
        require_cuda(pred, tgt, w)
        assert pred.shape == tgt.shape  # loss.py:16
        D = int(pred.shape[-1])           # norm over the last axis, whatever its length (loss.py:17)
        p, g = f32c(pred), f32c(tgt)
        n_joints = p.numel() // D
        if w is not None:
            assert w.shape[0] == pred.shape[0]  # loss.py:26
            wt, T, J, sn, st, sj = _weight_view(w, tuple(pred.shape[:-1]))
        else:
            wt, T, J, sn, st, sj = None, 1, 1, 0, 0, 0
        out = torch.empty((), dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device):
            if D == 3:
                check(lib().vp3d_mpjpe_fwd(_ptr(p), _ptr(g), n_joints, _ptr(wt), T, J, sn, st, sj,
                                           _ptr(_loss_workspace(p.device)), _ptr(out), _stream()), 'mpjpe_fwd')
            else:
                check(lib().vp3d_mpjpe_nd_fwd(_ptr(p), _ptr(g), n_joints, D, _ptr(wt), T, J, sn, st, sj,
                                              _ptr(_loss_workspace(p.device)), _ptr(out), _stream()), 'mpjpe_nd_fwd')
        ctx.save_for_backward(p, g, wt if wt is not None else torch.empty(0, device=p.device))
        ctx.meta = (n_joints, wt is not None, T, J, sn, st, sj, pred.shape, pred.dtype, D)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        p, g, wt = ctx.saved_tensors
        n_joints, has_w, T, J, sn, st, sj, shape, dtype, D = ctx.meta

        def grad_wrt_pred():
            go = f32c(grad_out).reshape(1)
            gp = torch.empty_like(p)
            with torch.cuda.device(p.device):
                if D == 3:
                    check(lib().vp3d_mpjpe_bwd(_ptr(p), _ptr(g), _ptr(go), n_joints, _ptr(wt) if has_w else None, T, J,
                                               sn, st, sj, _ptr(gp), _stream()), 'mpjpe_bwd')
                else:
                    check(lib().vp3d_mpjpe_nd_bwd(_ptr(p), _ptr(g), _ptr(go), n_joints, D, _ptr(wt) if has_w else None,
                                                  T, J, sn, st, sj, _ptr(gp), _stream()), 'mpjpe_nd_bwd')
            return gp.reshape(shape).to(dtype)

        grad_pred = grad_wrt_pred() if ctx.needs_input_grad[0] else None
        grad_tgt = None
        if ctx.needs_input_grad[1]:
            grad_tgt = -(grad_pred if grad_pred is not None else grad_wrt_pred())
        return grad_pred, grad_tgt, None


def mpjpe(pred, tgt, w=None):
    return _MpjpeFn.apply(pred, tgt, w)


class _ReprojFn(torch.autograd.Function):
    """mpjpe(project_to_2d(pose + traj, cam), target_2d) in one kernel each way (vp3d_reproj_mpjpe_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, pose, traj, cam, tgt2, linear):
        p, c, t2 = f32c(pose), f32c(cam), f32c(tgt2)
        tr = None if traj is None else f32c(traj)
        n_pts = p.numel() // 3
        per_cam = max(n_pts // max(c.shape[0], 1), 1)
        per_traj = 0 if tr is None else n_pts // (tr.numel() // 3)
        out = torch.empty((), dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device):
            check(lib().vp3d_reproj_mpjpe_fwd(_ptr(p), _ptr(tr), n_pts, per_traj, _ptr(c), per_cam, 1 if linear else 0,
                                              _ptr(t2), _ptr(_loss_workspace(p.device)), _ptr(out), _stream()),
                  'reproj_mpjpe_fwd')
        ctx.save_for_backward(p, tr if tr is not None else torch.empty(0, device=p.device), c, t2)
        ctx.meta = (n_pts, per_traj, per_cam, linear, tr is not None, pose.shape, None if traj is None else traj.shape)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        p, tr, c, t2 = ctx.saved_tensors
        n_pts, per_traj, per_cam, linear, has_traj, pshape, tshape = ctx.meta
        want_p = ctx.needs_input_grad[0]
        want_t = has_traj and ctx.needs_input_grad[1]
        if not (want_p or want_t):
            return None, None, None, None, None
        gp = torch.empty_like(p) if want_p else None
        gt = torch.zeros_like(tr) if want_t else None
        with torch.cuda.device(p.device):
            check(lib().vp3d_reproj_mpjpe_bwd(_ptr(p), _ptr(tr) if has_traj else None, n_pts, per_traj, _ptr(c), per_cam,
                                              1 if linear else 0, _ptr(t2), _ptr(f32c(grad_out).reshape(1)), _ptr(gp),
                                              _ptr(gt), _stream()), 'reproj_mpjpe_bwd')
        return (gp.reshape(pshape) if want_p else None, gt.reshape(tshape) if want_t else None, None, None, None)


def reproj_mpjpe(pose, camera_params, target_2d, traj=None, linear=False):
    """Reprojection loss mpjpe(project_to_2d(pose + traj, camera_params), target_2d) (camera.py:37-67 + loss.py:11-17)
    fused into one kernel per direction. pose (N, ..., J, 3) camera-space points, traj (N, ..., 1, 3) or None (added to
    every joint of its pose), camera_params (N, 9), target_2d (N, ..., J, 2). Differentiable wrt pose and traj."""
    require_cuda(pose, camera_params, target_2d, traj)
    assert pose.shape[-1] == 3 and target_2d.shape[-1] == 2 and tuple(pose.shape[:-1]) == tuple(target_2d.shape[:-1])
    assert camera_params.dim() == 2 and camera_params.shape[-1] == 9 and camera_params.shape[0] == pose.shape[0]
    if traj is not None:
        assert traj.shape[-1] == 3 and traj.shape[-2] == 1 and tuple(traj.shape[:-2]) == tuple(pose.shape[:-2]), \
            'traj must be (..., 1, 3) with the leading dimensions of pose'
    return _ReprojFn.apply(pose, traj, camera_params, target_2d, linear)


class _NMpjpeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, tgt):
        p, g = f32c(pred), f32c(tgt)
        J = p.shape[2]
        n_poses = p.shape[0] * p.shape[1]
        out = torch.empty((), dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device):
            check(lib().vp3d_n_mpjpe_fwd(_ptr(p), _ptr(g), n_poses, J, _ptr(_loss_workspace(p.device)), _ptr(out),
                                         _stream()), 'n_mpjpe_fwd')
        ctx.save_for_backward(p, g)
        ctx.meta = (n_poses, J, pred.shape, pred.dtype)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        p, g = ctx.saved_tensors
        n_poses, J, shape, dtype = ctx.meta
        if ctx.needs_input_grad[1]:
            raise RuntimeError('vp3d_b200 n_mpjpe: gradient wrt the target is not implemented')
        gp = torch.empty_like(p)
        with torch.cuda.device(p.device):
            check(lib().vp3d_n_mpjpe_bwd(_ptr(p), _ptr(g), _ptr(f32c(grad_out).reshape(1)), n_poses, J, _ptr(gp),
                                         _stream()), 'n_mpjpe_bwd')
        return gp.reshape(shape).to(dtype), None


def n_mpjpe(pred, tgt):
    require_cuda(pred, tgt)
    assert pred.shape == tgt.shape  # loss.py:75
    assert pred.dim() == 4 and pred.shape[-1] == 3, 'n_mpjpe expects (N, T, J, 3) (loss.py:77 reduces dims 3 and 2)'
    return _NMpjpeFn.apply(pred, tgt)


def p_mpjpe(pred, tgt):
    """(n_poses, J, 3) CUDA tensors -> 0-dim tensor: Procrustes-aligned MPJPE (loss.py:29-68) without leaving the device."""
    require_cuda(pred, tgt)
    assert pred.shape == tgt.shape
    assert pred.dim() == 3 and pred.shape[-1] == 3, 'p_mpjpe expects (n_poses, J, 3) (loss.py:36 averages over axis 1)'
    p, g = f32c(pred.detach()), f32c(tgt.detach())
    out = torch.empty((), dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        check(lib().vp3d_p_mpjpe_fwd(_ptr(p), _ptr(g), p.shape[0], p.shape[1], _ptr(_loss_workspace(p.device)),
                                     _ptr(out), _stream()), 'p_mpjpe_fwd')
    return out


def mean_velocity_error(pred, tgt):
    """(T, ..., D) CUDA tensors -> 0-dim tensor (loss.py:82-91: first differences along axis 0)."""
    require_cuda(pred, tgt)
    assert pred.shape == tgt.shape
    assert pred.dim() >= 2 and pred.shape[0] >= 2
    p, g = f32c(pred.detach()), f32c(tgt.detach())
    T, D = p.shape[0], p.shape[-1]
    inner = p.numel() // (T * D)
    out = torch.empty((), dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        check(lib().vp3d_velocity_error(_ptr(p), _ptr(g), T, inner, D, _ptr(_loss_workspace(p.device)), _ptr(out),
                                        _stream()), 'velocity_error')
    return out


# ------------------------------------------------------------------------------------------------ K1 plumbing
def pack_rows(dt, src, c_pad, ones_col=None):
    """fp32 (rows, c) -> operand type (rows, c_pad) zero padded; `ones_col`: that padding column holds 1.0 instead
    (vp3d_pack_rows_ones: the Gram matrix of the rows then carries their column sums and count)."""
    src = f32c(src)
    rows, c = src.shape
    dst = torch.empty((rows, c_pad), dtype=torch_dtype(dt), device=src.device)
    with torch.cuda.device(src.device):
        if ones_col is None:
            check(lib().vp3d_pack_rows(dt, _ptr(src), _ptr(dst), rows, c, c_pad, _stream()), 'pack_rows')
        else:
            check(lib().vp3d_pack_rows_ones(dt, _ptr(src), _ptr(dst), rows, c, c_pad, int(ones_col), _stream()),
                  'pack_rows_ones')
    return dst


def pack_conv_weight(dt, w, rows_pad, k_pad_per_tap, transpose=0):
    """transpose: 0 forward operand, 1 / 2 data-gradient operands (see vp3d_pack_conv_weight)."""
    w = f32c(w.detach())
    c_out, c_in, taps = w.shape
    transpose = int(transpose)
    k_total = k_pad_per_tap if transpose == 1 else taps * k_pad_per_tap
    dst = torch.empty((rows_pad, k_total), dtype=torch_dtype(dt), device=w.device)
    with torch.cuda.device(w.device):
        check(lib().vp3d_pack_conv_weight(dt, _ptr(w), _ptr(dst), c_out, c_in, taps, rows_pad, k_pad_per_tap,
                                          transpose, _stream()), 'pack_conv_weight')
    return dst


def pack_conv_weight_scaled(dt, w, row_scale, rows_pad, k_pad_per_tap):
    """Forward operand with every output-channel row multiplied by row_scale[n] (folded BatchNorm scale)."""
    w = f32c(w.detach())
    c_out, c_in, taps = w.shape
    dst = torch.empty((rows_pad, taps * k_pad_per_tap), dtype=torch_dtype(dt), device=w.device)
    with torch.cuda.device(w.device):
        check(lib().vp3d_pack_conv_weight_scaled(dt, _ptr(w), _ptr(f32c(row_scale)), _ptr(dst), c_out, c_in, taps,
                                                 rows_pad, k_pad_per_tap, _stream()), 'pack_conv_weight_scaled')
    return dst


def bn_fold(bn, c_pad):
    """nn.BatchNorm1d container (eval statistics) -> (scale, shift) fp32 [c_pad]."""
    c = bn.num_features
    dev = bn.weight.device
    scale = torch.empty(c_pad, dtype=torch.float32, device=dev)
    shift = torch.empty(c_pad, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib().vp3d_bn_fold(_ptr(f32c(bn.weight.detach())), _ptr(f32c(bn.bias.detach())),
                                 _ptr(f32c(bn.running_mean)), _ptr(f32c(bn.running_var)), float(bn.eps),
                                 _ptr(scale), _ptr(shift), c, c_pad, _stream()), 'bn_fold')
    return scale, shift


def conv_block(dt, a, a_view, w, taps, tap_row_step, k_per_tap, rows_out, out, out_view, block_n=256, a_row_off=0,
               scale=None, shift=None, relu=False, res=None, res_view=None, out_f32=False, n_valid=None,
               stat_sum=None, stat_sqsum=None, out_round_tf32=False, res_rows=0, res_col_off=0, res_cols=0,
               w_mn_major=None, dyn_offsets=None, out_rows_total=0, drop=None, side=None, side_view=None, side_mode=0,
               side_scale=1.0, fin=None):
    """One vp3d_conv_block_fwd launch.
    a_view   = (seqs, rows, kdim, row_stride, seq_stride)    out_view = (row_stride, seq_stride)
    res_view = (row_stride, seq_stride, row_mul, row_off)    side_view = (row_stride, seq_stride, rows, row_off)
    drop (a native.Dropout) / side: the fused epilogue of the CTA-pair kernel (see vp3d_conv_args)."""
    args = ConvArgs()
    args.dtype, args.block_n = dt, block_n
    args.a = a.data_ptr()
    args.a_seqs, args.a_rows, args.a_kdim, args.a_row_stride, args.a_seq_stride = a_view
    args.a_row_off = a_row_off
    args.w = w.data_ptr()
    args.taps, args.tap_row_step, args.k_per_tap = taps, tap_row_step, k_per_tap
    if w_mn_major is None:
        args.n_pad, args.k_total = w.shape[0], w.shape[1]
    else:
        # w is the forward-packed [k rows][row_stride columns] matrix read as W^T: (n_pad, tap column step)
        args.n_pad, args.w_tap_col_step = w_mn_major
        args.w_mn_major, args.w_row_stride, args.k_total = 1, w.shape[1], 0
    args.rows_out = rows_out
    args.out = out.data_ptr()
    args.out_f32 = 1 if out_f32 else 0
    args.out_row_stride, args.out_seq_stride = out_view
    args.n_valid = n_valid if n_valid is not None else args.n_pad
    args.out_round_tf32 = 1 if out_round_tf32 else 0
    args.scale = None if scale is None else scale.data_ptr()
    args.shift = None if shift is None else shift.data_ptr()
    args.relu = 1 if relu else 0
    if res is not None:
        args.res = res.data_ptr()
        args.res_row_stride, args.res_seq_stride, args.res_row_mul, args.res_row_off = res_view
        args.res_rows, args.res_col_off, args.res_cols = res_rows, res_col_off, res_cols
    if dyn_offsets is not None:       # device int[4] (a view into the streaming offsets table)
        args.dyn_offsets, args.out_rows_total = dyn_offsets.data_ptr(), out_rows_total
    args.stat_sum = None if stat_sum is None else stat_sum.data_ptr()
    args.stat_sqsum = None if stat_sqsum is None else stat_sqsum.data_ptr()
    if fin is not None:           # a native.BnFin (make_bn_fin): vp3d_bn_finalize in the tail of this launch
        args.fin = C.addressof(fin)
    if drop is not None and drop.p > 0:
        args.drop = C.addressof(drop)
    if side is not None:
        args.side, args.side_mode, args.side_scale = side.data_ptr(), int(side_mode), float(side_scale)
        args.side_row_stride, args.side_seq_stride, args.side_rows, args.side_row_off = side_view
    with torch.cuda.device(a.device):
        check(lib().vp3d_conv_block_fwd(C.byref(args), _stream()), 'conv_block_fwd')
    return out


# ------------------------------------------------------------------------------------------------ training path
def make_dropout(p, seed, stream, step_counter=None):
    """step_counter: optional int64 device tensor (one element) the kernels read as the training-step number."""
    d = Dropout()
    d.p, d.seed, d.stream = float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream) & 0xFFFFFFFFFFFFFFFF
    d.step_counter = None if step_counter is None else step_counter.data_ptr()
    # the struct carries a raw device pointer; the backward recomputes the forward's masks from it long after the caller's
    # local tensor has gone out of scope, so the tensor lives as long as the description does
    d._keepalive = step_counter
    return d


def counter_add(counter, inc=1):
    with torch.cuda.device(counter.device):
        check(lib().vp3d_counter_add(_ptr(counter), int(inc), _stream()), 'counter_add')


def wgrad(dt, dz, dz_view, a, a_view, co_pad, ci_pad, taps, dw_packed, b_row_off=0, b_tap_row_step=0,
          b_tap_col_step=0, block_n=256, dz_cols=0, max_slices=0):
    """One vp3d_wgrad launch. dz_view = (seqs, rows, row_stride, seq_stride); a_view = (rows, cols, row_stride,
    seq_stride). dw_packed: zero-filled fp32 [taps][co_pad][ci_pad]."""
    args = WgradArgs()
    args.dtype, args.block_n = dt, block_n
    args.dz = dz.data_ptr()
    args.dz_seqs, args.dz_rows, args.dz_row_stride, args.dz_seq_stride = dz_view
    args.co_pad = co_pad
    args.a = a.data_ptr()
    args.a_rows, args.a_cols, args.a_row_stride, args.a_seq_stride = a_view
    args.ci_pad = ci_pad
    args.taps, args.b_row_off, args.b_tap_row_step, args.b_tap_col_step = taps, b_row_off, b_tap_row_step, b_tap_col_step
    args.dw_packed = dw_packed.data_ptr()
    args.dz_cols = dz_cols
    args.max_slices = max_slices
    with torch.cuda.device(dz.device):
        check(lib().vp3d_wgrad(C.byref(args), _stream()), 'wgrad')
    return dw_packed


def _grad_out(out, shape, device):
    """`out` (a caller-owned fp32 buffer of that many elements, e.g. a slot of the data-parallel exchange buffer) viewed
    as `shape`, or a fresh tensor."""
    if out is None:
        return torch.empty(shape, dtype=torch.float32, device=device)
    assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == math.prod(shape), (out.shape, shape)
    return out.view(shape)


def wgrad_finish(dw_packed, c_out, c_in, taps, co_pad, ci_pad, gscale_buf, tap_stride=None, row_stride=None, out=None):
    """packed [taps][co_pad][ci_pad] (or the strides given) -> nn.Conv1d layout (c_out, c_in, taps), un-scaled."""
    dw = _grad_out(out, (c_out, c_in, taps), dw_packed.device)
    tap_stride = co_pad * ci_pad if tap_stride is None else tap_stride
    row_stride = ci_pad if row_stride is None else row_stride
    with torch.cuda.device(dw.device):
        check(lib().vp3d_wgrad_finish(_ptr(dw_packed), _ptr(dw), c_out, c_in, taps, tap_stride, row_stride,
                                      _ptr(gscale_buf), _stream()), 'wgrad_finish')
    return dw


def col_stats(dt, z, stats):
    """stats (double [2][c_pad]) += per-channel sum / sum of squares of the stored matrix z [rows][c_pad]."""
    c_pad = z.shape[-1]
    rows = z.numel() // c_pad
    with torch.cuda.device(z.device):
        check(lib().vp3d_col_stats(dt, _ptr(z), rows, c_pad, _ptr(stats[0]), _ptr(stats[1]), _stream()), 'col_stats')
    return stats


def _bn_momentum(bn):
    """nn.BatchNorm1d.momentum for the kernels: None (cumulative moving average, factor 1 / num_batches_tracked) is
    passed as a negative value and resolved on the device from the counter."""
    return -1.0 if bn.momentum is None else float(bn.momentum)


def _bump_running_stats(bn):
    """The kernels write running_mean / running_var / num_batches_tracked through raw pointers; tell autograd's version
    counters, which key the folded eval-mode cache (temporal.packed_for)."""
    for b in (bn.running_mean, bn.running_var, bn.num_batches_tracked):
        if b is not None:
            torch.autograd.graph.increment_version(b)


def make_bn_fin(bn, count, c_pad, done_counter, update_running=True):
    """-> (native.BnFin for vp3d_conv_args.fin, (scale, shift, mean, invstd) fp32 [c_pad] it will fill).
    `done_counter`: a zeroed 4-byte device word (a view into the caller's zero-filled arena)."""
    dev = bn.weight.device
    out = torch.empty((4, c_pad), dtype=torch.float32, device=dev)
    track = update_running and bn.track_running_stats and bn.running_mean is not None
    f = BnFin()
    f.count = int(count)
    gamma, beta = f32c(bn.weight.detach()), f32c(bn.bias.detach())
    f.gamma, f.beta, f.eps, f.momentum = gamma.data_ptr(), beta.data_ptr(), float(bn.eps), _bn_momentum(bn)
    if track:
        f.running_mean, f.running_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
        f.num_batches_tracked = bn.num_batches_tracked.data_ptr()
    f.scale, f.shift, f.mean, f.invstd = (out[i].data_ptr() for i in range(4))
    f.c, f.done_counter = bn.num_features, done_counter.data_ptr()
    f._keepalive = (gamma, beta, out, done_counter)
    if track:
        _bump_running_stats(bn)
    return f, (out[0], out[1], out[2], out[3])


def bn_finalize(stat, count, bn, c_pad, update_running=True):
    """stat: double [2][c_pad] (sum, sum of squares) -> (scale, shift, mean, invstd) fp32 [c_pad]; updates the
    nn.BatchNorm1d container's running statistics in place like F.batch_norm(training=True)."""
    c = bn.num_features
    dev = stat.device
    out = torch.empty((4, c_pad), dtype=torch.float32, device=dev)
    track = update_running and bn.track_running_stats and bn.running_mean is not None
    with torch.cuda.device(dev):
        check(lib().vp3d_bn_finalize(_ptr(stat[0]), _ptr(stat[1]), int(count), _ptr(f32c(bn.weight.detach())),
                                     _ptr(f32c(bn.bias.detach())), float(bn.eps), _bn_momentum(bn),
                                     _ptr(bn.running_mean) if track else None, _ptr(bn.running_var) if track else None,
                                     _ptr(bn.num_batches_tracked) if track else None,
                                     _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(out[3]), c, c_pad, _stream()),
              'bn_finalize')
    if track:
        _bump_running_stats(bn)
    return out[0], out[1], out[2], out[3]


def keep_scale(p):
    """Multiplier of kept elements: the kernels quantise the drop probability to 1/256 (see dropout.cuh)."""
    thresh = int(float(p) * 256.0 + 0.5)
    return 256.0 / (256.0 - thresh) if thresh > 0 else 1.0


def expand_bn_stats(dt, gram, w, k_total, ones_col, bn, c_pad, update_running=True):
    """Train-mode BatchNorm of the expand layer from the Gram matrix of its input (vp3d_expand_bn_stats)
    -> (scale, shift, mean, invstd [c_pad], wg [c_pad][256])."""
    dev = gram.device
    out = torch.empty((4, c_pad), dtype=torch.float32, device=dev)
    wg = torch.empty((c_pad, 256), dtype=torch.float32, device=dev)
    track = update_running and bn.track_running_stats and bn.running_mean is not None
    with torch.cuda.device(dev):
        check(lib().vp3d_expand_bn_stats(dt, _ptr(gram), _ptr(w), int(k_total), int(ones_col),
                                         _ptr(f32c(bn.weight.detach())), _ptr(f32c(bn.bias.detach())), float(bn.eps),
                                         _bn_momentum(bn), _ptr(bn.running_mean) if track else None,
                                         _ptr(bn.running_var) if track else None,
                                         _ptr(bn.num_batches_tracked) if track else None, _ptr(out[0]), _ptr(out[1]),
                                         _ptr(out[2]), _ptr(out[3]), _ptr(wg), bn.num_features, c_pad, _stream()),
              'expand_bn_stats')
    if track:
        _bump_running_stats(bn)
    return out[0], out[1], out[2], out[3], wg


def expand_bwd_finish(dt, p_packed, wg, gram, w, k_total, ones_col, scale, mean, invstd, gscale_buf, c, c_pad, c_in,
                      c_in_pad, taps, out=(None, None, None)):
    """-> (dW (c, c_in, taps), d_gamma [c], d_beta [c]) of the expand layer (vp3d_expand_bwd_finish)."""
    dev = p_packed.device
    dw = _grad_out(out[0], (c, c_in, taps), dev)
    dgb = (_grad_out(out[1], (c,), dev), _grad_out(out[2], (c,), dev))
    with torch.cuda.device(dev):
        check(lib().vp3d_expand_bwd_finish(dt, _ptr(p_packed), _ptr(wg), _ptr(gram), _ptr(w), int(k_total), int(ones_col),
                                           _ptr(scale), _ptr(mean), _ptr(invstd), _ptr(gscale_buf), c, c_pad, c_in,
                                           c_in_pad, taps, _ptr(dw), _ptr(dgb[0]), _ptr(dgb[1]), _stream()),
              'expand_bwd_finish')
    return dw, dgb[0], dgb[1]


def bn_act_fwd(dt, z, scale, shift, seqs, rows_per_seq, drop, res=None, res_seq_rows=0, res_row_mul=1, res_row_off=0,
               want_mask=False):
    """-> a, or (a, keep_mask uint8 [rows][c_pad / 8]) with want_mask: the keep bits the backward reads instead of
    recomputing the dropout stream and the ReLU decision (vp3d_bn_act_fwd_mask)."""
    c_pad = z.shape[-1]
    a = torch.empty_like(z)
    with torch.cuda.device(z.device):
        if want_mask:
            mask = torch.empty((seqs * rows_per_seq, c_pad // 8), dtype=torch.uint8, device=z.device)
            check(lib().vp3d_bn_act_fwd_mask(dt, _ptr(z), _ptr(scale), _ptr(shift), _ptr(res), seqs, rows_per_seq,
                                             res_seq_rows, res_row_mul, res_row_off, c_pad, C.byref(drop), _ptr(a),
                                             _ptr(mask), _stream()), 'bn_act_fwd_mask')
            return a, mask
        check(lib().vp3d_bn_act_fwd(dt, _ptr(z), _ptr(scale), _ptr(shift), _ptr(res), seqs, rows_per_seq, res_seq_rows,
                                    res_row_mul, res_row_off, c_pad, C.byref(drop), _ptr(a), _stream()), 'bn_act_fwd')
    return a


def bn_finalize_act_fwd(dt, z, stat, count, bn, seqs, rows_per_seq, drop, res=None, res_seq_rows=0, res_row_mul=1,
                        res_row_off=0, update_running=True):
    """bn_finalize + bn_act_fwd in one launch -> (a, scale, shift, mean, invstd); same results as the two calls."""
    c_pad = z.shape[-1]
    c = bn.num_features
    dev = z.device
    out = torch.empty((4, c_pad), dtype=torch.float32, device=dev)
    a = torch.empty_like(z)
    track = update_running and bn.track_running_stats and bn.running_mean is not None
    if bn.momentum is None:
        raise RuntimeError('vp3d_b200: bn_finalize_act_fwd needs a numeric BatchNorm momentum (momentum=None, the '
                           'cumulative average, is served by bn_finalize + bn_act_fwd)')
    momentum = float(bn.momentum)
    with torch.cuda.device(dev):
        check(lib().vp3d_bn_finalize_act_fwd(
            dt, _ptr(z), _ptr(stat[0]), _ptr(stat[1]), int(count), _ptr(f32c(bn.weight.detach())),
            _ptr(f32c(bn.bias.detach())), float(bn.eps), momentum, _ptr(bn.running_mean) if track else None,
            _ptr(bn.running_var) if track else None, _ptr(bn.num_batches_tracked) if track else None,
            _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(out[3]), c, _ptr(res), seqs, rows_per_seq, res_seq_rows,
            res_row_mul, res_row_off, c_pad, C.byref(drop), _ptr(a), _stream()), 'bn_finalize_act_fwd')
    if track:
        _bump_running_stats(bn)
    return a, out[0], out[1], out[2], out[3]


def bn_act_bwd(dt, g, z, scale, shift, mean, invstd, rows, c, drop, gscale_buf, count=None, group=None, sums=None,
               out=(None, None), mask=None):
    """-> (dz operand-typed [rows][c_pad], d_gamma [c], d_beta [c]). `count` (>= rows) is the number of rows the
    batch statistics were taken over; with `group` the per-channel sums are all-reduced first (SyncBN). The kernel
    then writes the GLOBAL sums as d_gamma / d_beta; they are divided by the group size here so that the gradient
    AVERAGE a data-parallel hook takes afterwards (ddp.GradSync, like torch DDP) leaves the gradient of the global mean
    loss -- the same convention as every other parameter, whose per-rank gradients are local sums."""
    c_pad = z.shape[-1]
    dev = z.device
    if sums is None:     # [2][c_pad] doubles, zero on entry (callers with many layers pass slices of one arena)
        sums = torch.zeros((2, c_pad), dtype=torch.float64, device=dev)
    dz = torch.empty_like(z)
    dgb = (_grad_out(out[0], (c,), dev), _grad_out(out[1], (c,), dev))
    with torch.cuda.device(dev):
        if mask is not None:
            # the forward stored the keep bits: no dropout stream, no affine comparison (vp3d_bn_act_bwd_*_mask)
            ks = float(keep_scale(drop.p))
            check(lib().vp3d_bn_act_bwd_reduce_mask(dt, _ptr(g), _ptr(z), _ptr(mask), _ptr(mean), _ptr(invstd), ks, rows,
                                                    c_pad, _ptr(sums[0]), _ptr(sums[1]), _stream()),
                  'bn_act_bwd_reduce_mask')
        else:
            check(lib().vp3d_bn_act_bwd_reduce(dt, _ptr(g), _ptr(z), _ptr(scale), _ptr(shift), _ptr(mean), _ptr(invstd),
                                               rows, c_pad, C.byref(drop), _ptr(sums[0]), _ptr(sums[1]), _stream()),
                  'bn_act_bwd_reduce')
        if group is not None:
            import torch.distributed as dist
            dist.all_reduce(sums, group=group)
        if mask is not None:
            check(lib().vp3d_bn_act_bwd_apply_mask(dt, _ptr(g), _ptr(z), _ptr(mask), _ptr(scale), _ptr(mean),
                                                   _ptr(invstd), ks, rows, int(count or rows), c, c_pad, _ptr(sums[0]),
                                                   _ptr(sums[1]), _ptr(gscale_buf), _ptr(dz), _ptr(dgb[0]),
                                                   _ptr(dgb[1]), _stream()), 'bn_act_bwd_apply_mask')
        else:
            check(lib().vp3d_bn_act_bwd_apply(dt, _ptr(g), _ptr(z), _ptr(scale), _ptr(shift), _ptr(mean), _ptr(invstd),
                                              rows, int(count or rows), c, c_pad, C.byref(drop), _ptr(sums[0]),
                                              _ptr(sums[1]), _ptr(gscale_buf), _ptr(dz), _ptr(dgb[0]), _ptr(dgb[1]),
                                              _stream()), 'bn_act_bwd_apply')
        if group is not None:
            import torch.distributed as dist
            torch._foreach_mul_(list(dgb), 1.0 / dist.get_world_size(group))
    return dz, dgb[0], dgb[1]


def grad_scale(dy):
    """dy fp32 contiguous -> device buffer {gscale, 1 / gscale, max|dy|}."""
    buf = torch.empty(4, dtype=torch.float32, device=dy.device)
    with torch.cuda.device(dy.device):
        check(lib().vp3d_grad_scale(_ptr(dy), dy.numel(), _ptr(buf), _stream()), 'grad_scale')
    return buf


def grad_pack_rows(dt, src, c_pad, gscale_buf, want_col_sum=False, col_sum_out=None):
    rows, c = src.shape
    dst = torch.empty((rows, c_pad), dtype=torch_dtype(dt), device=src.device)
    col_sum = None
    if want_col_sum:
        col_sum = (torch.zeros(c, dtype=torch.float32, device=src.device) if col_sum_out is None
                   else _grad_out(col_sum_out, (c,), src.device).zero_())
    with torch.cuda.device(src.device):
        check(lib().vp3d_grad_pack_rows(dt, _ptr(src), _ptr(dst), rows, c, c_pad, _ptr(gscale_buf), _ptr(col_sum),
                                        _stream()), 'grad_pack_rows')
    return dst, col_sum
