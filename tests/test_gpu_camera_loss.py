"""K5 (projection) and K6 (losses) on the GPU against the reference golden vectors and the NumPy/torch oracle.
2-D projections must match fp32 to 1e-5 (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import load_golden  # noqa: E402
from common import camera as cam  # noqa: E402
from common import loss as closs  # noqa: E402
from common.quaternion import qinverse, qrot  # noqa: E402
from oracle import camera as ocam  # noqa: E402
from oracle import loss as oloss  # noqa: E402

PROJ_TOL = 1e-5


def test_rigid_transforms_against_golden():
    z = load_golden('camera.npz')
    X, q, t = z['X'], z['q'], z['t']
    np.testing.assert_allclose(cam.world_to_camera(X, q, t), z['w2c'], atol=PROJ_TOL)
    np.testing.assert_allclose(cam.camera_to_world(X, q, t), z['c2w'], atol=PROJ_TOL)
    assert cam.world_to_camera(X, q, t).dtype == np.float32
    Xc, qf, tf = torch.from_numpy(X).cuda(), torch.from_numpy(z['qf']).cuda(), torch.from_numpy(z['tf']).cuda()
    qb = qf[:, None, :].expand(-1, X.shape[1], -1).contiguous()
    np.testing.assert_allclose(qrot(qb, Xc).cpu().numpy(), z['qrot_f'], atol=PROJ_TOL)
    np.testing.assert_allclose(cam.world_to_camera(Xc, qf, tf).cpu().numpy(), z['w2c_f'], atol=PROJ_TOL)
    np.testing.assert_array_equal(qinverse(qf).cpu().numpy(), z['qinv_f'])
    rt = cam.camera_to_world(cam.world_to_camera(Xc, qf, tf), qf, tf)
    np.testing.assert_allclose(rt.cpu().numpy(), X, atol=1e-5)
    with pytest.raises(AssertionError):
        qrot(qb[:, :3], Xc)


def test_projection_against_golden_including_edge_cases():
    z = load_golden('camera.npz')
    Xc, cams = torch.from_numpy(z['Xc']).cuda(), torch.from_numpy(z['cams']).cuda()
    p = cam.project_to_2d(Xc, cams).cpu().numpy()
    pl = cam.project_to_2d_linear(Xc, cams).cpu().numpy()
    np.testing.assert_allclose(p, z['proj'], atol=PROJ_TOL, equal_nan=True)
    np.testing.assert_allclose(pl, z['proj_linear'], atol=PROJ_TOL, equal_nan=True)
    assert np.isnan(p[0, 0, 1]).all() and np.isfinite(p[0, 0, 0]).all()   # 0/0 -> NaN, x/0 -> clamp(+-inf)
    with pytest.raises(AssertionError):
        cam.project_to_2d(Xc, cams[:2])
    with pytest.raises(AssertionError):
        cam.project_to_2d(Xc, cams[:, :8])
    # NumPy path through wrap(..., unsqueeze=True) as data/prepare_data_h36m.py:161 calls it
    from common.utils import wrap
    one = wrap(cam.project_to_2d, z['Xc'][0], z['cams'][0], unsqueeze=True)
    np.testing.assert_allclose(one, z['proj'][0], atol=PROJ_TOL, equal_nan=True)


@pytest.mark.parametrize('per_frame_intrinsics', [False, True])
@pytest.mark.parametrize('exact', [True, False])
def test_fused_dynamic_camera_projection(per_frame_intrinsics, exact):
    rng = np.random.default_rng(4)
    S, T, J = 5, 243, 17
    X = (rng.standard_normal((S, T, J, 3)) * 0.4 + np.array([0, 0, 4.0])).astype(np.float32)
    q = rng.standard_normal((S, T, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=-1, keepdims=True)
    q = q * 0.05 + np.array([1, 0, 0, 0], dtype=np.float32)
    q = (q / np.linalg.norm(q, axis=-1, keepdims=True)).astype(np.float32)
    t = (rng.standard_normal((S, T, 3)) * 0.2).astype(np.float32)
    base = np.array([2.2900989, 2.2875624, 0.025083065, 0.028902981, -0.20709892, 0.24777518, -0.0030751503,
                     -0.00097569887, -0.0014244716], dtype=np.float32)
    if per_frame_intrinsics:
        cams = (base * (1 + 0.01 * rng.standard_normal((S, T, 1)))).astype(np.float32)
    else:
        cams = (base * (1 + 0.01 * rng.standard_normal((S, 1)))).astype(np.float32)
    Xc_ref = ocam.world_to_camera(X.reshape(S * T, J, 3), q.reshape(S * T, 4), t.reshape(S * T, 3))
    if per_frame_intrinsics:
        p_ref = ocam.project_to_2d(Xc_ref, cams.reshape(S * T, 9)).reshape(S, T, J, 2)
    else:
        p_ref = ocam.project_to_2d(Xc_ref.reshape(S, T, J, 3), cams)
    # exact: operation by operation like the reference; default: fused multiply-adds (VP3D_PT_FAST), same 1e-5 budget
    x3, x2 = cam.world_to_image(*(torch.from_numpy(v).cuda() for v in (X, q, t, cams)), exact=exact)
    np.testing.assert_allclose(x3.cpu().numpy().reshape(S * T, J, 3), Xc_ref, atol=PROJ_TOL)
    np.testing.assert_allclose(x2.cpu().numpy(), p_ref, atol=PROJ_TOL)
    # odd point counts / unaligned views take the scalar path and must agree bit for bit
    x3b, x2b = cam.world_to_image(*(torch.from_numpy(v).cuda() for v in (X[:, :3, :], q[:, :3], t[:, :3],
                                                                          cams[:, :3] if per_frame_intrinsics else cams)),
                                  exact=exact)
    assert torch.equal(x2b, x2[:, :3]) and torch.equal(x3b, x3[:, :3])


def test_losses_against_golden_values_and_gradients():
    z = load_golden('loss.npz')
    pred, tgt = torch.from_numpy(z['pred']).cuda(), torch.from_numpy(z['tgt']).cuda()
    p = pred.clone().requires_grad_(True)
    l = closs.mpjpe(p, tgt)
    l.backward()
    np.testing.assert_allclose(l.item(), z['mpjpe'], rtol=1e-6)
    np.testing.assert_allclose(p.grad.cpu().numpy(), z['mpjpe_grad'], atol=1e-8, rtol=1e-5)
    assert (p.grad[0, 0, 0] == 0).all()
    for wname in ('w_n', 'w_nt1', 'w_ntj'):
        w = torch.from_numpy(z[wname]).cuda()
        p = pred.clone().requires_grad_(True)
        l = closs.weighted_mpjpe(p, tgt, w)
        l.backward()
        np.testing.assert_allclose(l.item(), z['wmpjpe_' + wname], rtol=1e-6)
        np.testing.assert_allclose(p.grad.cpu().numpy(), z['wmpjpe_grad_' + wname], atol=1e-8, rtol=1e-5)
    np.testing.assert_allclose(closs.n_mpjpe(pred, tgt).item(), z['n_mpjpe'], rtol=1e-5)
    with pytest.raises(AssertionError):
        closs.mpjpe(pred, tgt[:, :2])
    with pytest.raises(RuntimeError):
        closs.mpjpe(pred.cpu(), tgt.cpu())


def test_mpjpe_large_and_ragged_sizes_against_oracle():
    g = torch.Generator().manual_seed(3)
    for shape in [(1, 1, 1, 3), (7, 3, 17, 3), (1024, 1, 17, 3), (33, 243, 31, 3)]:
        pred, tgt = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
        ref = oloss.mpjpe(pred.double(), tgt.double()).item()
        got = closs.mpjpe(pred.cuda(), tgt.cuda()).item()
        assert abs(got - ref) < 1e-6 * max(1.0, abs(ref)), shape


def test_projection_backward_against_reference_autograd():
    """project_to_2d / project_to_2d_linear are differentiable wrt the camera-space points (camera.py:39-40); gradient
    against autograd of the reference itself (golden) and the oracle's closed form."""
    from oracle import camera as ocam
    z = load_golden('camera_grad.npz')
    cams = torch.from_numpy(z['cams']).cuda()
    W = torch.from_numpy(z['W']).cuda()
    for key, fn, lin in (('grad', cam.project_to_2d, False), ('grad_linear', cam.project_to_2d_linear, True)):
        X = torch.from_numpy(z['X']).cuda().requires_grad_(True)
        (fn(X, cams) * W).sum().backward()
        np.testing.assert_allclose(X.grad.cpu().numpy(), z[key], atol=PROJ_TOL)
        np.testing.assert_allclose(X.grad.cpu().numpy(), ocam.project_to_2d_grad(z['X'], z['cams'], z['W'], linear=lin),
                                   atol=PROJ_TOL)


def test_mpjpe_on_2d_points_value_and_gradient():
    """loss.py:17 norms the last axis whatever its length; upstream applies mpjpe to 2-D reprojections."""
    g = torch.Generator().manual_seed(2)
    pred = torch.randn(7, 3, 17, 2, generator=g)
    tgt = torch.randn(7, 3, 17, 2, generator=g)
    w = torch.rand(7, 1, 1, generator=g)
    for weights in (None, w):
        pc = pred.clone().requires_grad_(True)
        pg = pred.cuda().requires_grad_(True)
        if weights is None:
            ref = torch.mean(torch.norm(pc - tgt, dim=3))
            got = closs.mpjpe(pg, tgt.cuda())
        else:
            ref = torch.mean(weights * torch.norm(pc - tgt, dim=3))
            got = closs.weighted_mpjpe(pg, tgt.cuda(), weights.cuda())
        ref.backward()
        got.backward()
        assert abs(got.item() - ref.item()) < 1e-6
        np.testing.assert_allclose(pg.grad.cpu().numpy(), pc.grad.numpy(), atol=1e-8, rtol=1e-5)


def test_n_mpjpe_gradient_matches_autograd_of_the_reference_formula():
    """n_mpjpe differentiates through its own scale factor (loss.py:77-80)."""
    g = torch.Generator().manual_seed(6)
    for J in (17, 31, 40):
        pred = torch.randn(5, 4, J, 3, generator=g)
        tgt = pred * 0.7 + torch.randn(5, 4, J, 3, generator=g) * 0.2
        pc = pred.clone().requires_grad_(True)
        oloss.n_mpjpe(pc, tgt).backward()
        pg = pred.cuda().requires_grad_(True)
        val = closs.n_mpjpe(pg, tgt.cuda())
        val.backward()
        assert abs(val.item() - oloss.n_mpjpe(pred, tgt).item()) < 1e-6
        np.testing.assert_allclose(pg.grad.cpu().numpy(), pc.grad.numpy(), atol=2e-8, rtol=2e-4)


def test_gpu_eval_metrics_match_the_numpy_reference_semantics():
    """p_mpjpe (Procrustes, 3x3 SVD per pose) and mean_velocity_error on the device against the oracle's NumPy
    restatement and the reference's golden values."""
    z = load_golden('loss.npz')
    pred, tgt = z['pred'], z['tgt']                      # (6, 5, 17, 3)
    p2, t2 = pred.reshape(-1, 17, 3), tgt.reshape(-1, 17, 3)
    got = closs.p_mpjpe(torch.from_numpy(p2).cuda(), torch.from_numpy(t2).cuda()).item()
    assert abs(got - float(z['p_mpjpe'])) < 2e-6 * max(1.0, abs(float(z['p_mpjpe'])))
    pv, tv = pred[0], tgt[0]                             # (5, 17, 3): frames along axis 0
    got_v = closs.mean_velocity_error(torch.from_numpy(pv).cuda(), torch.from_numpy(tv).cuda()).item()
    assert abs(got_v - oloss.mean_velocity_error(pv, tv)) < 1e-6
    rng = np.random.default_rng(3)
    for J in (17, 31):
        t = rng.standard_normal((300, J, 3)).astype(np.float32)
        # similarity-transformed + noisy copies, incl. a mirrored pose (reflection branch) and a planar pose (rank 2)
        ang = rng.standard_normal((300, 3))
        p = (t * 1.3 + 0.05 * rng.standard_normal(t.shape)).astype(np.float32)
        p[0, :, 0] *= -1
        p[1, :, 2] = 0
        t[1, :, 2] = 0
        want = oloss.p_mpjpe(p.copy(), t.copy())
        got = closs.p_mpjpe(torch.from_numpy(p).cuda(), torch.from_numpy(t).cuda()).item()
        assert abs(got - want) < 5e-6 * max(1.0, want), (J, got, want)
    # NumPy inputs keep the reference's host semantics
    assert abs(closs.p_mpjpe(p2.copy(), t2.copy()) - float(z['p_mpjpe'])) < 1e-6


@pytest.mark.parametrize('S,T,J,linear,per_frame', [
    (1, 1, 2, False, False),      # one frame, smallest joint count the frame kernel takes
    (2, 3, 5, True, False),       # fewer frames than one 4-frame tile: ragged tile only
    (3, 50, 31, False, True),     # CMU-sized skeleton, intrinsics per frame
    (1, 9, 300, False, False),    # very wide frames: tiles of 4 frames
    (4, 243, 17, False, False),   # several full tiles per CTA + a ragged last tile
    (2, 17, 1, False, False),     # J = 1 is left to the generic kernel
])
def test_frame_kernel_edge_shapes_against_the_exact_path(S, T, J, linear, per_frame):
    """project_frames_kernel (bulk-copy staged tiles, VP3D_PT_FAST arithmetic) against the un-contracted generic kernel
    (exact=True, itself pinned to the reference's golden values) on ragged / tiny / wide shapes."""
    rng = np.random.default_rng(S * 1000 + T * 10 + J)
    X = (rng.standard_normal((S, T, J, 3)) * 0.4 + np.array([0, 0, 4.0])).astype(np.float32)
    q = rng.standard_normal((S, T, 4)).astype(np.float32) * 0.05 + np.array([1, 0, 0, 0], dtype=np.float32)
    q = (q / np.linalg.norm(q, axis=-1, keepdims=True)).astype(np.float32)
    t = (rng.standard_normal((S, T, 3)) * 0.2).astype(np.float32)
    base = np.array([2.2900989, 2.2875624, 0.025083065, 0.028902981, -0.20709892, 0.24777518, -0.0030751503,
                     -0.00097569887, -0.0014244716], dtype=np.float32)
    cams = (base * (1 + 0.01 * rng.standard_normal((S, T, 1) if per_frame else (S, 1)))).astype(np.float32)
    args = [torch.from_numpy(v).cuda() for v in (X, q, t, cams)]
    c3, p2 = cam.world_to_image(*args, linear=linear)
    c3e, p2e = cam.world_to_image(*args, linear=linear, exact=True)
    np.testing.assert_allclose(c3.cpu().numpy(), c3e.cpu().numpy(), atol=PROJ_TOL)
    np.testing.assert_allclose(p2.cpu().numpy(), p2e.cpu().numpy(), atol=PROJ_TOL)
    _, p2_only = cam.world_to_image(*args, linear=linear, return_camera_space=False)
    assert torch.equal(p2_only, p2)


def test_frame_kernel_special_values_follow_the_reference():
    """camera.py:59 clamps x/z to [-1, 1]: z = 0 saturates (x != 0) or gives NaN (0/0), z < 0 flips the sign; the frame
    kernel's reciprocal-multiply must land on the same values as the exact path (project_to_2d semantics)."""
    J = 8
    X = np.zeros((1, 4, J, 3), np.float32)
    X[..., 2] = 4.0
    X[0, 0, 0] = [1.0, -2.0, 0.0]        # x/0 -> +inf -> 1, y/0 -> -inf -> -1
    X[0, 0, 1] = [0.0, 0.0, 0.0]         # 0/0 -> NaN
    X[0, 1, 2] = [3.0, 1.0, -2.0]        # behind the camera: ratios change sign, x saturates at -1
    X[0, 2, 3] = [np.inf, 1.0, 2.0]      # inf / z -> inf -> 1
    X[0, 3, 4] = [np.nan, 1.0, 2.0]      # NaN propagates through the clamp
    q = np.tile(np.array([1, 0, 0, 0], np.float32), (1, 4, 1))      # identity pose: camera space == world space
    t = np.zeros((1, 4, 3), np.float32)
    cams = np.array([[2.29, 2.2876, 0.0251, 0.0289, -0.2071, 0.2478, -0.0031, -0.00098, -0.0014]], np.float32)
    args = [torch.from_numpy(v).cuda() for v in (X, q, t, cams)]
    _, p2 = cam.world_to_image(*args)
    _, p2e = cam.world_to_image(*args, exact=True)
    # the reference composition: qrot by the (identity) quaternion spreads an inf / NaN coordinate to the whole point
    # (0 * inf inside the cross products, quaternion.py:21-24), exactly as the rotation-matrix form of the frame kernel
    with np.errstate(invalid='ignore', divide='ignore'):
        xc = ocam.world_to_camera(X.reshape(4, J, 3), q.reshape(4, 4), t.reshape(4, 3))
        want = ocam.project_to_2d(xc.reshape(1, 4 * J, 3), cams).reshape(1, 4, J, 2)
    got, gote = p2.cpu().numpy(), p2e.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isnan(gote), np.isnan(want))
    ok = ~np.isnan(want)
    np.testing.assert_allclose(got[ok], want[ok], atol=PROJ_TOL)
    np.testing.assert_allclose(gote[ok], want[ok], atol=PROJ_TOL)


@pytest.mark.parametrize('with_traj', [False, True])
@pytest.mark.parametrize('linear', [False, True])
def test_fused_reprojection_loss_matches_the_composition(with_traj, linear):
    """vp3d_reproj_mpjpe_fwd / _bwd (projection with the loss fused in, north-star item 3) against
    mpjpe(project_to_2d(pose + traj, cam), target) composed from the drop-in functions -- value and both gradients --
    and against the CPU oracle; includes points whose x / z saturates the clamp (zero gradient through it)."""
    from common.camera import project_to_2d, project_to_2d_linear
    from common.loss import mpjpe, reprojection_mpjpe
    from oracle import camera as ocam
    from oracle import loss as oloss
    g = torch.Generator().manual_seed(91)
    N, T, J = 37, 3, 31
    pose = (torch.randn(N, T, J, 3, generator=g) * 0.4)
    traj = torch.randn(N, T, 1, 3, generator=g) * 0.2 + torch.tensor([0.0, 0.0, 3.0])
    if not with_traj:
        pose = pose + torch.tensor([0.0, 0.0, 3.0])
    pose[0, 0, :4, 0] = 9.0                      # |x / z| > 1: clamped, no gradient through x
    cam = torch.tensor([2.29, 2.2876, 0.0251, 0.0289, -0.2071, 0.2478, -0.0031, -0.00098, -0.0014]).repeat(N, 1)
    cam[:, 0] += torch.randn(N, generator=g) * 0.05
    tgt = torch.randn(N, T, J, 2, generator=g) * 0.3
    pc, tc = pose.cuda().requires_grad_(True), traj.cuda().requires_grad_(True)
    fused = reprojection_mpjpe(pc, cam.cuda(), tgt.cuda(), trajectory=tc if with_traj else None, linear=linear)
    fused.backward()
    gp_f, gt_f = pc.grad.clone(), (tc.grad.clone() if with_traj else None)
    pc2, tc2 = pose.cuda().requires_grad_(True), traj.cuda().requires_grad_(True)
    X = pc2 + tc2 if with_traj else pc2
    proj = (project_to_2d_linear if linear else project_to_2d)(X, cam.cuda())
    ref = mpjpe(proj, tgt.cuda())
    ref.backward()
    assert abs(fused.item() - ref.item()) < 1e-6 * max(1.0, abs(ref.item()))
    assert (gp_f - pc2.grad).abs().max().item() < 1e-7 + 1e-4 * pc2.grad.abs().max().item()
    if with_traj:
        assert (gt_f - tc2.grad).abs().max().item() < 1e-7 + 1e-4 * tc2.grad.abs().max().item()
    assert gp_f[0, 0, :4, 0].abs().max().item() == 0
    Xn = (pose + traj).numpy() if with_traj else pose.numpy()
    p_ref = (ocam.project_to_2d_linear if linear else ocam.project_to_2d)(Xn, cam.numpy())
    want = oloss.mpjpe(torch.from_numpy(p_ref), tgt).item()
    assert abs(fused.item() - want) < 1e-5
