"""Generate the golden fixtures in this directory by importing the REAL reference from /root/reference.

Run once in the build container (the reference is not available on the GPU box):
    python tests/golden/make_golden.py
The reference ships no tests or golden vectors of its own, so these outputs of the reference itself,
on seeded synthetic inputs, are the parity pin for oracle/ and, through it, for the CUDA path.
Large-channel cases store only (seed, input, output, state checksum): the state_dict is re-derived from the seed by
oracle.temporal_model.init_state and checked against the checksum.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, '/root/reference')
sys.path.insert(1, ROOT)

from common.models.TemporalModel import TemporalModel, TemporalModelOptimized1f  # noqa: E402  (reference)
from common import camera as rcam  # noqa: E402
from common import loss as rloss  # noqa: E402
from common.quaternion import qrot as r_qrot, qinverse as r_qinverse  # noqa: E402
from common.utils import wrap  # noqa: E402

from oracle import temporal_model as otm  # noqa: E402

torch.set_num_threads(8)


def state_checksum(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def ref_model(cls, sd, j_in, feat, j_out, fw, channels, **kw):
    m = cls(j_in, feat, j_out, fw, dropout=0.0, channels=channels, **kw)
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m


def small_models():
    out = {}
    fw, ch, j = [3, 3, 3], 32, 17
    sd = otm.init_state(j, 2, j, fw, channels=ch, seed=11)
    for k, v in sd.items():
        out['sd/' + k] = v.numpy()
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 40, j, 2, generator=g) * 2 - 1
    out['x'] = x.numpy()
    with torch.no_grad():
        out['y_full'] = ref_model(TemporalModel, sd, j, 2, j, fw, ch).eval()(x).numpy()
        out['y_causal'] = ref_model(TemporalModel, sd, j, 2, j, fw, ch, causal=True).eval()(x).numpy()
        xw = x[:, :27].contiguous()
        out['y_1f'] = ref_model(TemporalModelOptimized1f, sd, j, 2, j, fw, ch).eval()(xw).numpy()
        out['y_1f_causal'] = ref_model(TemporalModelOptimized1f, sd, j, 2, j, fw, ch, causal=True).eval()(xw).numpy()
    # dense ablation has its own weight shapes
    sdd = otm.init_state(j, 2, j, fw, channels=ch, dense=True, seed=12)
    for k, v in sdd.items():
        out['sdd/' + k] = v.numpy()
    with torch.no_grad():
        out['y_dense'] = ref_model(TemporalModel, sdd, j, 2, j, fw, ch, dense=True).eval()(x).numpy()

    # one training step (dropout 0): outputs, loss, parameter gradients, BN running statistics after the step
    for name, cls, xin in (('1f', TemporalModelOptimized1f, x[:, :27].contiguous()), ('full', TemporalModel, x)):
        m = ref_model(cls, sd, j, 2, j, fw, ch).train()
        pred = m(xin)
        tgt = torch.rand(pred.shape, generator=g) - 0.5
        loss = rloss.mpjpe(pred, tgt)
        loss.backward()
        out['train_%s/target' % name] = tgt.numpy()
        out['train_%s/pred' % name] = pred.detach().numpy()
        out['train_%s/loss' % name] = loss.detach().numpy()
        for k, p in m.named_parameters():
            out['train_%s/grad/%s' % (name, k)] = p.grad.numpy()
        for k, b in m.named_buffers():
            out['train_%s/buf/%s' % (name, k)] = b.detach().numpy()
    # API helpers
    m = ref_model(TemporalModel, sd, j, 2, j, fw, ch, causal=True)
    out['api/receptive_field'] = np.int64(m.receptive_field())
    out['api/total_causal_shift_full_causal'] = np.int64(m.total_causal_shift())
    m = ref_model(TemporalModelOptimized1f, sd, j, 2, j, fw, ch, causal=True)
    out['api/total_causal_shift_1f_causal'] = np.int64(m.total_causal_shift())
    np.savez_compressed(os.path.join(HERE, 'temporal_small.npz'), **out)


def seeded_large(name, fw, t_in, seed, j_in=17, j_out=17, causal=False, n=1):
    ch = 1024
    sd = otm.init_state(j_in, 2, j_out, fw, channels=ch, seed=seed)
    g = torch.Generator().manual_seed(seed + 1000)
    x = torch.rand(n, t_in, j_in, 2, generator=g) * 2 - 1
    with torch.no_grad():
        y = ref_model(TemporalModel, sd, j_in, 2, j_out, fw, ch, causal=causal).eval()(x).numpy()
        rf = otm.receptive_field(otm.make_plan(fw))
        y1f = ref_model(TemporalModelOptimized1f, sd, j_in, 2, j_out, fw, ch, causal=causal).eval()(
            x[:, :rf].contiguous()).numpy()
    np.savez_compressed(os.path.join(HERE, name), x=x.numpy(), y=y, y_1f=y1f, seed=np.int64(seed),
                        filter_widths=np.array(fw), causal=np.bool_(causal), j_in=np.int64(j_in),
                        j_out=np.int64(j_out), checksum=np.array(state_checksum(sd)))


def camera_cases():
    rng = np.random.default_rng(1234)
    out = {}
    T, J = 50, 17
    X = (rng.standard_normal((T, J, 3)) * 0.5 + np.array([0.2, -0.1, 4.0])).astype(np.float32)
    q = rng.standard_normal(4).astype(np.float32)
    q /= np.linalg.norm(q)
    t = rng.standard_normal(3).astype(np.float32)
    out['X'], out['q'], out['t'] = X, q, t
    out['w2c'] = rcam.world_to_camera(X, q, t)
    out['c2w'] = rcam.camera_to_world(X, q, t)
    # per-frame quaternions through the reference's qrot/qinverse with explicit broadcasting
    qf = rng.standard_normal((T, 4)).astype(np.float32)
    qf /= np.linalg.norm(qf, axis=-1, keepdims=True)
    tf = rng.standard_normal((T, 3)).astype(np.float32)
    out['qf'], out['tf'] = qf, tf
    qb = np.ascontiguousarray(np.broadcast_to(qf[:, None, :], (T, J, 4)))
    out['qrot_f'] = wrap(r_qrot, qb, X)
    out['w2c_f'] = wrap(r_qrot, wrap(r_qinverse, qb), X - tf[:, None, :])
    out['qinv_f'] = wrap(r_qinverse, qf)
    # projection: H36M cam 0 normalised intrinsics with distortion and a CMU-style linear camera
    h36m = np.array([2.2900989, 2.2875624, 0.025083065, 0.028902981, -0.20709892, 0.24777518, -0.0030751503,
                     -0.00097569887, -0.0014244716], dtype=np.float32)
    cmu = np.array([1.5625, 1.5625, 0, 0, 0, 0, 0, 0, 0], dtype=np.float32)
    cams = np.stack([h36m, cmu, h36m * np.float32(1.1)])
    Xc = (rng.standard_normal((3, T, J, 3)) * 0.7 + np.array([0.0, 0.0, 3.0])).astype(np.float32)
    # edge cases: z = 0 (-> +-inf -> clamp), 0/0 (-> NaN), saturation, negative depth
    Xc[0, 0, 0] = [1.0, -1.0, 0.0]
    Xc[0, 0, 1] = [0.0, 0.0, 0.0]
    Xc[0, 0, 2] = [5.0, -7.0, 1.0]
    Xc[0, 0, 3] = [0.3, 0.2, -2.0]
    out['Xc'], out['cams'] = Xc, cams
    with np.errstate(all='ignore'):
        out['proj'] = wrap(rcam.project_to_2d, Xc, cams)
        out['proj_linear'] = wrap(rcam.project_to_2d_linear, Xc, cams)
    px = (rng.random((T, J, 2)) * 1000).astype(np.float32)
    out['px'] = px
    out['norm_sc'] = rcam.normalize_screen_coordinates(px, w=1000, h=1002)
    out['img_sc'] = rcam.image_coordinates(out['norm_sc'], w=1000, h=1002)
    np.savez_compressed(os.path.join(HERE, 'camera.npz'), **out)


def camera_grad_cases():
    """Autograd of the reference's project_to_2d / project_to_2d_linear (camera.py:37-90) wrt the camera-space points:
    d sum(W * proj(X)) / dX, incl. saturated (clamped) ratios and negative depth."""
    import torch
    rng = np.random.default_rng(4321)
    h36m = np.array([2.2900989, 2.2875624, 0.025083065, 0.028902981, -0.20709892, 0.24777518, -0.0030751503,
                     -0.00097569887, -0.0014244716], dtype=np.float32)
    cams = np.stack([h36m, h36m * np.float32(0.9), h36m * np.float32(1.1)])
    X = (rng.standard_normal((3, 20, 17, 3)) * 0.7 + np.array([0.0, 0.0, 3.0])).astype(np.float32)
    X[0, 0, 0] = [5.0, -7.0, 1.0]       # both ratios saturate: gradient 0
    X[0, 0, 1] = [0.3, 0.2, -2.0]       # negative depth
    X[0, 0, 2] = [2.5, 0.1, 2.0]        # x saturates, y does not
    W = rng.standard_normal((3, 20, 17, 2)).astype(np.float32)
    out = {'X': X, 'cams': cams, 'W': W}
    for name, fn in (('grad', rcam.project_to_2d), ('grad_linear', rcam.project_to_2d_linear)):
        xt = torch.from_numpy(X).clone().requires_grad_(True)
        (fn(xt, torch.from_numpy(cams)) * torch.from_numpy(W)).sum().backward()
        out[name] = xt.grad.numpy()
    np.savez_compressed(os.path.join(HERE, 'camera_grad.npz'), **out)


def generator_cases():
    """Epoch order and window placement of the reference's ChunkedGenerator (generators.py:11-137) on a small synthetic
    set: the (seq, start_3d, end_3d) chunks of the first batches of two epochs, and one assembled 2-D batch row."""
    from common.generators import ChunkedGenerator
    rng = np.random.default_rng(99)
    lens = [37, 64, 21, 50]
    p3 = [rng.standard_normal((n, 17, 3)).astype(np.float32) for n in lens]
    p2 = [rng.standard_normal((n, 17, 2)).astype(np.float32) for n in lens]
    cams = [{'extrinsics': rng.standard_normal((n, 3, 4)).astype(np.float32),
             'intrinsics': {'focal_length': (1.5, 1.5), 'center': (0.0, 0.0)}} for n in lens]
    out = {'lens': np.array(lens)}
    for chunk, pad, shift, tag in ((1, 13, 0, 'a'), (3, 4, 4, 'b')):
        gen = ChunkedGenerator(16, cams, p3, p2, chunk, pad=pad, causal_shift=shift, shuffle=True, random_seed=1234)
        out['pairs_' + tag] = np.array([[int(a), int(b), int(c)] for a, b, c in gen.pairs])
        order = []
        rows = []
        for epoch in range(2):
            st = np.random.RandomState(1234) if False else None
            start_idx, pairs = gen.next_pairs()
            order.append(np.array(pairs[:48]).astype(np.int64))
            for _cam, b3, b2 in gen.next_epoch():
                rows.append(np.array(b2[:4]).astype(np.float32))
                break
        out['order_' + tag] = np.stack(order)
        out['batch2d_' + tag] = np.stack(rows)
        out['params_' + tag] = np.array([chunk, pad, shift])
    for i, a in enumerate(p2):
        out['p2_%d' % i] = a
    np.savez_compressed(os.path.join(HERE, 'generator.npz'), **out)


def unchunked_generator_cases():
    """The reference's UnchunkedGenerator (generators.py:140-205) on synthetic dynamic-camera sequences: per sequence
    the 2-D input it yields (edge-padded by (pad + causal_shift, pad - causal_shift)), the padded K @ [R|t] matrices and
    the 3-D target, where the generator's inputs are themselves made with the reference's own camera functions from
    world-space joints + one quaternion / translation per frame (qinverse, qrot, project_to_2d; run.py:72-74 for the
    root-relative target). This is the end-to-end expectation for vp3d_b200.feeder.DeviceSequenceFeeder."""
    from common.generators import UnchunkedGenerator
    rng = np.random.default_rng(123)
    lens = [45, 70, 33]
    J = 17
    intr = np.array([1.5625, 1.5625, 0.0, 0.0, 0, 0, 0, 0, 0], np.float32)      # CMUMocapDataset.py:53-69, normalised
    out = {'lens': np.array(lens), 'intrinsics': intr}
    cams, p3, p2 = [], [], []
    for i, n in enumerate(lens):
        X = (np.cumsum(rng.normal(0, 0.02, (n, J, 3)), axis=0) + np.array([0, 0, 4.0])).astype(np.float32)
        q = np.array([1, 0, 0, 0], np.float32) + np.cumsum(rng.normal(0, 0.004, (n, 4)), axis=0).astype(np.float32)
        q = (q / np.linalg.norm(q, axis=-1, keepdims=True)).astype(np.float32)
        t = np.cumsum(rng.normal(0, 0.01, (n, 3)), axis=0).astype(np.float32)
        qt, tt = torch.from_numpy(q), torch.from_numpy(t)
        qi = r_qinverse(qt)                                                        # (n, 4)
        xc = r_qrot(qi[:, None, :].expand(n, J, 4).contiguous(), torch.from_numpy(X) - tt[:, None, :])
        x2 = rcam.project_to_2d(xc, torch.from_numpy(intr)[None].expand(n, 9).contiguous())
        # extrinsics [R | -R c] of the same pose: columns of R = qrot(qinverse(q), e_k)
        eye = torch.eye(3)
        R = torch.stack([r_qrot(qi, eye[k].expand(n, 3).contiguous()) for k in range(3)], dim=-1)      # (n, 3, 3)
        tc = r_qrot(qi, -tt)
        E = torch.cat([R, tc[:, :, None]], dim=-1).numpy().astype(np.float32)
        cams.append({'extrinsics': E, 'intrinsics': {'focal_length': (float(intr[0]), float(intr[1])),
                                                     'center': (float(intr[2]), float(intr[3]))},
                     'cam_velocity': rng.normal(0, 1, 3), 'cam_acceleration': rng.normal(0, 1, 3),
                     'cam_angular_velocity': rng.normal(0, 1, 3), 'cam_angular_acceleration': rng.normal(0, 1, 3)})
        xcn = xc.numpy()
        p3.append((xcn - xcn[:, :1]).astype(np.float32))                           # run.py:72-74
        p2.append(x2.numpy().astype(np.float32))
        out['world_%d' % i], out['q_%d' % i], out['t_%d' % i] = X, q, t
    for pad, shift, tag in ((13, 0, 'a'), (13, 13, 'b')):
        gen = UnchunkedGenerator(cams, p3, p2, pad=pad, causal_shift=shift)
        assert gen.num_frames() == sum(lens)
        for i, (bc, b3, b2, info) in enumerate(gen.next_epoch()):
            out['cam_%s_%d' % (tag, i)] = bc.astype(np.float32)
            out['b3d_%s_%d' % (tag, i)] = b3.astype(np.float32)
            out['b2d_%s_%d' % (tag, i)] = b2.astype(np.float32)
            assert info['cam_velocity'] is cams[i]['cam_velocity']
        out['params_' + tag] = np.array([pad, shift])
    np.savez_compressed(os.path.join(HERE, 'generator_unchunked.npz'), **out)


def loss_cases():
    g = torch.Generator().manual_seed(77)
    out = {}
    N, T, J = 6, 5, 17
    pred = torch.randn(N, T, J, 3, generator=g)
    tgt = torch.randn(N, T, J, 3, generator=g)
    tgt[0, 0, 0] = pred[0, 0, 0]  # zero distance: gradient must be 0 there
    out['pred'], out['tgt'] = pred.numpy(), tgt.numpy()
    p = pred.clone().requires_grad_(True)
    l = rloss.mpjpe(p, tgt)
    l.backward()
    out['mpjpe'], out['mpjpe_grad'] = l.detach().numpy(), p.grad.numpy()
    for wname, w in (('w_n', torch.rand(N, generator=g).view(N, 1, 1) + 0.5),
                     ('w_nt1', torch.rand(N, T, 1, generator=g) + 0.5),
                     ('w_ntj', torch.rand(N, T, J, generator=g) + 0.5)):
        p = pred.clone().requires_grad_(True)
        l = rloss.weighted_mpjpe(p, tgt, w)
        l.backward()
        out[wname] = w.numpy()
        out['wmpjpe_' + wname] = l.detach().numpy()
        out['wmpjpe_grad_' + wname] = p.grad.numpy()
    out['n_mpjpe'] = rloss.n_mpjpe(pred, tgt).numpy()
    P, Tg = pred.numpy().reshape(-1, J, 3), tgt.numpy().reshape(-1, J, 3)
    out['p_mpjpe'] = np.float64(rloss.p_mpjpe(P.copy(), Tg.copy()))
    out['mve'] = np.float64(rloss.mean_velocity_error(P[:, 0].copy(), Tg[:, 0].copy()))
    np.savez_compressed(os.path.join(HERE, 'loss.npz'), **out)


if __name__ == '__main__':
    small_models()
    seeded_large('temporal_27f_1024.npz', [3, 3, 3], 60, seed=7)
    seeded_large('temporal_243f_1024.npz', [3, 3, 3, 3, 3], 250, seed=8)
    seeded_large('temporal_243f_1024_causal.npz', [3, 3, 3, 3, 3], 247, seed=9, causal=True)
    seeded_large('temporal_243f_j31.npz', [3, 3, 3, 3, 3], 245, seed=10, j_in=31, j_out=31)
    camera_cases()
    camera_grad_cases()
    generator_cases()
    unchunked_generator_cases()
    loss_cases()
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)))
