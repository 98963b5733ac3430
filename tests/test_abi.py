"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/vp3d_b200.h declares,
and the Python mirror of the reference API behaves like the reference where no GPU is needed."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, state_from_npz
from vp3d_b200 import native


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'vp3d_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(vp3d_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(native.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    handle = ctypes.CDLL(native.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(handle, s), 'missing export: ' + s
    assert sorted(native.declared_symbols()) == syms, 'ctypes signature table out of sync with the header'
    assert native.lib().vp3d_version() >= 100
    assert native.lib().vp3d_loss_workspace_bytes() > 0


def test_integration_doc_names_every_entry_point_and_knob():
    """INTEGRATION.md is the maintainer's map of the boundary: every exported function and every VP3D_* environment
    variable the package reads must appear in it (families may be written `vp3d_x_fwd/bwd`, `vp3d_x_fwd` / `vp3d_x_bwd`)."""
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    doc = re.sub(r'(vp3d_[a-z0-9_]+?)_fwd/bwd', r'\1_fwd \1_bwd', doc)
    missing = [s for s in _header_symbols() if s not in doc]
    assert not missing, missing
    pkg = os.path.join(ROOT, 'dynamic-camera-augmented-videopose3d_b200')
    knobs = set()
    for sub in ('vp3d_b200', 'csrc', 'common'):
        for dirpath, _, files in os.walk(os.path.join(pkg, sub)):
            if 'build' in dirpath:
                continue
            for f in files:
                if f.endswith(('.py', '.cu', '.cuh', '.h')):
                    src = open(os.path.join(dirpath, f)).read()
                    knobs.update(re.findall(r"environ\.get\(\s*'(VP3D_[A-Z0-9_]+)'", src))
                    knobs.update(re.findall(r'getenv\("(VP3D_[A-Z0-9_]+)"\)', src))
    assert len(knobs) >= 10
    undocumented = sorted(k for k in knobs if k not in doc)
    assert not undocumented, undocumented


def test_conv_args_struct_matches_header_field_order():
    text = open(os.path.join(ROOT, 'include', 'vp3d_b200.h')).read()
    body = text[text.index('typedef struct vp3d_conv_args {'):text.index('} vp3d_conv_args;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = re.findall(r'\b(\w+);', body)
    assert fields == [f[0] for f in native.ConvArgs._fields_]


def test_model_api_matches_reference_layout():
    from common.models.TemporalModel import TemporalModel, TemporalModelBase, TemporalModelOptimized1f
    z = load_golden('temporal_small.npz')
    sd = state_from_npz(z, 'sd/')
    for cls in (TemporalModel, TemporalModelOptimized1f):
        m = cls(17, 2, 17, [3, 3, 3], channels=32)
        assert isinstance(m, TemporalModelBase)
        assert list(m.state_dict().keys()) == list(sd.keys())
        m.load_state_dict(sd, strict=True)
        assert m.receptive_field() == int(z['api/receptive_field'])
        m.set_bn_momentum(0.03)
        assert m.expand_bn.momentum == 0.03 and all(bn.momentum == 0.03 for bn in m.layers_bn)
    assert TemporalModel(17, 2, 17, [3, 3, 3], causal=True, channels=32).total_causal_shift() == \
        int(z['api/total_causal_shift_full_causal'])
    assert TemporalModelOptimized1f(17, 2, 17, [3, 3, 3], causal=True, channels=32).total_causal_shift() == \
        int(z['api/total_causal_shift_1f_causal'])
    m = TemporalModel(17, 2, 17, [3, 3, 3, 3, 3])
    assert sum(p.numel() for p in m.parameters()) == 16952371
    assert m.pad == [1, 3, 9, 27, 81] and m.causal_shift == [0] * 5
    assert [c.dilation[0] for c in m.layers_conv[::2]] == [3, 9, 27, 81]
    with pytest.raises(AssertionError):
        TemporalModel(17, 2, 17, [3, 4, 3])
    # same default initialisation stream as the reference: identical RNG consumption order
    torch.manual_seed(0)
    a = TemporalModel(17, 2, 17, [3, 3], channels=16)
    torch.manual_seed(0)
    b = TemporalModel(17, 2, 17, [3, 3], channels=16)
    assert all(torch.equal(p, q) for p, q in zip(a.parameters(), b.parameters()))


def test_no_cpu_fallback():
    from common.models.TemporalModel import TemporalModel
    from common import loss as closs
    m = TemporalModel(17, 2, 17, [3, 3, 3], channels=32).eval()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 27, 17, 2))
    with pytest.raises(RuntimeError):
        closs.mpjpe(torch.zeros(2, 1, 17, 3), torch.zeros(2, 1, 17, 3))
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 27, 16, 2))


def test_numpy_metrics_against_golden():
    from common import loss as closs
    from common import camera as cam
    z = load_golden('loss.npz')
    P, T = z['pred'].reshape(-1, 17, 3), z['tgt'].reshape(-1, 17, 3)
    np.testing.assert_allclose(closs.p_mpjpe(P.copy(), T.copy()), float(z['p_mpjpe']), rtol=1e-6)
    np.testing.assert_allclose(closs.mean_velocity_error(P[:, 0], T[:, 0]), float(z['mve']), rtol=1e-6)
    c = load_golden('camera.npz')
    ns = cam.normalize_screen_coordinates(c['px'], w=1000, h=1002)
    assert ns.dtype == np.float64
    np.testing.assert_array_equal(ns, c['norm_sc'])
    np.testing.assert_array_equal(cam.image_coordinates(ns, w=1000, h=1002), c['img_sc'])
    from common.utils import deterministic_random
    assert deterministic_random(0, 100, 'S1/Walking') == deterministic_random(0, 100, 'S1/Walking')


@pytest.mark.parametrize('fw,j,ch,cls', [([3, 3, 3, 3, 3], 17, 1024, '1f'), ([3, 3, 3], 31, 256, '1f'),
                                         ([3, 3, 3, 3], 17, 512, 'dilated'), ([5, 3], 17, 1024, 'dilated'),
                                         ([3], 17, 256, '1f')])
def test_backward_arena_bound_covers_every_accumulator(fw, j, ch, cls):
    """training._arena_floats (the zero arena of the backward is allocated during the forward, before the saved layers
    exist) must not be smaller than what the backward takes: BatchNorm sums, one split-K accumulator per convolution in
    either of its two layouts, the shrink layer's, 16-byte alignment slack."""
    from common.models.TemporalModel import TemporalModel, TemporalModelOptimized1f
    from vp3d_b200 import training
    from vp3d_b200.temporal import K_ALIGN, N_TILE, _round_up
    model = (TemporalModelOptimized1f if cls == '1f' else TemporalModel)(j, 2, j, fw, channels=ch)
    c_in = 2 * j
    n_layers = 2 * len(fw) - 1
    c_pad, c_in_pad = _round_up(ch, N_TILE), _round_up(c_in, K_ALIGN)
    takes = [n_layers * 2 * c_pad * 2]                                        # sums_all: [layers][2][c_pad] doubles
    takes.append(training.SHRINK_PAD * c_pad)                                  # shrink weight gradient
    takes.append(max(c_pad * 256, fw[0] * c_pad * c_in_pad))                   # expand: fused / narrow or generic layout
    for conv in model.layers_conv:
        takes.append(conv.kernel_size[0] * c_pad * c_pad)
    need = sum((t + 3) // 4 * 4 for t in takes)
    assert training._arena_floats(model, c_in, n_layers) >= need
