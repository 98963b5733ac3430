"""Host-side logic of the GPU feeder against the reference's ChunkedGenerator (tests/golden/generator.npz, produced by
importing common/generators.py): lineage pairs, RandomState(1234) epoch order over successive epochs, and the window
placement / edge padding rule the device kernel implements (clamp of start_3d - pad - causal_shift + k)."""
import numpy as np

from conftest import load_golden
from vp3d_b200.feeder import build_pairs


def test_pairs_and_epoch_order_match_reference_generator():
    z = load_golden('generator.npz')
    lens = [int(v) for v in z['lens']]
    for tag in ('a', 'b'):
        chunk, pad, shift = [int(v) for v in z['params_' + tag]]
        pairs = build_pairs(lens, chunk)
        np.testing.assert_array_equal(np.array([[int(a), int(b), int(c)] for a, b, c in pairs]), z['pairs_' + tag])
        # the golden script drew permutations 1 and 3 of RandomState(1234) for `order` (2 and 4 went to next_epoch)
        rs = np.random.RandomState(1234)
        p1 = rs.permutation(pairs)
        p2 = rs.permutation(pairs)
        p3 = rs.permutation(pairs)
        np.testing.assert_array_equal(np.asarray(p1[:48]).astype(np.int64), z['order_' + tag][0])
        np.testing.assert_array_equal(np.asarray(p3[:48]).astype(np.int64), z['order_' + tag][1])
        # window placement: frame k of sample (s, start_3d) is source frame clamp(start_3d - pad - shift + k, 0, len - 1)
        window = chunk + 2 * pad
        for i in range(4):
            s, a3, _e3 = [int(v) for v in p2[i]]
            src = np.clip(a3 - pad - shift + np.arange(window), 0, lens[s] - 1)
            np.testing.assert_array_equal(z['p2_%d' % s][src], z['batch2d_' + tag][0][i])
