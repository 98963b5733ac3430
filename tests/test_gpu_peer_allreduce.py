"""vp3d_peer_allreduce_f32 (csrc/allreduce.cu), the gradient exchange of data-parallel training, on ONE device: the
"ranks" are separate buffers of the same GPU and every rank's kernel runs on its own stream, so the rank barrier (flag
words in peer memory), the ownership split and the write-to-all-replicas step are exercised through the peer-pointer
path exactly as over NVLink (the multicast path needs a fabric; tools/ddp_peer_check.py checks it under torchrun).
No reference counterpart: the reference trains in one process (run.py:473-487)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

from vp3d_b200 import native  # noqa: E402

FLAGS = 64 * 16      # flag words at the head of every buffer (PeerGradSync.FLAG_FLOATS)


def _call(bufs, rank, off, n, ctas, stream, scale=None, timeout_s=5.0):
    world = len(bufs)
    ptrs = (C.c_void_p * world)(*[b.data_ptr() for b in bufs])
    a = native.AllReduceArgs()
    a.multicast = None
    a.peers = ptrs
    a.flags = ptrs
    a.rank, a.world, a.offset, a.count = rank, world, off, n
    a.scale, a.ctas, a.timeout_s = (1.0 / world if scale is None else scale), ctas, timeout_s
    return native.lib().vp3d_peer_allreduce_f32(C.byref(a), stream.cuda_stream)


@pytest.mark.parametrize('world,ctas', [(1, 2), (2, 2), (4, 4), (8, 2)])
def test_emulated_ranks_agree_with_the_sum(world, ctas):
    g = torch.Generator().manual_seed(world)
    total = FLAGS + 3 * 1000 * 1000 + 64
    data = [torch.randn(total - FLAGS, generator=g) for _ in range(world)]
    bufs = []
    for k in range(world):
        b = torch.zeros(total, device='cuda')
        b[FLAGS:] = data[k].cuda()
        bufs.append(b)
    streams = [torch.cuda.Stream() for _ in range(world)]
    torch.cuda.synchronize()
    # three slices, issued back to back on every rank's stream (the flag words are reused from call to call); the last
    # one is tiny (fewer vectors than threads) and the middle one is not a multiple of world * 4 floats
    slices = [(FLAGS, 2_000_000), (FLAGS + 2_000_000, 999_996), (FLAGS + 2_999_996, 68)]
    for off, n in slices:
        for k in range(world):
            native.check(_call(bufs, k, off, n, ctas, streams[k]), 'peer_allreduce')
    torch.cuda.synchronize()
    want = torch.stack([d.double() for d in data]).sum(0) / world
    for k in range(world):
        got = bufs[k][FLAGS:].cpu()
        assert torch.equal(got, bufs[0][FLAGS:].cpu()), 'replicas differ'
        lo = slices[0][0] - FLAGS
        hi = slices[-1][0] + slices[-1][1] - FLAGS
        err = (got[lo:hi].double() - want[lo:hi]).abs().max().item()
        assert err < 1e-6, err
        assert torch.equal(got[hi:], data[k][hi:]), 'elements outside the slices were touched'
        assert int(bufs[k][:FLAGS].view(torch.int32).abs().sum()) == 0, 'flag words not back to zero'


def test_argument_errors():
    b = torch.zeros(FLAGS + 64, device='cuda')
    s = torch.cuda.current_stream()
    assert _call([b], 0, FLAGS + 2, 8, 2, s) == 1          # offset not a multiple of 4 floats
    assert _call([b], 0, FLAGS, 8, 3, s) == 1              # odd CTA count (clusters of two)
    assert _call([b], 1, FLAGS, 8, 2, s) == 1              # rank outside the world
    assert _call([b], 0, FLAGS, 0, 2, s) == 0              # nothing to do
    torch.cuda.synchronize()


def test_graph_replays_reuse_the_flag_words():
    """The exchange sits inside the captured training step: replays must find the flag words as the last one left them."""
    world, ctas, n = 2, 2, 4096
    bufs = [torch.zeros(FLAGS + n, device='cuda') for _ in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    graphs = []
    for k in range(world):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.stream(streams[k]):
            with torch.cuda.graph(gr, stream=streams[k]):
                native.check(_call(bufs, k, FLAGS, n, ctas, streams[k], scale=1.0), 'peer_allreduce')
        graphs.append(gr)
    for it in range(3):
        for k in range(world):
            bufs[k][FLAGS:] = float(k + 1 + it)
        torch.cuda.synchronize()
        for k in range(world):
            with torch.cuda.stream(streams[k]):
                graphs[k].replay()
        torch.cuda.synchronize()
        for k in range(world):
            assert torch.all(bufs[k][FLAGS:] == float(3 + 2 * it))
