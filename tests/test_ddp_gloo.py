"""Host-side multi-GPU logic on CPU: world_size-2 gloo process groups (no CUDA involved).
 * shard_range: balanced contiguous sequence shards, no overlap, no gap (inference has no collective);
 * GradSync: the hook protocol the training backward drives -- large tensors all-reduced immediately and asynchronously,
   small ones in one flat bucket, result = average over ranks, identical on every rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vp3d_b200 import ddp, training


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 65, 1000):
        for world in (1, 2, 3, 8):
            spans = [ddp.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        sync = ddp.enable_grad_sync()
        assert training.grad_ready_hook is sync and training.grad_finish_hook is not None
        g = torch.Generator().manual_seed(100 + rank)
        shapes = [(1024, 1024, 3), (1024,), (1024,), (51, 1024, 1), (51,), (1024, 1024, 1), (1024, 34, 3)]
        grads = [torch.randn(s, generator=g) for s in shapes]
        params = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
        for p_, g_ in zip(params, grads):           # what _StackTrainFn.backward does, layer by layer
            training.grad_ready_hook(p_, g_)
        training.grad_finish_hook()
        # expected: average over ranks of the per-rank tensors
        want = []
        for s_i, s in enumerate(shapes):
            acc = torch.zeros(s)
            for r in range(world):
                gr = torch.Generator().manual_seed(100 + r)
                ts = [torch.randn(sh, generator=gr) for sh in shapes]
                acc += ts[s_i]
            want.append(acc / world)
        err = max((a - b).abs().max().item() for a, b in zip(grads, want))
        # two large tensors (12 MB, 4 MB) + one bucket for the five small ones
        out[rank] = (err, sync.collectives, sync.bytes_reduced)
        # second step reuses the object
        for p_, g_ in zip(params, grads):
            training.grad_ready_hook(p_, g_)
        training.grad_finish_hook()
        # buffers / parameters broadcast
        m = torch.nn.BatchNorm1d(8)
        m.running_mean.fill_(float(rank + 1))
        ddp.broadcast_buffers(m, src=0)
        assert torch.all(m.running_mean == 1.0)
        ddp.disable_grad_sync()
        assert training.grad_ready_hook is None
        # bf16 on the wire: same protocol, half the bytes, fp32 gradients written back
        sync16 = ddp.enable_grad_sync(compress='bf16')
        g2 = [torch.randn(s, generator=torch.Generator().manual_seed(100 + rank)) for s in shapes[:2]]
        for p_, g_ in zip(params[:2], g2):
            training.grad_ready_hook(p_, g_)
        training.grad_finish_hook()
        want2 = sum(torch.randn(shapes[0], generator=torch.Generator().manual_seed(100 + r)) for r in range(world)) / world
        rel = ((g2[0] - want2).norm() / want2.norm()).item()
        assert g2[0].dtype == torch.float32 and rel < 1e-2, rel
        assert sync16.bytes_reduced == 2 * sum(g.numel() for g in g2)
        ddp.disable_grad_sync()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_grad_sync_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        err, n_coll, n_bytes = out[r]
        assert err < 1e-6, err
        assert n_coll == 3, n_coll
        total = sum(4 * n for n in (1024 * 1024 * 3, 1024, 1024, 51 * 1024, 51, 1024 * 1024, 1024 * 34 * 3))
        assert n_bytes == total
