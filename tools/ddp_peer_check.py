"""Peer-memory gradient exchange on real GPUs (run under torchrun, one rank per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_peer_check.py
Checks vp3d_b200.ddp.PeerGradSync (multicast and peer-pointer paths) against dist.all_reduce(AVG) on random gradients
and times a 67.8 MB exchange (the 1f model's parameter gradients) for several CTA counts next to NCCL. Rank 0 prints one
JSON line."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dynamic-camera-augmented-videopose3d_b200'))

from vp3d_b200 import ddp  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
    dist.init_process_group('nccl')
    dev = torch.device('cuda', torch.cuda.current_device())
    out = {'world': world}
    sizes = [3 * 1024 * 1024, 1024 * 1024, 1024 * 1024 * 3, 1024, 1024, 51 * 1024, 51]
    params = [torch.nn.Parameter(torch.zeros(n, device=dev)) for n in sizes]
    for use_mc in (True, False):
        try:
            sync = ddp.PeerGradSync(params, ctas=8, use_multicast=use_mc, timeout_s=10.0)
        except Exception as e:   # noqa: BLE001
            out['setup_error'] = repr(e)
            break
        key = 'multicast' if (use_mc and sync.multicast) else ('peer' if not use_mc else 'multicast_unavailable')
        if key == 'multicast_unavailable':
            out[key] = True
            continue
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        grads = [torch.randn(n, device=dev, generator=g) for n in sizes]
        want = [x.clone() for x in grads]
        for w in want:
            dist.all_reduce(w, op=dist.ReduceOp.AVG)
        torch.cuda.synchronize()
        sync.begin()
        for p, x in zip(params, grads):
            slot = sync.alloc(p)
            slot.copy_(x)
            sync(p, slot.view_as(p))
        sync.finish()
        torch.cuda.synchronize()
        err = max(float((sync.alloc(p) - w).abs().max()) for p, w in zip(params, want))
        # every rank holds the same bits
        chk = torch.stack([sync.alloc(p).double().sum() for p in params])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        out[key] = {'max_abs_err_vs_nccl': err, 'replicas_identical': bool(torch.equal(lo, hi)),
                    'collectives': sync.collectives}
        # timing: one 67.8 MB slice (16.95 M floats)
        n_big = 16_952_320
        big = [torch.nn.Parameter(torch.zeros(n_big, device=dev))]
        times = {}
        for ctas in (4, 8, 16):
            s2 = ddp.PeerGradSync(big, ctas=ctas, use_multicast=use_mc, reserve_sms=0)
            off, n = s2.slots[id(big[0])]

            def run(s2=s2, off=off, n=n):
                s2._exchange(off, n)
                torch.cuda.current_stream().wait_stream(s2.comm)
            times[ctas] = timed(run)
            del s2
        out[key]['ms_67.8MB_by_ctas'] = times
    t = torch.zeros(16_952_320, device=dev)
    out['nccl_ms_67.8MB'] = timed(lambda: dist.all_reduce(t, op=dist.ReduceOp.AVG))
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
