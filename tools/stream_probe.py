"""Streaming latency probe (BASELINE configs[3]): ms per frame of CausalStream.step / step_world for a few stream counts,
fused kernel against the GEMM-launch graph, device-timed (CUDA events around N steps) and host wall clock."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dynamic-camera-augmented-videopose3d_b200')); sys.path.insert(0, ROOT)
import json
import torch
from common.models.TemporalModel import TemporalModel
from oracle import temporal_model as otm
from vp3d_b200.streaming import CausalStream

FW = [3, 3, 3, 3, 3]
dev = torch.device('cuda')
model = TemporalModel(17, 2, 17, FW, causal=True, channels=1024)
model.load_state_dict(otm.init_state(17, 2, 17, FW, channels=1024, seed=1234))
model = model.to(dev).eval()
out = {}
for S in (1, 2, 4, 8, 16, 128, 1024):
    x = (torch.rand(S, 17, 2, device=dev) * 2 - 1)
    for fused in ((True, False) if S <= 8 else (False,)):
        st = CausalStream(model, S)
        st.fused = fused and st.fused
        with torch.no_grad():
            for _ in range(30):
                st.step(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 300
            t0 = time.perf_counter()
            e0.record()
            for _ in range(n):
                st.step(x)
            e1.record()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) / n * 1e3
        out['S%d_%s' % (S, 'fused' if st.fused else 'graph')] = {'ms_device': e0.elapsed_time(e1) / n, 'ms_wall': wall}
# kernel time without the host: 20 fused frames captured as one CUDA graph
for S in (1, 8):
    x = (torch.rand(S, 17, 2, device=dev) * 2 - 1).reshape(S, 34).contiguous()
    st = CausalStream(model, S)
    if not st.fused:
        continue
    with torch.no_grad():
        for _ in range(3):
            st._issue_fused(x)
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g, stream=side):
                for _ in range(20):
                    st._issue_fused(x)
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            out['S%d_fused_kernel_only' % S] = {'ms_device': e0.elapsed_time(e1) / 200, 'ms_wall': None}
        except Exception as e:   # cooperative launches may be refused inside a capture
            out['S%d_fused_kernel_only' % S] = {'error': str(e)[:200]}
print(json.dumps(out))
