"""A/B probe of the dynamic-camera projection (world_to_image): time per launch over rotating inputs larger than L2 and
the largest difference from the un-contracted (exact=True) path. The frame kernel's tuning is read from
VP3D_PROJ_TUNE ("stages,max_points,ctas"; "0" = generic kernel only), so run once per setting:
    VP3D_PROJ_TUNE=3,2304,2 python tools/proj_probe.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'dynamic-camera-augmented-videopose3d_b200'))
from common.camera import world_to_image  # noqa: E402

peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
dev = torch.device('cuda')
tune = os.environ.get('VP3D_PROJ_TUNE', 'default')
for B, T, J, n_sets in ((1024, 243, 17, 6), (4096, 243, 17, 2), (1024, 243, 31, 4)):
    sets = []
    for i in range(n_sets):
        X = torch.randn(B, T, J, 3, device=dev) * 0.3
        X[..., 2] += 4
        q = torch.randn(B, T, 4, device=dev)
        q = q / q.norm(dim=-1, keepdim=True)
        t = torch.randn(B, T, 3, device=dev) * 0.1
        cam = torch.tensor([2.29, 2.2876, 0.0251, 0.0289, -0.2071, 0.2478, -0.0031, -0.00098, -0.0014],
                           device=dev).repeat(B, 1)
        sets.append((X, q, t, cam))
    _, ref = world_to_image(*sets[0], return_camera_space=False, exact=True)
    c3, got = world_to_image(*sets[0], return_camera_space=True)
    c3r, _ = world_to_image(*sets[0], return_camera_space=True, exact=True)
    err = (got - ref).abs().max().item()
    err3 = (c3 - c3r).abs().max().item()
    g = torch.cuda.CUDAGraph()
    for s in sets:
        world_to_image(*s, return_camera_space=False)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for rep in range(5):
            for s in sets:
                world_to_image(*s, return_camera_space=False)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (5 * n_sets)
    gb = (20 * J + 64) * B * T / 1e9
    print('tune=%s B=%d J=%d: %.4f ms  %.0f GB/s  %.3f of HBM peak   max|d2|=%.2e max|d3|=%.2e' %
          (tune, B, J, ms, gb / ms * 1e3, gb / ms * 1e3 / peak, err, err3), flush=True)
