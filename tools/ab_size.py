#!/usr/bin/env python
"""Launch-size sweep of the closest-hit query: time = a + n / r?  (development aid)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mallie_b200 as M  # noqa: E402
from mallie_b200.procedural import bumpy_sphere  # noqa: E402

W, H = 1920, 1080
v, f = bumpy_sphere(500)
sc = M.Scene(v, f)
frame = M.camera_frame((0, 0, 3), (0, 0, 0), width=W, height=H)
stream = torch.cuda.ExternalStream(sc.stream())
rays = sc.generate_rays_grid(frame, 0, 0, W, H)
idx = np.arange(W * H).reshape(H // 4, 4, W // 8, 8).transpose(0, 2, 1, 3).reshape(-1)
rays = rays[idx]                                     # 8x4 tile order, as the frame kernels see them
base = torch.from_numpy(np.ascontiguousarray(rays)).cuda()
for rep, frac in ((1, 0.125), (1, 0.25), (1, 0.5), (1, 1.0), (2, 1.0), (4, 1.0), (8, 1.0)):
    n0 = int(W * H * frac) // 32 * 32
    # a fraction = every k-th tile row, like one rank of a multi-GPU frame
    if frac < 1.0:
        k = int(round(1 / frac))
        sel = torch.arange(W * H, device="cuda").reshape(H // 4, -1)[::k].reshape(-1)
        d_rays = base[sel].contiguous()
    else:
        d_rays = base.repeat(rep, 1).contiguous()
    n = d_rays.shape[0]
    d_hits = torch.empty(n * 4, dtype=torch.float64, device="cuda")
    ts = []
    for it in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sc.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr())
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"n = {n:9d} rays: {min(ts[2:]):.3f} ms -> {n/min(ts[2:])/1e3:.0f} Mrays/s")
