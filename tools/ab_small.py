#!/usr/bin/env python
"""Fixed cost of one traversal launch: closest-hit query on 32 ... 262144 rays (development aid)."""
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
rays = rays[idx]
c0 = (H // 8) * (W // 8) * 32 + (W // 16) * 32          # a tile at the image centre
for n in (32, 1024, 32768, 262144):
    d_rays = torch.from_numpy(np.ascontiguousarray(rays[c0:c0 + n])).cuda()
    d_hits = torch.empty(n * 4, dtype=torch.float64, device="cuda")
    ts = []
    for it in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sc.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr())
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"n = {n:7d} central rays: {min(ts[2:])*1e3:.1f} us")
