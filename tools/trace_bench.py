#!/usr/bin/env python
"""Device-resident timing of K2 (closest hit) on the 1 M-triangle scene: un-jittered 1080p primary rays in
row-major and in 8x4-tile order, plus an incoherent ray set.  Development aid (kernel variants are picked
with the MB200_* environment knobs); bench.py is the judged harness."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mallie_b200 as M  # noqa: E402
from mallie_b200.procedural import bumpy_sphere  # noqa: E402

N = int(os.environ.get("SPHERE_N", "500"))
W, H = 1920, 1080
v, f = bumpy_sphere(N)
sc = M.Scene(v, f)
frame = M.camera_frame((0, 0, 3), (0, 0, 0), width=W, height=H)
stream = torch.cuda.ExternalStream(sc.stream())
n = W * H
rays = sc.generate_rays_grid(frame, 0, 0, W, H)
# 8x4 tile-major order
idx = np.arange(n).reshape(H // 4, 4, W // 8, 8).transpose(0, 2, 1, 3).reshape(-1)
sets = {"rowmajor": rays, "tile8x4": rays[idx]}
rng = np.random.default_rng(1)
d = rng.normal(size=(n, 3))
d /= np.linalg.norm(d, axis=1, keepdims=True)
org = d * 3.0
tgt = rng.uniform(-1, 1, (n, 3)) * 0.9
dr = tgt - org
dr /= np.linalg.norm(dr, axis=1, keepdims=True)
sets["incoherent"] = np.concatenate([org, dr], 1)
REP = int(os.environ.get("REP", "1"))   # REP > 1: the ray set repeated (a 16-spp frame has 16x the rays per launch)
if REP > 1:
    sets = {k + f"x{REP}": np.tile(r, (REP, 1)) for k, r in sets.items() if k != "incoherent"}
    n = n * REP
tag = " ".join(f"{k}={os.environ[k]}" for k in sorted(os.environ) if k.startswith("MB200_"))
ref = None
for name, r in sets.items():
    d_rays = torch.from_numpy(np.ascontiguousarray(r)).cuda()
    d_hits = torch.empty(n * 4, dtype=torch.float64, device="cuda")
    ts = []
    for it in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sc.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr())
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    hits = d_hits.cpu().numpy().view(M.capi.HIT_DTYPE)
    nh = int((hits["faceID"] != 0xFFFFFFFF).sum())
    chk = int(hits["faceID"].astype(np.uint64).sum())
    print(f"[{tag}] {name:10s} best {min(ts[2:]):.3f} ms median {np.median(ts[2:]):.3f} ms -> {n/min(ts[2:])/1e3:.0f} Mrays/s  hits {nh} chk {chk}")
