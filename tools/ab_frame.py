#!/usr/bin/env python
"""A/B timing of one kernel variant (development aid; variants are picked with MB200_TRACE_* in a `make DEV=1`
build): un-jittered 1080p closest-hit query + the bench frame (16 spp primary+shadow) with per-kernel times."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mallie_b200 as M  # noqa: E402
from mallie_b200.procedural import bumpy_sphere  # noqa: E402

N = int(os.environ.get("SPHERE_N", "500"))
W, H, SPP = 1920, 1080, 16
v, f = bumpy_sphere(N)
sc = M.Scene.build(v, f, want_bvh=False)
frame = M.camera_frame((0, 0, 3), (0, 0, 0), width=W, height=H)
stream = torch.cuda.ExternalStream(sc.stream())
n = W * H
d_rays = torch.empty(n * 6, dtype=torch.float64, device="cuda")
d_hits = torch.empty(n * 4, dtype=torch.float64, device="cuda")
L, C = M.capi.lib(), M.capi.C
M.capi.check(L.mb200_generate_rays_grid(sc.h, C.byref(frame), 0, 0, W, H, M.capi._p(d_rays.data_ptr())))
d_img = torch.zeros(n * 3, dtype=torch.float32, device="cuda")
d_cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
p = sc.render_params(frame, W, H, shader=M.SHADER_PRIMARY_SHADOW, light=(2.0, 4.0, 3.0))
shader = os.environ.get("AB_SHADER")
if shader == "path":
    p = sc.render_params(frame, W, H, shader=M.SHADER_PATHTRACE)
    SPP = 2


def closest():
    sc.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr())


def frame_():
    M.capi.check(L.mb200_render_frame(sc.h, C.byref(p), SPP, M.capi._p(d_img.data_ptr()), M.capi._p(d_cnt.data_ptr()), None))


out = {}
for name, fn, reps in (("closest", closest, 12), ("frame", frame_, 6)):
    for _ in range(3):
        fn()
    sc.synchronize()
    sc.timing(True)
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    kt = sc.kernel_times()
    sc.timing(False)
    out[name] = (min(ts), float(np.median(ts)), {k: round(v / reps, 3) for k, v in kt.items() if k.endswith("_ms") and v})
hits = d_hits.cpu().numpy().view(M.capi.HIT_DTYPE)
chk = int(hits["faceID"].astype(np.uint64).sum())
img_sum = float(d_img.double().sum())
import hashlib  # noqa: E402
img_fnv = hashlib.sha1(d_img.cpu().numpy().tobytes()).hexdigest()[:12]
tag = " ".join(f"{k[6:]}={os.environ[k]}" for k in sorted(os.environ) if k.startswith("MB200_")) or "production"
print(f"[{tag}] closest {out['closest'][0]:.3f} ms ({n/out['closest'][0]/1e3:.0f} Mray/s) | frame {out['frame'][0]:.3f} ms "
      f"med {out['frame'][1]:.3f} {out['frame'][2]} | chk {chk} img {img_sum:.3f} {img_fnv}")
