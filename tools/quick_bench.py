#!/usr/bin/env python
"""Quick device-resident timing of the hot kernels (development aid; bench.py is the judged harness)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mallie_b200 as M  # noqa: E402
from mallie_b200.procedural import bumpy_sphere  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 500
W, H = 1920, 1080
t0 = time.time()
v, f = bumpy_sphere(N)
t1 = time.time()
sc = M.Scene(v, f)
t2 = time.time()
print(f"mesh {t1-t0:.2f}s build+upload {t2-t1:.2f}s tris {len(f)} device MB {sc.device_bytes()/1e6:.1f} f32 {sc.uses_f32_vertices()}")
frame = M.camera_frame((0, 0, 3), (0, 0, 0), width=W, height=H)
stream = torch.cuda.ExternalStream(sc.stream())
n = W * H
d_rays = torch.empty(n * 6, dtype=torch.float64, device="cuda")
d_hits = torch.empty(n * 4, dtype=torch.float64, device="cuda")
M.capi.check(M.capi.lib().mb200_generate_rays_grid(sc.h, M.capi.C.byref(frame), 0, 0, W, H, M.capi._p(d_rays.data_ptr())))


def timeit(fn, reps=10):
    fn()
    sc.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


best, med = timeit(lambda: sc.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr()))
print(f"trace_closest primary {n} rays: best {best:.3f} ms median {med:.3f} ms -> {n/best/1e3:.1f} Mrays/s")
hits = d_hits.cpu().numpy().view(M.capi.HIT_DTYPE)
print("hits", int((hits['faceID'] != 0xFFFFFFFF).sum()))

d_img = torch.zeros(n * 3, dtype=torch.float32, device="cuda")
d_cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
for shader, name in ((M.SHADER_PRIMARY_ONLY, "primary_only"), (M.SHADER_PRIMARY_SHADOW, "primary_shadow"), (M.SHADER_PATHTRACE, "pathtrace")):
    p = sc.render_params(frame, W, H, shader=shader, light=(2.0, 4.0, 3.0))
    for spp in (1, 16):
        if shader == M.SHADER_PATHTRACE and spp == 16:
            continue
        best, med = timeit(lambda: M.capi.check(M.capi.lib().mb200_render_accumulate(sc.h, M.capi.C.byref(p), spp, M.capi._p(d_img.data_ptr()), M.capi._p(d_cnt.data_ptr()), None)), reps=5)
        _, _, st = sc.render_accumulate(p, spp, d_img.data_ptr(), d_cnt.data_ptr(), stats=True)
        rays = st["primary_rays"] + st["bounce_rays"] + st["shadow_rays"]
        print(f"render {name} spp={spp}: best {best:.3f} ms median {med:.3f} -> {rays/best/1e3:.1f} Mrays/s {st}")
print("launches", M.capi.launches_issued())
