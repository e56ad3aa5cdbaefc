#!/usr/bin/env python
"""One bench frame (1 M-triangle sphere, 1080p, 16 spp, primary+shadow) for ncu captures of k_trace_sm (development aid).
SPHERE_N / PROF_W / PROF_H / PROF_SPP / PROF_FRAMES override the workload."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mallie_b200 as M  # noqa: E402
from mallie_b200.procedural import bumpy_sphere  # noqa: E402

N = int(os.environ.get("SPHERE_N", "500"))
W, H, SPP = int(os.environ.get("PROF_W", "1920")), int(os.environ.get("PROF_H", "1080")), int(os.environ.get("PROF_SPP", "16"))
v, f = bumpy_sphere(N)
sc = M.Scene.build(v, f, want_bvh=False)
frame = M.camera_frame((0, 0, 3), (0, 0, 0), width=W, height=H)
p = sc.render_params(frame, W, H, shader=M.SHADER_PRIMARY_SHADOW, light=(2.0, 4.0, 3.0))
for _ in range(int(os.environ.get("PROF_FRAMES", "1"))):
    img, cnt, st = sc.render_frame(p, SPP, stats=False)     # stats=True would launch the counting instantiations
print(float(img.sum()))
sc.close()
