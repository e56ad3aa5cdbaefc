#!/usr/bin/env python
"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / initcheck): every kernel of the library
once, on a scene small enough to finish under the sanitizer.  Usage: compute-sanitizer --tool memcheck python tools/sanitize.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mallie_b200 as M  # noqa: E402
from mallie_b200.procedural import bumpy_sphere  # noqa: E402

W, H = 96, 64
v, f = bumpy_sphere(12)
sc = M.Scene(v, f)
fg = M.camera_frame((0.3, 0.2, 3.0), (0, 0, 0), width=W, height=H)
rays = sc.generate_rays_grid(fg, 0, 0, W, H)
hits, cnt = sc.trace_closest(rays, counters=True)
full = sc.trace_closest_full(rays)
occ = sc.trace_occluded(rays, np.full(len(rays), 2.5))
env = sc.generate_rays_env((0, 0, 0), W, H, np.arange(50.0), np.arange(50.0), stereo=True)
for shader, kw in ((M.SHADER_PRIMARY_SHADOW, dict(light=(2, 4, 3))), (M.SHADER_PATHTRACE, dict(max_path_length=6)),
                   (M.SHADER_PRIMARY_ONLY, {}), (M.SHADER_PATHTRACE_ENV, dict(camera_mode=M.CAMERA_ENV))):
    p = sc.render_params(fg, W, H, shader=shader, plane=M.plane_from_bounds(*sc.bounds()), **kw)
    sc.render_frame(p, 3)
p = sc.render_params(fg, W, H, shader=M.SHADER_PRIMARY_SHADOW, light=(2, 4, 3), step=4)
sc.render_pass(p)
p = sc.render_params(fg, W, H, shader=M.SHADER_PRIMARY_SHADOW, light=(2, 4, 3), bands=(4, 3, 1), compact=True)
sc.render_pass(p)
sc.close()
print("sanitize workload done:", int((hits["faceID"] != 0xFFFFFFFF).sum()), "hits", cnt)
