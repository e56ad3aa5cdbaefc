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
# round 2: frames large enough for the row cut + chunked copy-back and the longest-rays-first order (>= 1024 tiles, twice
# the same layout), the device LDR resolves, the measured-peaks probes
Wb, Hb = 512, 264
fb = M.camera_frame((0.3, 0.2, 3.0), (0, 0, 0), width=Wb, height=Hb)
pb = sc.render_params(fb, Wb, Hb, shader=M.SHADER_PRIMARY_SHADOW, light=(2, 4, 3))
for _ in range(2):
    img, cntb, _ = sc.render_frame(pb, 2)
for mode in (M.capi.LDR_RGB8_LINEAR, M.capi.LDR_BGRA8_GAMMA22):
    sc.resolve_ldr(img, cntb, Wb, Hb, mode)
    sc.render_frame_ldr(pb, 2, mode)
pp = sc.render_params(fb, Wb, Hb, shader=M.SHADER_PATHTRACE, max_path_length=5, plane=M.plane_from_bounds(*sc.bounds()))
sc.render_frame(pp, 2)
p = sc.render_params(fg, W, H, shader=M.SHADER_PRIMARY_SHADOW, light=(2, 4, 3), bands=(4, 3, 1), compact=True)
sc.render_pass(p)
sc.close()
# the device builder: a mesh large enough for single-segment CTAs (block-wide reductions, shared-memory histograms)
# at the top levels and many small segments (per-element atomics) below; then the device-side layout and a clone
v2, f2 = bumpy_sphere(40)
db = M.HostBVH.build_device(v2, f2)
hb = M.HostBVH.build(v2, f2)
assert db.arrays()[0].tobytes() == hb.arrays()[0].tobytes() and (db.arrays()[1] == hb.arrays()[1]).all()
ds = M.Scene.build(v2 + 1e-11, f2, min_leaf=4, bin_size=16)          # f64 records
ds2 = M.Scene.build(v2, f2)
cl = ds2.clone(0)
assert cl.trace_closest(rays).tobytes() == ds2.trace_closest(rays).tobytes()
for s in (ds, ds2, cl):
    s.close()
print("sanitize workload done:", int((hits["faceID"] != 0xFFFFFFFF).sum()), "hits", cnt)
