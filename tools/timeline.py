#!/usr/bin/env python
"""The bench frame's launches as the timing hooks see them (MB200_TIMING_DUMP=1): per launch the time its stream reached
it and the time it completed, ms since the frame's first event (development aid)."""
import os
import sys

os.environ["MB200_TIMING_DUMP"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mallie_b200 as M  # noqa: E402
from mallie_b200.procedural import bumpy_sphere  # noqa: E402
import torch  # noqa: E402

W, H, SPP = 1920, 1080, 16
v, f = bumpy_sphere(500)
sc = M.Scene.build(v, f, want_bvh=False)
frame = M.camera_frame((0, 0, 3), (0, 0, 0), width=W, height=H)
L, C = M.capi.lib(), M.capi.C
d_img = torch.zeros(H * W * 3, dtype=torch.float32, device="cuda")
d_cnt = torch.zeros(H * W, dtype=torch.int32, device="cuda")
p = sc.render_params(frame, W, H, shader=M.SHADER_PRIMARY_SHADOW, light=(2.0, 4.0, 3.0))
for _ in range(3):
    M.capi.check(L.mb200_render_frame(sc.h, C.byref(p), SPP, M.capi._p(d_img.data_ptr()), M.capi._p(d_cnt.data_ptr()), None))
sc.synchronize()
sc.timing(True)
M.capi.check(L.mb200_render_frame(sc.h, C.byref(p), SPP, M.capi._p(d_img.data_ptr()), M.capi._p(d_cnt.data_ptr()), None))
sc.synchronize()
print(sc.kernel_times())
