#!/usr/bin/env python
"""What does one rank of an N-GPU frame cost on its own?  Renders band_index 0 of N interleaved 4-row bands of
the bench frame on ONE GPU, with and without an L2 flush before every frame (development aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mallie_b200 as M  # noqa: E402
from mallie_b200.procedural import bumpy_sphere  # noqa: E402

W, H, SPP = 1920, 1080, 16
v, f = bumpy_sphere(500)
sc = M.Scene.build(v, f, want_bvh=False)
frame = M.camera_frame((0, 0, 3), (0, 0, 0), width=W, height=H)
stream = torch.cuda.ExternalStream(sc.stream())
L, C = M.capi.lib(), M.capi.C
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for N in tuple(int(x) for x in os.environ.get("AB_N", "1,2,4,8").split(",")):
    bands = (4, N, 0) if N > 1 else None
    p = sc.render_params(frame, W, H, shader=M.SHADER_PRIMARY_SHADOW, light=(2.0, 4.0, 3.0), bands=bands, compact=N > 1)
    rows = sc.band_local_rows(p) if N > 1 else H
    d_img = torch.zeros(rows * W * 3, dtype=torch.float32, device="cuda")
    d_cnt = torch.zeros(rows * W, dtype=torch.int32, device="cuda")
    res = {}
    for do_flush in (False, True):
        ts = []
        sc.timing(True)
        for it in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                if do_flush:
                    flush.zero_()
                e0.record(stream)
                M.capi.check(L.mb200_render_frame(sc.h, C.byref(p), SPP, M.capi._p(d_img.data_ptr()), M.capi._p(d_cnt.data_ptr()), None))
                e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        kt = sc.kernel_times()
        sc.timing(False)
        res[do_flush] = (min(ts[2:]), kt["camera_trace_ms"] / 10, kt["shadow_trace_ms"] / 10)
    print(f"N={N} rows {rows}: warm L2 frame {res[False][0]:.3f} ms (cam {res[False][1]:.3f} shd {res[False][2]:.3f}) | "
          f"flushed {res[True][0]:.3f} ms (cam {res[True][1]:.3f} shd {res[True][2]:.3f}) | ideal {res[False][0] if N==1 else 0:.3f}")
