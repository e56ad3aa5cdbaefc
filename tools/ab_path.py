#!/usr/bin/env python
"""Path-traced frames with and without the bounce-queue regrouping (MB200_SORT_BOUNCES, read once per process, so this
script is run once per setting): BASELINE config 3 (cornell box, 1080p, max_path_length 5) and a path-traced frame of
the 1 M-triangle sphere over its ground plane.  Prints ms per frame, per-class kernel times and an image digest
(the digest must not depend on the setting)."""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mallie_b200 as M  # noqa: E402
from mallie_b200.procedural import bumpy_sphere  # noqa: E402
from tests import common as T  # noqa: E402

W, H = 1920, 1080
L, C = M.capi.lib(), M.capi.C
tag = "sort=" + os.environ.get("MB200_SORT_BOUNCES", "default")


def run(name, sc, frame, spp, **kw):
    stream = torch.cuda.ExternalStream(sc.stream())
    p = sc.render_params(frame, W, H, shader=M.SHADER_PATHTRACE, **kw)
    d_img = torch.zeros(W * H * 3, dtype=torch.float32, device="cuda")
    d_cnt = torch.zeros(W * H, dtype=torch.int32, device="cuda")
    st = M.capi.RenderStats()

    def go(stats=None):
        M.capi.check(L.mb200_render_frame(sc.h, C.byref(p), spp, M.capi._p(d_img.data_ptr()), M.capi._p(d_cnt.data_ptr()), stats))

    for _ in range(2):
        go()
    sc.synchronize()
    sc.timing(True)
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        go()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    kt = {k: round(v / 5, 3) for k, v in sc.kernel_times().items() if k.endswith("_ms") and v}
    sc.timing(False)
    go(C.byref(st))
    sc.synchronize()
    rays = st.primary_rays + st.bounce_rays
    dig = hashlib.blake2b(d_img.cpu().numpy().tobytes(), digest_size=8).hexdigest()
    print(f"[{tag}] {name}: {min(ts):.3f} ms (median {float(np.median(ts)):.3f}) {rays / min(ts) / 1e3:.0f} Mrays/s "
          f"({st.primary_rays} camera + {st.bounce_rays} bounce rays) {kt} img {dig}")


m = T.load_mesh("cornellbox")
sc = M.Scene(m["vertices"], m["faces"], m["material_ids"], m["normals"], m["uvs"])
run("config 3: cornell box 1080p 16 spp max_path_length 5", sc, M.camera_frame((0, 0, 20), (0, 0, 0), width=W, height=H), 16,
    max_path_length=5)
sc.close()
v, f = bumpy_sphere(500)
sc = M.Scene.build(v, f, want_bvh=False)
pl = M.plane_from_bounds(*sc.bounds())
run("1 M-triangle sphere + plane, 1080p 4 spp max_path_length 5", sc, M.camera_frame((0, 0, 3), (0, 0, 0), width=W, height=H), 4,
    max_path_length=5, plane=pl)
sc.close()
