#!/usr/bin/env python
"""How far are path-traced frames and panorama rays from the oracle's (glibc) bits?  The only arithmetic the device does not
share with the reference's host is libm (acos / sin / cos / atan2); this prints the fraction of bit-identical pixels and
rays (development aid; the bars the tests hold are in tests/test_gpu_render.py / test_gpu_fullsize.py)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mallie_b200 as M  # noqa: E402
from oracle import orabind as O  # noqa: E402
from tests import common as T  # noqa: E402

CORES = len(os.sched_getaffinity(0))
for name, eye, lookat, W, H, L in (("cornellbox", (0, 0, 20), (0, 0, 0), 1920, 1080, 5), ("cornellbox", (0, 0, 20), (0, 0, 0), 512, 512, 16),
                                   ("teapot", (5, 40, 150), (5, 40, 0), 960, 540, 8)):
    m = T.load_mesh(name)
    sc = M.Scene.build(m["vertices"], m["faces"], m["material_ids"], m["normals"], m["uvs"], want_bvh=False)
    om, ob = T.oracle_scene(name)
    fg = M.camera_frame(eye, lookat, width=W, height=H)
    fo = O.camera_frame(eye, lookat, width=W, height=H)
    pl = M.plane_from_bounds(*sc.bounds())
    p = sc.render_params(fg, W, H, shader=M.SHADER_PATHTRACE, max_path_length=L, pass_index=7, plane=pl)
    img, cnt, st = sc.render_pass(p)
    oimg, _, oc = ob.render_pass(fo, W, H, rng_mode=1, pass_index=7, skip_zombies=1, shader=0, max_path_length=L, plane=pl,
                                 nthreads=CORES)
    same = (img.view(np.uint32) == oimg.view(np.uint32)).all(axis=2)
    print(f"{name} {W}x{H} max_path_length {L}: {int((~same).sum())} of {W * H} pixels differ ({1 - same.mean():.2e}); "
          f"rays {st['primary_rays'] + st['bounce_rays']} vs {oc['trace_calls']}", flush=True)
    sc.close()

# panorama rays
m = T.load_mesh("sphere40")
sc = M.Scene(m["vertices"], m["faces"])
W, H = 2048, 1024
rng = np.random.default_rng(3)
px, py = rng.uniform(0, W, 200000), rng.uniform(0, H, 200000)
fr = O.camera_frame((0.3, -0.2, 2.5), (0, 0, 0), width=W, height=H)
for stereo in (False, True):
    got = sc.generate_rays_env((0.3, -0.2, 2.5), W, H, px, py, stereo=stereo)
    want = O.generate_env(fr[0], W, H, px, py, stereo=stereo)
    same = (got.view(np.uint64) == want.view(np.uint64)).all(axis=1)
    print(f"env rays stereo={stereo}: {int((~same).sum())} of {len(px)} rays differ ({1 - same.mean():.2e}), max abs diff {np.abs(got - want).max():.2e}")
sc.close()
