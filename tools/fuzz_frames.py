#!/usr/bin/env python
"""Randomised frames (development aid, GPU): random image sizes (not multiples of the 8x4 tiles), sample counts, first
pass, shader (primary+shadow, path tracing with a random max_path_length and plane), rectangle, row bands
(compact or in place) -- through mb200_render_frame into host buffers, against the oracle's passes accumulated on the CPU.
MB200_FRAME_BATCH_ITEMS (read once per process) is set small so that even these frames are cut into many batches: the row
cut, the chunk reports, the copy-back by chunks and the longest-rays-first order of repeated frames are all exercised.

    [MB200_FRAME_BATCH_ITEMS=30000] python tools/fuzz_frames.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np

os.environ.setdefault("MB200_FRAME_BATCH_ITEMS", "30000")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mallie_b200 as M  # noqa: E402
from oracle import orabind as O  # noqa: E402
from tests import common as T  # noqa: E402

VIEWS = {"cornellbox": ((0, 0, 20), (0, 0, 0), (0.0, 8.0, 5.0)), "teapot": ((5, 40, 150), (5, 40, 0), (100.0, 200.0, 150.0)),
         "sphere40": ((0.2, 0.1, 3.0), (0, 0, 0), (2.0, 4.0, 3.0))}


def run(budget, seed, verbose=True):
    rng = np.random.default_rng(seed)
    scenes = {}
    for name in VIEWS:
        m = T.load_mesh(name)
        scenes[name] = (M.Scene.build(m["vertices"], m["faces"], m["material_ids"], m["normals"], m["uvs"]), T.oracle_scene(name)[1])
    t_end = time.time() + budget
    frames = 0
    while time.time() < t_end:
        name = str(rng.choice(list(VIEWS)))
        sc, ob = scenes[name]
        eye, lookat, light = VIEWS[name]
        W, H = int(rng.integers(17, 420)), int(rng.integers(9, 300))
        spp, pass0 = int(rng.integers(1, 7)), int(rng.integers(0, 30))
        shader = int(rng.choice([M.SHADER_PRIMARY_SHADOW, M.SHADER_PATHTRACE]))
        kw, okw = {}, {}
        if shader == M.SHADER_PATHTRACE:
            L = int(rng.integers(1, 7))
            kw["max_path_length"], okw["max_path_length"] = L, L
            if rng.random() < 0.5:
                kw["plane"] = M.plane_from_bounds(*sc.bounds())
                nodes, _ = ob.arrays()
                okw["plane"] = O.plane_from_bbox(nodes[0]["bmin"], nodes[0]["bmax"])
        tile = bands = None
        compact = False
        r = rng.random()
        if r < 0.3:
            x0, y0 = int(rng.integers(0, W - 1)), int(rng.integers(0, H - 1))
            tile = (x0, y0, int(rng.integers(x0 + 1, W + 1)), int(rng.integers(y0 + 1, H + 1)))
        elif r < 0.6:
            cnt = int(rng.integers(2, 6))
            bands = (int(rng.choice([4, 8, 12])), cnt, int(rng.integers(0, cnt)))
            compact = bool(rng.random() < 0.5)
        tag = f"{name} {W}x{H} spp {spp} pass0 {pass0} shader {shader} {kw.get('max_path_length', '')} plane {'plane' in kw} tile {tile} bands {bands} compact {compact}"
        fg = M.camera_frame(eye, lookat, width=W, height=H)
        fo = O.camera_frame(eye, lookat, width=W, height=H)
        p = sc.render_params(fg, W, H, shader=shader, light=light, pass_index=pass0, tile=tile, bands=bands, compact=compact, **kw)
        want = np.zeros((H, W, 3), np.float32)
        oshader = {M.SHADER_PRIMARY_SHADOW: 1, M.SHADER_PATHTRACE: 0}[shader]
        for k in range(spp):
            oimg, _, _ = ob.render_pass(fo, W, H, rng_mode=1, pass_index=pass0 + k, skip_zombies=1, shader=oshader, light=light,
                                        tile=tile, **okw)
            want += oimg
        wcnt = np.zeros((H, W), np.int32)
        x0, y0, x1, y1 = tile if tile else (0, 0, W, H)
        wcnt[y0:y1, x0:x1] = spp
        if bands:
            br, bc, bi = bands
            mine = ((np.arange(H) // br) % bc) == bi
            if compact:
                want, wcnt = want[mine], wcnt[mine]
            else:
                want[~mine], wcnt[~mine] = 0, 0
        for rep in range(2):                    # the second frame of a layout runs with the longest-rays-first order
            img, cnt, st = sc.render_frame(p, spp)
            assert np.array_equal(cnt, wcnt), "count: " + tag
            if want.size == 0:                  # a rank without rows
                assert img.size == 0, "empty band: " + tag
            elif shader == M.SHADER_PATHTRACE:
                same = (img.view(np.uint32) == want.view(np.uint32)).all(axis=2)
                assert same.mean() >= 0.9995, f"{1 - same.mean():.2e} of the pixels differ: " + tag
            else:
                assert img.tobytes() == want.tobytes(), "image: " + tag
        frames += 1
        if verbose:
            print("ok " + tag, flush=True)
    for sc, _ in scenes.values():
        sc.close()
    return frames


if __name__ == "__main__":
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    print(f"FUZZ FRAMES OK: {run(budget, seed)} frame configurations, seed {seed}, MB200_FRAME_BATCH_ITEMS={os.environ['MB200_FRAME_BATCH_ITEMS']}")
