#!/usr/bin/env python
"""north_star names "Woop-packed triangles"; the production records keep Moeller-Trumbore in the reference's operation
order instead, because hit records have to be bit-identical.  This measures what an FP64 Woop record (traverse.cuh:
tri_test_woop, 96 B, three 256-bit loads) would buy and what it would cost, on the golden ray sets (development build:
make -C mallie_b200/csrc DEV=1).  Per ray set: primID mismatches against the exact kernel, hit / miss flips, the largest
relative t and absolute u / v difference where both agree on the triangle, and the closest-hit kernel time of both."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mallie_b200 as M  # noqa: E402
from tests import common as T  # noqa: E402  (meshes / golden camera set-ups only)

CASES = [("cornellbox", "cornellbox_512"), ("teapot", "teapot_1080p"), ("sphere40", "sphere40_256"), ("sphere500", "sphere500_1080p")]

print("| ray set | rays | hits | primID differs | hit/miss flips | max rel dt | max abs du | max abs dv | exact kernel ms | Woop ms |")
print("|---|---|---|---|---|---|---|---|---|---|")
for mesh, entry in CASES:
    g = T.golden()[entry]
    m = T.load_mesh(mesh)
    W, H = g["width"], g["height"]
    os.environ.pop("MB200_TRI_LAYOUT", None)
    exact = M.Scene.build(m["vertices"], m["faces"], m["material_ids"], m["normals"], m["uvs"], want_bvh=False)
    os.environ["MB200_TRI_LAYOUT"] = "woop"
    woop = M.Scene.build(m["vertices"], m["faces"], m["material_ids"], m["normals"], m["uvs"], want_bvh=False)
    os.environ.pop("MB200_TRI_LAYOUT", None)
    frame = M.camera_frame(g["eye"], g["lookat"], width=W, height=H)
    sets = [("primary " + entry, exact.generate_rays_grid(frame, 0, 0, W, H))]
    rng = np.random.default_rng(7)
    sets.append(("incoherent " + mesh, T.random_rays(rng, 500_000, *exact.bounds())))
    for name, rays in sets:
        d_rays = torch.from_numpy(np.ascontiguousarray(rays)).cuda()
        d_hits = torch.empty(len(rays) * 4, dtype=torch.float64, device="cuda")
        res = {}
        for tag, sc in (("exact", exact), ("woop", woop)):
            for _ in range(2):
                sc.trace_closest_device(d_rays.data_ptr(), len(rays), d_hits.data_ptr())
            sc.synchronize()
            sc.timing(True)
            for _ in range(5):
                sc.trace_closest_device(d_rays.data_ptr(), len(rays), d_hits.data_ptr())
            kt = sc.kernel_times()
            sc.timing(False)
            res[tag] = (d_hits.cpu().numpy().view(M.capi.HIT_DTYPE).copy(), kt["query_trace_ms"] / 5)
        a, b = res["exact"][0], res["woop"][0]
        ha, hb = a["faceID"] != 0xFFFFFFFF, b["faceID"] != 0xFFFFFFFF
        both = ha & hb
        same = both & (a["faceID"] == b["faceID"])
        relt = float(np.max(np.abs(a["t"][same] - b["t"][same]) / np.abs(a["t"][same]))) if same.any() else 0.0
        du = float(np.max(np.abs(a["u"][same] - b["u"][same]))) if same.any() else 0.0
        dv = float(np.max(np.abs(a["v"][same] - b["v"][same]))) if same.any() else 0.0
        print(f"| {name} | {len(rays)} | {int(ha.sum())} | {int((both & ~same).sum())} | {int((ha != hb).sum())} | {relt:.2e} | "
              f"{du:.2e} | {dv:.2e} | {res['exact'][1]:.3f} | {res['woop'][1]:.3f} |")
    exact.close()
    woop.close()
