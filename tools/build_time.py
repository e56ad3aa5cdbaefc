#!/usr/bin/env python
"""Scene set-up (BVHAccel::Build + upload) on the host vs on the device, for the bench scenes.
  [MB200_BUILD_TIMING=1] python tools/build_time.py [N ...]     N = bumpy_sphere resolution (500 -> 1 M triangles)
Prints a markdown table (committed as profiles/r1_build.md)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mallie_b200 as M  # noqa: E402
from mallie_b200.procedural import bumpy_sphere  # noqa: E402


def best(fn, reps=3):
    t, out = 1e30, None
    for _ in range(reps):
        if out is not None and hasattr(out, "close"):
            out.close()
        t0 = time.perf_counter()
        out = fn()
        t = min(t, time.perf_counter() - t0)
    return t * 1e3, out


rows = []
for n in [int(a) for a in sys.argv[1:]] or [500]:
    v, f = bumpy_sphere(n)
    M.Scene.build(v[:30], f[:1] * 0).close()          # context
    t_hb, hb = best(lambda: M.HostBVH.build(v, f))
    nodes, idx = hb.arrays()
    t_up, hs = best(lambda: M.Scene(v, f, nodes=nodes, indices=idx))
    t_db, db = best(lambda: M.HostBVH.build_device(v, f))
    t_ds, ds = best(lambda: M.Scene.build(v, f, want_bvh=False))
    t_dsb, dsb = best(lambda: M.Scene.build(v, f, want_bvh=True))
    same = (hs.layout()[1].tobytes() == ds.layout()[1].tobytes() and hs.layout()[2].tobytes() == ds.layout()[2].tobytes()
            and nodes.tobytes() == dsb.nodes.tobytes() and (idx == dsb.indices).all() and nodes.tobytes() == db.arrays()[0].tobytes())
    rows.append((len(f), len(nodes), t_hb, t_up, t_db, t_ds, t_dsb, same))
    for s in (hs, ds, dsb):
        s.close()
    print(rows[-1], flush=True)

print(f"\n| triangles | nodes | host build ({os.cpu_count()} threads) | host relayout + upload | host total | device build, tree downloaded "
      "| device build + layout (mb200_scene_build) | same, tree downloaded too | identical |")
print("|---|---|---|---|---|---|---|---|---|")
for r in rows:
    print(f"| {r[0]} | {r[1]} | {r[2]:.1f} ms | {r[3]:.1f} ms | {r[2]+r[3]:.1f} ms | {r[4]:.1f} ms | {r[5]:.1f} ms | {r[6]:.1f} ms | {r[7]} |")
