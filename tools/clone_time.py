#!/usr/bin/env python
"""Per-GPU replicas for the multi-GPU frame: a second host upload (mb200_scene_create) vs a device-to-device copy
(mb200_scene_clone).  Needs 2 GPUs.   python tools/clone_time.py [N]    N = bumpy_sphere resolution"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mallie_b200 as M  # noqa: E402
from mallie_b200.procedural import bumpy_sphere  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1581
v, f = bumpy_sphere(n)
M.Scene.build(v[:30], f[:1] * 0, device=0).close()
M.Scene.build(v[:30], f[:1] * 0, device=1).close()          # both contexts up
first = M.Scene.build(v, f, device=0, want_bvh=True)
first.clone(1).close()                                        # peer mapping set up
best_c = best_u = 1e30
for _ in range(3):
    t0 = time.perf_counter(); c = first.clone(1); best_c = min(best_c, time.perf_counter() - t0)
    t0 = time.perf_counter(); u = M.Scene(v, f, nodes=first.nodes, indices=first.indices, device=1); best_u = min(best_u, time.perf_counter() - t0)
    same = c.layout()[1].tobytes() == u.layout()[1].tobytes() and c.layout()[2].tobytes() == u.layout()[2].tobytes()
    c.close(); u.close()
print(f"{len(f)} triangles, {first.device_bytes()/1e6:.0f} MB resident: replica by host relayout + upload {best_u*1e3:.1f} ms, "
      f"by mb200_scene_clone {best_c*1e3:.1f} ms ({first.device_bytes()/best_c/1e9:.0f} GB/s), identical {same}")
first.close()
