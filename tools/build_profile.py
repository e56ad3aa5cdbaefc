#!/usr/bin/env python
"""One device-side scene build, for `ncu --metrics gpu__time_duration.sum` (per-kernel times of the builder).
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/build_launches.csv python tools/build_profile.py [N]
  python tools/build_profile.py --summarise gpurun_out/build_launches.csv      # aggregate per kernel (markdown)"""
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 2 and sys.argv[1] == "--summarise":
    rows = list(csv.reader(l for l in open(sys.argv[2]) if l.startswith('"')))
    head = rows[0]
    k_name, k_val, k_unit = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        v = float(r[k_val].replace(",", ""))
        v = v / 1e3 if r[k_unit] in ("ns", "nsecond") else v * (1e3 if r[k_unit] in ("ms", "msecond") else 1.0)
        a = agg.setdefault(r[k_name].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    print(f"{sum(a[0] for a in agg.values())} launches, {total/1e3:.3f} ms of kernel time (ncu: serialised, cold caches)\n")
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {us:.1f} | {us/total:.1%} |")
    sys.exit(0)

import mallie_b200 as M  # noqa: E402
from mallie_b200.procedural import bumpy_sphere  # noqa: E402

v, f = bumpy_sphere(int(sys.argv[1]) if len(sys.argv) > 1 else 500)
M.Scene.build(v, f, want_bvh=False).close()
print("built", len(f), "triangles")
