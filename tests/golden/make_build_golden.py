#!/usr/bin/env python
"""Tree fingerprints of the UNMODIFIED reference builder (BVHAccel::Build through oracle/_ref/libmallie_ref.so, default
BVHBuildOptions, and the option sets of tests/common.py::build_option_cases) for the builder edge cases of tests/common.py::build_cases -- triangle soups, degenerate extents,
tie-heavy grids, node sizes around minLeafPrimitives.  Authoring container only (needs oracle/_ref):

    python tests/golden/make_build_golden.py        -> tests/golden/build_golden.json (committed)

The host builder, the oracle's builder and the numpy model are held against these on CPU (tests/test_build_golden.py),
the device builder on the GPU (tests/test_gpu_build.py)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refbind as R  # noqa: E402
from tests import common as T  # noqa: E402

out = {}
for name, (v, f) in T.build_cases().items():
    rs = R.RefScene.from_arrays(v, f)
    rs.build()
    nodes, idx = rs.bvh()
    fp = T.tree_fingerprint(nodes, idx)
    fp["stats"] = rs.stats()
    fp["num_triangles"] = int(len(f))
    out[name] = fp
    rs.close()
    print(name, fp)
# the same builder under non-default BVHBuildOptions (ref_scene_build_opts)
cases = T.build_cases()
for name, opt in T.build_option_cases():
    v, f = cases[name]
    rs = R.RefScene.from_arrays(v, f)
    rs.build(**opt)
    nodes, idx = rs.bvh()
    fp = T.tree_fingerprint(nodes, idx)
    fp["stats"] = rs.stats()
    fp["num_triangles"] = int(len(f))
    out[T.build_option_key(name, opt)] = fp
    rs.close()
    print(T.build_option_key(name, opt), fp)
with open(os.path.join(HERE, "build_golden.json"), "w") as fp:
    json.dump(out, fp, indent=1, sort_keys=True)
