"""The level-synchronous build ALGORITHM (tests/lsbuild_model.py, the numpy model of device/bvh_build_gpu.cu) against the
host builder and the reference's tree fingerprints -- runs without a GPU.  The CUDA implementation itself is checked in
tests/test_gpu_build.py."""
import numpy as np
import pytest

import mallie_b200 as M
from tests import common as T
from tests import lsbuild_model as L


@pytest.mark.parametrize("mesh,entry", [("cornellbox", "cornellbox_512"), ("teapot", "teapot_1080p"), ("sphere40", "sphere40_256")])
def test_level_synchronous_build_is_the_reference_build(mesh, entry):
    m = T.load_mesh(mesh)
    nodes, idx = L.build(m["vertices"], m["faces"])
    g = T.golden()[entry]
    assert len(nodes) == g["num_nodes"]
    assert T.fnv(idx) == g["indices_fnv"] and T.fnv(T.mask_leaf_axis(nodes)) == g["nodes_fnv"]
    hn, hi = M.HostBVH.build(m["vertices"], m["faces"]).arrays()
    assert np.array_equal(hi, idx) and T.mask_leaf_axis(hn).tobytes() == T.mask_leaf_axis(nodes).tobytes()


@pytest.mark.parametrize("opt,kw", [(dict(min_leaf=4), dict(min_leaf=4)), (dict(nb=8, taabb=0.5), dict(bin_size=8, cost_taabb=0.5)),
                                    (dict(max_depth=3), dict(max_depth=3))])
def test_level_synchronous_build_options_and_soup(opt, kw):
    rng = np.random.default_rng(12)
    c = rng.uniform(-1, 1, (700, 1, 3))
    v = (c + rng.uniform(-0.2, 0.2, (700, 3, 3))).reshape(-1, 3).astype(np.float32).astype(np.float64)
    f = np.arange(2100, dtype=np.uint32).reshape(700, 3)
    nodes, idx = L.build(v, f, **opt)
    hn, hi = M.HostBVH.build(v, f, **kw).arrays()
    assert np.array_equal(hi, idx) and T.mask_leaf_axis(hn).tobytes() == T.mask_leaf_axis(nodes).tobytes()
    # identical triangles: the object-median fallback at every level
    v1 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float64)
    f1 = np.tile(np.array([[0, 1, 2]], np.uint32), (70, 1))
    nodes, idx = L.build(v1, f1, **opt)
    hn, hi = M.HostBVH.build(v1, f1, **kw).arrays()
    assert np.array_equal(hi, idx) and T.mask_leaf_axis(hn).tobytes() == T.mask_leaf_axis(nodes).tobytes()
