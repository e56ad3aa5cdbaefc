"""BVHAccel::Build on the device (mb200_bvh_build_device) against the host builder (mb200_bvh_build), which the CPU
suite pins to the reference's own tree (tests/test_oracle_vs_reference.py, tests/test_host_builder.py): nodes, bounds
and the triangle index order must be identical bit for bit -- the tie-breaks of the traversal depend on them."""
import time

import numpy as np
import pytest

import mallie_b200 as M
from tests import common as T

pytestmark = pytest.mark.gpu


def same_tree(v, f, **opt):
    hb = M.HostBVH.build(v, f, **opt)
    t0 = time.perf_counter()
    db = M.HostBVH.build_device(v, f, **opt)
    dt = time.perf_counter() - t0
    hn, hi = hb.arrays()
    dn, di = db.arrays()
    assert len(hn) == len(dn) and len(hi) == len(di)
    assert np.array_equal(hi, di), "triangle index order differs"
    assert hn.tobytes() == dn.tobytes(), "node array differs"
    assert hb.stats() == db.stats()
    return dt, len(hn)


@pytest.mark.parametrize("name", ["cornellbox", "teapot", "sphere40"])
def test_device_build_matches_host_small(name):
    m = T.load_mesh(name)
    same_tree(m["vertices"], m["faces"])


@pytest.mark.parametrize("opt", [dict(bin_size=16), dict(min_leaf=4), dict(max_depth=5), dict(cost_taabb=0.5, bin_size=128),
                                 dict(min_leaf=2, bin_size=8)])
def test_device_build_options(opt):
    m = T.load_mesh("teapot")
    same_tree(m["vertices"], m["faces"], **opt)


def test_device_build_degenerate_inputs():
    # all triangles identical: no plane separates them -> object-median fallback at every level
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float64)
    f = np.tile(np.array([[0, 1, 2]], np.uint32), (100, 1))
    same_tree(v, f)
    # a flat mesh (zero extent along z), fewer triangles than a leaf holds, and the empty mesh
    g = np.arange(12, dtype=np.float64)
    vv = np.stack([np.repeat(g, 12), np.tile(g, 12), np.zeros(144)], axis=1)
    q = np.array([[i * 12 + j, i * 12 + j + 1, (i + 1) * 12 + j] for i in range(11) for j in range(11)], np.uint32)
    same_tree(vv, q)
    same_tree(vv, q[:5])
    db = M.HostBVH.build_device(vv, q[:0])
    assert len(db.arrays()[0]) == 0 and len(db.arrays()[1]) == 0
    with pytest.raises(M.MallieB200Error):
        M.HostBVH.build_device(vv, q, min_leaf=1)
    with pytest.raises(M.MallieB200Error):
        M.HostBVH.build_device(vv, np.array([[0, 1, 999]], np.uint32))


def test_device_build_matches_host_1m():
    """The bench scene (1 M triangles): identical tree; the timing is printed for DESIGN.md, not asserted."""
    v, f = T.bumpy_sphere(500)
    t0 = time.perf_counter()
    M.HostBVH.build(v, f)
    host = time.perf_counter() - t0
    same_tree(v, f)                 # first call pays context + allocation warm-up
    dt, nn = same_tree(v, f)
    print(f"\n1M triangles: {nn} nodes, host build {host*1e3:.0f} ms, device build {dt*1e3:.0f} ms (incl. copies)")
