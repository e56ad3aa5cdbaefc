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


@pytest.mark.parametrize("mesh,entry", [("cornellbox", "cornellbox_512"), ("teapot", "teapot_1080p"),
                                        ("sphere40", "sphere40_256"), ("sphere500", "sphere500_1080p")])
def test_device_builder_is_the_reference_builder(mesh, entry):
    """Against the committed fingerprints of the REFERENCE's own trees (tests/golden/golden.json, generated from
    oracle/_ref) and, for the small scenes, against the oracle's builder node by node."""
    g = T.golden()[entry]
    m = T.load_mesh(mesh)
    db = M.HostBVH.build_device(m["vertices"], m["faces"])
    nodes, idx = db.arrays()
    assert db.stats() == g["stats"]
    assert len(nodes) == g["num_nodes"] and len(idx) == g["num_indices"]
    assert T.fnv(idx) == g["indices_fnv"] and T.fnv(T.mask_leaf_axis(nodes)) == g["nodes_fnv"]
    if mesh != "sphere500":
        _, ob = T.oracle_scene(mesh)
        on, oi = ob.arrays()
        assert np.array_equal(oi, idx) and T.mask_leaf_axis(on).tobytes() == T.mask_leaf_axis(nodes).tobytes()
    db.close()


@pytest.mark.parametrize("opt", [dict(bin_size=16), dict(min_leaf=4), dict(max_depth=5), dict(cost_taabb=0.5, bin_size=128),
                                 dict(min_leaf=2, bin_size=8)])
def test_device_build_options(opt):
    m = T.load_mesh("teapot")
    same_tree(m["vertices"], m["faces"], **opt)


soup = T.soup


@pytest.mark.parametrize("name", sorted(T.build_cases()))
def test_device_builder_edge_cases_match_the_reference_tree(name):
    """Fingerprints of the reference's own trees for these inputs (tests/golden/build_golden.json)."""
    v, f = T.build_cases()[name]
    g = T.build_golden()[name]
    db = M.HostBVH.build_device(v, f)
    assert T.tree_fingerprint(*db.arrays()) == {k: g[k] for k in ("num_nodes", "nodes_fnv", "indices_fnv")}
    assert db.stats() == g["stats"]
    db.close()


DEVICE_OPTION_CASES = [(n, o) for n, o in T.build_option_cases() if o.get("min_leaf", 16) >= 2]


@pytest.mark.parametrize("name,opt", DEVICE_OPTION_CASES, ids=[T.build_option_key(n, o) for n, o in DEVICE_OPTION_CASES])
def test_device_builder_matches_the_reference_tree_under_options(name, opt):
    """The reference's own trees under non-default BVHBuildOptions (goldens made by ref_scene_build_opts)."""
    v, f = T.build_cases()[name]
    g = T.build_golden()[T.build_option_key(name, opt)]
    db = M.HostBVH.build_device(v, f, **opt)
    assert T.tree_fingerprint(*db.arrays()) == {k: g[k] for k in ("num_nodes", "nodes_fnv", "indices_fnv")}
    assert db.stats() == g["stats"]
    db.close()


@pytest.mark.parametrize("case", [
    dict(n=5000, seed=1),                                   # overlapping triangles, many straddle the planes
    dict(n=5000, seed=2, tri=0.8),                          # huge triangles: most bins shared, frequent median fallback
    dict(n=3000, seed=3, scale=1e-12),                      # extents below kEPS * 1024: bin scale 0 (bvh_accel.cc:105-112)
    dict(n=3000, seed=4, scale=1e6, offset=-3e6),           # large negative coordinates
    dict(n=4000, seed=5, opt=dict(bin_size=1024)),          # more bins than the shared-memory histogram holds
    dict(n=4000, seed=6, opt=dict(bin_size=2)),             # a single candidate plane per axis
    dict(n=4000, seed=7, opt=dict(max_depth=1)),
    dict(n=15, seed=8), dict(n=16, seed=9), dict(n=17, seed=10), dict(n=1, seed=11),
])
def test_device_build_triangle_soups(case):
    v, f = soup(case["n"], case["seed"], case.get("scale", 1.0), case.get("offset", 0.0), case.get("tri", 0.05))
    same_tree(v, f, **case.get("opt", {}))


def test_device_build_quantised_centroids():
    """Many triangles with exactly equal centroid sums and bounds (a regular grid): ties in the predicate and in the
    SAH costs; the plane / axis choice and the partition order must still follow the reference."""
    g = np.arange(65, dtype=np.float64)
    vv = np.stack([np.repeat(g, 65), np.tile(g, 65), np.zeros(65 * 65)], axis=1)
    q = np.array([[i * 65 + j, i * 65 + j + 1, (i + 1) * 65 + j] for i in range(64) for j in range(64)], np.uint32)
    q = np.concatenate([q, q[::-1]])                         # every triangle twice
    same_tree(vv, q)
    same_tree(vv[:, [2, 0, 1]].copy(), q, min_leaf=2)


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


def same_scene(m, **opt):
    """mb200_scene_build against mb200_scene_create(mb200_bvh_build): resident layout, tree and hit records."""
    v, f = m["vertices"], m["faces"]
    host = M.Scene(v, f, m.get("material_ids"), m.get("normals"), m.get("uvs"))
    dev = M.Scene.build(v, f, m.get("material_ids"), m.get("normals"), m.get("uvs"), **opt)
    try:
        hi, hp, ht = host.layout()
        di, dp, dt = dev.layout()
        assert hi == di
        assert hp.tobytes() == dp.tobytes(), "pair nodes differ"
        assert ht.tobytes() == dt.tobytes(), "triangle records differ"
        if not opt:
            assert host.nodes.tobytes() == dev.nodes.tobytes() and np.array_equal(host.indices, dev.indices)
        assert host.uses_f32_vertices() == dev.uses_f32_vertices()
        lo, hi_ = host.bounds()
        dlo, dhi = dev.bounds()
        assert np.array_equal(lo, dlo) and np.array_equal(hi_, dhi)
        c = (lo + hi_) / 2
        ext = float(np.max(hi_ - lo))
        fr = M.camera_frame(c + np.array([0.1 * ext, 0.2 * ext, 1.6 * ext]), c, width=160, height=120)
        rays = host.generate_rays_grid(fr, 0, 0, 160, 120)
        a, b = host.trace_closest(rays), dev.trace_closest(rays)
        assert a.tobytes() == b.tobytes()
        assert int((a["faceID"] != 0xFFFFFFFF).sum()) > 0
        ia, ib = host.trace_closest_full(rays), dev.trace_closest_full(rays)   # BuildIntersection: normals / uvs / materials
        for x, y in zip(ia if isinstance(ia, tuple) else (ia,), ib if isinstance(ib, tuple) else (ib,)):
            assert np.asarray(x).tobytes() == np.asarray(y).tobytes()
    finally:
        host.close()
        dev.close()


@pytest.mark.parametrize("name", ["cornellbox", "teapot", "sphere40"])
def test_device_scene_matches_host_scene(name):
    same_scene(T.load_mesh(name))


def test_device_scene_f64_records_and_empty():
    m = dict(T.load_mesh("sphere40"))
    m["vertices"] = m["vertices"] + 1e-11 * np.arange(m["vertices"].size).reshape(-1, 3)   # not float-exact -> 80-byte records
    same_scene(m)
    e = M.Scene.build(m["vertices"], m["faces"][:0])
    assert e.layout()[0]["empty"] == 1
    e.close()
    with pytest.raises(M.MallieB200Error):
        M.Scene.build(m["vertices"], m["faces"], device=99)


def test_scene_clone_same_device():
    """mb200_scene_clone onto the same GPU (the 2-GPU case is in test_gpu_multi.py): an independent, identical scene."""
    m = T.load_mesh("cornellbox")
    a = M.Scene.build(m["vertices"], m["faces"], m["material_ids"], m["normals"], m["uvs"])
    b = a.clone(0)
    ia, pa, ta = a.layout()
    ib, pb, tb = b.layout()
    assert ia == ib and pa.tobytes() == pb.tobytes() and ta.tobytes() == tb.tobytes()
    assert a.device_bytes() == b.device_bytes()
    fr = M.camera_frame((0, 0, 20), (0, 0, 0), width=96, height=96)
    rays = a.generate_rays_grid(fr, 0, 0, 96, 96)
    want = a.trace_closest_full(rays)
    a.close()                                  # the clone owns its memory
    got = b.trace_closest_full(rays)
    assert want[0].tobytes() == got[0].tobytes() and np.array_equal(want[1], got[1])
    b.close()
    with pytest.raises(M.MallieB200Error):
        M.Scene.build(m["vertices"], m["faces"]).clone(77)
