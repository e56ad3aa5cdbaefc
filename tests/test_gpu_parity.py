"""GPU parity: the CUDA path (through the C ABI) against the oracle and the golden vectors.

Bar (BASELINE.json north_star): faceID / materialID bit-exact, t/u/v within 1e-5 relative --
the FP64 kernels are expected to be, and are checked to be, BIT-IDENTICAL.
"""
import numpy as np
import pytest

import mallie_b200 as M
from mallie_b200.procedural import bumpy_sphere
from oracle import orabind as O
from tests import common as T

pytestmark = pytest.mark.gpu

SCENES = [("cornellbox", "cornellbox_512"), ("teapot", "teapot_1080p"), ("sphere40", "sphere40_256"),
          ("sphere500", "sphere500_1080p")]


def make_scene(name):
    m = T.load_mesh(name)
    return M.Scene(m["vertices"], m["faces"], m["material_ids"], m["normals"], m["uvs"])


@pytest.fixture(scope="module")
def scenes():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = make_scene(name)
        return cache[name]
    yield get
    for s in cache.values():
        s.close()


@pytest.mark.parametrize("mesh,entry", SCENES)
def test_primary_rays_match_golden_and_oracle(scenes, mesh, entry):
    g = T.golden()[entry]
    sc = scenes(mesh)
    W, H = g["width"], g["height"]
    frame = M.camera_frame(g["eye"], g["lookat"], width=W, height=H)
    for a, b in zip(frame.arrays(), T.golden_frame(g)):
        assert a.tobytes() == b.tobytes()
    rays = sc.generate_rays_grid(frame, 0, 0, W, H)          # K1 on the device
    assert T.fnv(rays) == g["rays_fnv"]
    hits, cnt = sc.trace_closest(rays, counters=True)        # K2
    mask = hits["faceID"] != 0xFFFFFFFF
    assert int(mask.sum()) == g["hits"]
    assert T.fnv(hits["faceID"]) == g["faceid_fnv"]
    assert T.fnv(np.stack([hits["t"], hits["u"], hits["v"]], 1)[mask]) == g["tuv_fnv"]
    for s in g["spots"]:
        r = hits[s["y"] * W + s["x"]]
        assert int(r["faceID"]) == s["faceID"]
        assert float(r["t"]) == float.fromhex(s["t"]) and float(r["u"]) == float.fromhex(s["u"])
    # sampled full records from the reference itself
    idx, ghits, gmask, gis = T.sample(entry)
    T.assert_hits_equal(hits[idx], ghits, f"{entry} sample")
    # the oracle on the same rays: records and traversal counters
    om, ob = T.oracle_scene(mesh)
    o = ob.trace(rays, row=W)
    T.assert_hits_equal(hits, o["hits"], entry)
    assert cnt["nodes_tested"] == o["n_node"] and cnt["tris_tested"] == o["n_tri"] and cnt["rays"] == W * H


@pytest.mark.parametrize("mesh,entry", SCENES[:3])
def test_full_intersection_records(scenes, mesh, entry):
    g = T.golden()[entry]
    sc = scenes(mesh)
    W, H = g["width"], g["height"]
    frame = M.camera_frame(g["eye"], g["lookat"], width=W, height=H)
    rays = sc.generate_rays_grid(frame, 0, 0, W, H)
    isects, mask = sc.trace_closest_full(rays)               # K2 + K3
    assert int(mask.sum()) == g["hits"]
    for f, h in g["isect_fnv"].items():
        assert T.fnv(isects[f][mask]) == h, f
    idx, ghits, gmask, gis = T.sample(entry)
    assert np.array_equal(mask[idx], gmask)
    for f in ("position", "geometricNormal", "normal", "texcoord", "f0", "f1", "f2", "faceID", "t", "u", "v"):
        assert np.ascontiguousarray(isects[f][idx][gmask]).tobytes() == np.ascontiguousarray(gis[f][gmask]).tobytes(), f


@pytest.mark.parametrize("mesh", ["cornellbox", "teapot", "sphere40"])
def test_incoherent_rays_and_occlusion(scenes, mesh):
    sc = scenes(mesh)
    om, ob = T.oracle_scene(mesh)
    bmin, bmax = sc.bounds()
    rng = np.random.default_rng(7)
    rays = T.random_rays(rng, 200_000, bmin, bmax)
    hits, cnt = sc.trace_closest(rays, counters=True)
    o = ob.trace(rays)
    T.assert_hits_equal(hits, o["hits"], mesh)
    assert cnt["nodes_tested"] == o["n_node"] and cnt["tris_tested"] == o["n_tri"]
    # occlusion: tmax drawn around the true hit distance, including exactly t (not occluded: t < tmax is strict)
    t = o["hits"]["t"].copy()
    t[~o["mask"]] = 10.0
    scale = rng.choice([0.5, 1.0, 1.0, 1.5, 1e30], size=t.shape)
    tmax = t * scale
    occ = sc.trace_occluded(rays, tmax)
    want = ob.occluded(rays, tmax)
    assert np.array_equal(occ, want)
    assert occ.sum() > 0 and (~occ).sum() > 0


def test_edge_case_rays(scenes):
    """Axis-parallel directions (1/0 = inf, 0*inf = NaN in the slab test), -0.0 components, rays starting
    inside / on box planes, grazing rays along mesh edges and through vertices, zero direction."""
    sc = scenes("cornellbox")
    om, ob = T.oracle_scene("cornellbox")
    rays = T.edge_case_rays("cornellbox")
    with np.errstate(all="ignore"):
        hits = sc.trace_closest(rays)
        o = ob.trace(rays)
    T.assert_hits_equal(hits, o["hits"], "edge cases")
    assert o["mask"].sum() > 100


def test_empty_inputs_and_empty_scene(scenes):
    sc = scenes("cornellbox")
    assert sc.trace_closest(np.zeros((0, 6))).shape == (0,)
    assert sc.trace_occluded(np.zeros((0, 6)), np.zeros(0)).shape == (0,)
    empty = M.Scene(np.zeros((0, 3)), np.zeros((0, 3), np.uint32))
    rays = np.array([[0, 0, 5, 0, 0, -1.0], [1, 2, 3, 0, 1, 0]])
    h = empty.trace_closest(rays)
    assert np.all(h["faceID"] == 0xFFFFFFFF) and np.all(h["t"] == np.finfo(np.float64).max)
    assert not empty.trace_occluded(rays, np.array([1e30, 1e30])).any()
    empty.close()
    # ragged sizes around the warp / block granularity
    g = T.golden()["cornellbox_512"]
    frame = M.camera_frame(g["eye"], g["lookat"], width=512, height=512)
    rays = sc.generate_rays_grid(frame, 0, 200, 512, 210)
    om, ob = T.oracle_scene("cornellbox")
    for n in (1, 31, 32, 33, 127, 129, 4097):
        T.assert_hits_equal(sc.trace_closest(rays[:n]), ob.trace(rays[:n])["hits"], f"n={n}")


def test_f64_vertex_path_and_single_leaf():
    """scene_scale != 1 makes vertices non-float-representable -> 80-byte double triangle records."""
    m = T.load_mesh("sphere40")
    v = m["vertices"] * 1.1
    sc = M.Scene(v, m["faces"])
    assert not sc.uses_f32_vertices()
    om = O.Mesh(v, m["faces"])
    ob = O.BVH.build(om)
    frame = M.camera_frame((0.3, 0.2, 3), (0, 0, 0), width=300, height=200)
    rays = sc.generate_rays_grid(frame, 0, 0, 300, 200)
    T.assert_hits_equal(sc.trace_closest(rays), ob.trace(rays)["hits"], "f64 records")
    sc.close()
    # a mesh smaller than minLeafPrimitives: the root is a leaf
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0.5]], np.float64)
    f = np.array([[0, 1, 2], [1, 3, 2]], np.uint32)
    sc = M.Scene(v, f)
    ob = O.BVH.build(O.Mesh(v, f))
    rng = np.random.default_rng(0)
    rays = np.concatenate([rng.uniform(-0.5, 1.5, (5000, 2)), np.full((5000, 1), 3.0),
                           np.tile([0, 0, -1.0], (5000, 1))], axis=1)
    T.assert_hits_equal(sc.trace_closest(rays), ob.trace(rays)["hits"], "root leaf")
    sc.close()


def chain_bvh(v, f, pad=2.2737367544323206e-13):
    """A hand-made, maximally unbalanced but valid BVH: branch k = {leaf(triangle k), branch k+1}."""
    n = len(f)
    nodes = np.zeros(2 * n - 1, M.capi.NODE_DTYPE)
    tri_lo = v[f].min(axis=1) - pad
    tri_hi = v[f].max(axis=1) + pad
    suffix_lo = np.minimum.accumulate(tri_lo[::-1], axis=0)[::-1]
    suffix_hi = np.maximum.accumulate(tri_hi[::-1], axis=0)[::-1]
    for k in range(n - 1):                      # branch k at index 2k, its leaf at 2k+1
        b = nodes[2 * k]
        b["bmin"], b["bmax"], b["flag"], b["axis"] = suffix_lo[k], suffix_hi[k], 0, k % 3
        b["data"] = (2 * k + 1, 2 * k + 2)
        l = nodes[2 * k + 1]
        l["bmin"], l["bmax"], l["flag"], l["data"] = tri_lo[k], tri_hi[k], 1, (1, k)
    l = nodes[2 * n - 2]
    l["bmin"], l["bmax"], l["flag"], l["data"] = tri_lo[n - 1], tri_hi[n - 1], 1, (1, n - 1)
    return nodes, np.arange(n, dtype=np.uint32)


def test_deep_tree_uses_big_stack():
    """A 300-level chain needs the 512-entry-stack kernels (the reference's kMaxStackDepth, bvh_accel.cc:548);
    overlapping triangles along the chain make the far-child stack actually fill up."""
    rng = np.random.default_rng(11)
    n = 300
    c = rng.uniform(-1, 1, (n, 1, 3)) * np.array([1.0, 1.0, 0.2])
    v = (c + rng.normal(0, 2.0, (n, 3, 3))).reshape(-1, 3).astype(np.float32).astype(np.float64)
    f = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    nodes, idx = chain_bvh(v, f)
    sc = M.Scene(v, f, nodes=nodes, indices=idx)
    ob = O.BVH.from_arrays(nodes, idx, O.Mesh(v, f))
    rays = T.random_rays(rng, 50_000, v.min(0), v.max(0))
    hits, cnt = sc.trace_closest(rays, counters=True)
    o = ob.trace(rays)
    T.assert_hits_equal(hits, o["hits"], "deep chain")
    assert cnt["nodes_tested"] == o["n_node"] and cnt["tris_tested"] == o["n_tri"]
    assert cnt["max_stack"] > 64, cnt
    tmax = np.where(o["mask"], o["hits"]["t"] * rng.choice([0.9, 1.0, 1.1], len(rays)), 5.0)
    assert np.array_equal(sc.trace_occluded(rays, tmax), ob.occluded(rays, tmax))
    sc.close()
    # deeper than the reference's own stack could handle -> rejected, not a hang
    n = 600
    v = rng.uniform(-1, 1, (3 * n, 3))
    f = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    nodes, idx = chain_bvh(v, f)
    with pytest.raises(M.MallieB200Error):
        M.Scene(v, f, nodes=nodes, indices=idx)


def test_tree_with_empty_leaves_min_leaf_primitives_1():
    """minLeafPrimitives = 1: the reference's tree hangs a 256-deep chain under every triangle, each level an EMPTY left
    leaf (a point box at the triangle's first vertex) and the triangle again.  The traversal must visit and count those
    boxes as the reference does (hits and counters against the oracle walking the very same tree; the tree itself is
    pinned to the reference's in tests/test_build_golden.py)."""
    rng = np.random.default_rng(5)
    for name in ("soup_17", "identical_100"):
        v, f = T.build_cases()[name]
        hb = M.HostBVH.build(v, f, min_leaf=1)
        nodes, idx = hb.arrays()
        assert hb.stats()["maxTreeDepth"] == 256 and ((nodes["flag"] == 1) & (nodes["data"][:, 0] == 0)).sum() > 1000
        sc = M.Scene(v, f, nodes=nodes, indices=idx)
        ob = O.BVH.from_arrays(nodes, idx, O.Mesh(v, f))
        rays = T.random_rays(rng, 20_000, v.min(0) - 0.1, v.max(0) + 0.1)
        # plus rays aimed at first vertices, i.e. straight through the point boxes of the empty leaves
        tgt = v[f[rng.integers(0, len(f), 5000), 0]]
        org = tgt + rng.normal(0, 1.0, tgt.shape)
        d = tgt - org
        rays = np.concatenate([rays, np.concatenate([org, d / np.linalg.norm(d, axis=1, keepdims=True)], axis=1)])
        hits, cnt = sc.trace_closest(rays, counters=True)
        o = ob.trace(rays)
        T.assert_hits_equal(hits, o["hits"], name)
        assert cnt["nodes_tested"] == o["n_node"] and cnt["tris_tested"] == o["n_tri"]
        assert o["mask"].any()
        tmax = np.where(o["mask"], o["hits"]["t"] * rng.choice([0.9, 1.0, 1.1], len(rays)), 5.0)
        assert np.array_equal(sc.trace_occluded(rays, tmax), ob.occluded(rays, tmax))
        sc.close()
        hb.close()


def test_malformed_bvh_is_rejected():
    m = T.load_mesh("sphere40")
    hb = M.HostBVH.build(m["vertices"], m["faces"])
    nodes, idx = hb.arrays()
    bad = nodes.copy()
    bad["data"][0][0] = 0                     # root's child points at the root: a cycle
    with pytest.raises(M.MallieB200Error):
        M.Scene(m["vertices"], m["faces"], nodes=bad, indices=idx)
    bad = nodes.copy()
    leaf = np.nonzero(bad["flag"] == 1)[0][0]
    bad["data"][leaf][0] = 10 ** 9            # leaf range past the index array
    with pytest.raises(M.MallieB200Error):
        M.Scene(m["vertices"], m["faces"], nodes=bad, indices=idx)
    bidx = idx.copy()
    bidx[3] = len(m["faces"]) + 5
    with pytest.raises(M.MallieB200Error):
        M.Scene(m["vertices"], m["faces"], nodes=nodes, indices=bidx)


def test_ten_million_triangles_4k_full_size():
    """BASELINE.json configs[4] geometry at full size: ~10 M-triangle bumpy sphere (N = 1581 -> 9 998 244
    triangles), 3840x2160 un-jittered primaries.  Pins: the hit count and the node / triangle visit averages the
    UNMODIFIED reference produced for this exact ray set (2 862 379 hits, 28.29 nodes/ray,
    7.91 tris/ray, 1 932 723 nodes, depth 35), and bit-exact hit records against the oracle on 40 000 sampled rays."""
    W, H = 3840, 2160
    v, f = bumpy_sphere(1581)
    assert len(f) == 9998244
    hb = M.HostBVH.build(v, f)
    st = hb.stats()
    assert st["numLeafNodes"] + st["numBranchNodes"] == 1932723 and st["maxTreeDepth"] == 35
    nodes, idx = hb.arrays()
    sc = M.Scene(v, f, nodes=nodes, indices=idx)
    fg = M.camera_frame((0, 0, 3), (0, 0, 0), width=W, height=H)
    rays = sc.generate_rays_grid(fg, 0, 0, W, H)
    hits, cnt = sc.trace_closest(rays, counters=True)
    mask = hits["faceID"] != 0xFFFFFFFF
    assert int(mask.sum()) == 2862379
    n = W * H
    assert abs(cnt["nodes_tested"] / n - 28.29) < 0.006 and abs(cnt["tris_tested"] / n - 7.91) < 0.006
    assert cnt["max_stack"] <= 37
    # occlusion: a ray is occluded iff the closest hit lies before tmax
    sel = np.random.default_rng(5).choice(n, 40000, replace=False)
    sel.sort()
    tmax = np.where(mask[sel], hits["t"][sel] * 1.5, 1.0)
    tmax[::2] = np.where(mask[sel][::2], hits["t"][sel][::2] * 0.5, 1.0)
    occ = sc.trace_occluded(rays[sel], tmax)
    assert np.array_equal(occ, mask[sel] & (hits["t"][sel] < tmax))
    # oracle on the sample (same tree: the oracle BVH is loaded from the arrays the host builder produced,
    # which test_abi pins to the reference builder on the smaller scenes)
    ob = O.BVH.from_arrays(nodes, idx, O.Mesh(v, f))
    o = ob.trace(rays[sel], row=4000)
    T.assert_hits_equal(hits[sel], o["hits"], "10M sample")
    sc.close()
    hb.close()
