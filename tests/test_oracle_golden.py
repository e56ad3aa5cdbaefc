"""The oracle (oracle/mallie_oracle.c) against the golden vectors generated from the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only.

Everything here is bit-exact: BVH node/index arrays (FNV-1a-64), camera frames, ray sets, hit records,
BuildIntersection fields, and the deterministic OMP_NUM_THREADS=1 Render() images.
"""
import numpy as np
import pytest

from oracle import orabind as O
from tests import common as T

ENTRIES = [("cornellbox", "cornellbox_512"), ("teapot", "teapot_1080p"), ("sphere40", "sphere40_256"),
           ("sphere500", "sphere500_1080p")]


def test_fnv_known_answers():
    # FNV-1a-64 published test vectors: "" and "a"
    assert O.fnv1a64(np.zeros(0, np.uint8)) == 0xcbf29ce484222325
    assert O.fnv1a64(np.frombuffer(b"a", np.uint8)) == 0xaf63dc4c8601ec8c
    assert O.fnv1a64(np.frombuffer(b"foobar", np.uint8)) == 0x85944171f73967e8


def test_survey_appendix_b_hashes_are_the_committed_goldens():
    """The hashes written in SURVEY.md App. B (from the survey session) equal the regenerated fixtures."""
    g = T.golden()
    assert g["cornellbox_512"]["hits"] == 98889
    assert g["cornellbox_512"]["faceid_fnv"] == "aed66afebec55b2e"
    assert g["cornellbox_512"]["tuv_fnv"] == "18757eae0e11819f"
    assert g["teapot_1080p"]["hits"] == 569286
    assert g["teapot_1080p"]["faceid_fnv"] == "23ba070648e5b1dd"
    assert g["teapot_1080p"]["tuv_fnv"] == "dbcaedf52d31a371"
    assert g["sphere500_1080p"]["hits"] == 715539
    assert g["sphere500_1080p"]["faceid_fnv"] == "4bd8d6f04f9145e1"
    assert g["render_cornellbox_512_1thread"]["without_mtl"]["plane_off"]["fnv"] == "89f747731311c395"
    assert g["render_cornellbox_512_1thread"]["without_mtl"]["plane_on"]["fnv"] == "5569a14a30fb4f7b"


@pytest.mark.parametrize("mesh,entry", ENTRIES)
def test_builder_matches_reference_tree(mesh, entry):
    g = T.golden()[entry]
    om, ob = T.oracle_scene(mesh)
    nodes, idx = ob.arrays()
    assert len(nodes) == g["num_nodes"] and len(idx) == g["num_indices"]
    assert ob.stats() == g["stats"]
    assert T.fnv(idx) == g["indices_fnv"]
    assert T.fnv(T.mask_leaf_axis(nodes)) == g["nodes_fnv"]
    # structural invariants of the reference layout (bvh_accel.cc:420-427): pre-order, left child = parent + 1
    br = np.nonzero(nodes["flag"] == 0)[0]
    assert np.array_equal(nodes["data"][br, 0], br + 1)
    leaves = nodes[nodes["flag"] == 1]
    assert leaves["data"][:, 0].sum() == len(idx) and leaves["data"][:, 0].max() < 16
    assert sorted(idx.tolist()) == list(range(len(idx)))


def test_procedural_mesh_is_the_pinned_one():
    g = T.golden()["sphere500_1080p"]
    m = T.load_mesh("sphere500")
    assert m["faces"].shape == (1_000_000, 3) and m["vertices"].shape == (501_501, 3)
    assert T.fnv(m["vertices"]) == g["vertices_fnv"] and T.fnv(m["faces"]) == g["faces_fnv"]


@pytest.mark.parametrize("mesh,entry", ENTRIES)
def test_camera_rays_and_hits_match_reference(mesh, entry):
    g = T.golden()[entry]
    om, ob = T.oracle_scene(mesh)
    W, H = g["width"], g["height"]
    fr = O.camera_frame(g["eye"], g["lookat"], width=W, height=H)
    for a, b in zip(fr, T.golden_frame(g)):
        assert a.tobytes() == b.tobytes()
    rays = O.generate_grid(fr, W, H)
    assert T.fnv(rays) == g["rays_fnv"]
    o = ob.trace(rays, full=True, row=W)
    h, m = o["hits"], o["mask"]
    assert int(m.sum()) == g["hits"]
    assert T.fnv(h["faceID"]) == g["faceid_fnv"]
    assert T.fnv(np.stack([h["t"], h["u"], h["v"]], 1)[m]) == g["tuv_fnv"]
    for f, want in g["isect_fnv"].items():
        assert T.fnv(o["isects"][f][m]) == want, f
    for s in g["spots"]:
        r = h[s["y"] * W + s["x"]]
        assert int(r["faceID"]) == s["faceID"]
        assert float(r["t"]) == float.fromhex(s["t"])
        assert float(r["u"]) == float.fromhex(s["u"]) and float(r["v"]) == float.fromhex(s["v"])
    idx, ghits, gmask, gis = T.sample(entry)
    T.assert_hits_equal(h[idx], ghits, entry)
    assert np.array_equal(m[idx], gmask)
    for f in ("position", "geometricNormal", "normal", "texcoord", "f0", "f1", "f2"):
        assert np.ascontiguousarray(o["isects"][f][idx][gmask]).tobytes() == \
            np.ascontiguousarray(gis[f][gmask]).tobytes(), f
    # miss record (bvh_accel.cc:783-786)
    miss = h[~m]
    if len(miss):
        assert np.all(miss["t"] == np.finfo(np.float64).max) and np.all(miss["u"] == 0) and np.all(miss["v"] == 0)


def test_survey_spot_values():
    """Decimal spot values quoted in SURVEY.md App. B."""
    om, ob = T.oracle_scene("sphere500")
    fr = O.camera_frame((0, 0, 3), (0, 0, 0), width=1920, height=1080)
    rays = O.generate_rays(fr, [960, 700], [540, 300])
    h = ob.trace(rays)["hits"]
    assert h["faceID"].tolist() == [500500, 376645]
    assert h["t"][0] == 2.0 and h["u"][0] == h["v"][0] == 9.745494074354275e-15      # a vertex hit
    assert h["t"][1] == 2.210445456584494 and h["u"][1] == 0.37545304199839735 and h["v"][1] == 0.14227995413796055
    om, ob = T.oracle_scene("teapot")
    fr = O.camera_frame((5, 40, 150), (5, 40, 0), width=1920, height=1080)
    assert fr[1].tolist() == [-955.0, 580.0, -1153.6753060591777]
    h = ob.trace(O.generate_rays(fr, [960, 640], [540, 540]))["hits"]
    assert h["faceID"].tolist() == [543, 663]
    assert h["t"][0] == 106.58221157553848 and h["u"][0] == 0.16594723709385203 and h["v"][0] == 0.46262344120372662
    assert h["t"][1] == 115.72984520487498


def test_reference_traversal_counters():
    """nodes/ray and tris/ray of SURVEY §6 (the figures the roofline's algorithmic bytes are built from)."""
    om, ob = T.oracle_scene("sphere500")
    fr = O.camera_frame((0, 0, 3), (0, 0, 0), width=1920, height=1080)
    o = ob.trace(O.generate_grid(fr, 1920, 1080), row=1920)
    n = 1920 * 1080
    assert abs(o["n_node"] / n - 24.91) < 0.01 and abs(o["n_tri"] / n - 8.27) < 0.01
    assert o["max_stack"] == 16


@pytest.mark.parametrize("mtl,plane", [("with_mtl", False), ("with_mtl", True), ("without_mtl", False),
                                       ("without_mtl", True)])
def test_deterministic_render_matches_reference(mtl, plane):
    """Render() with the reference's sequential RNG stream (one OpenMP thread) -- bit-identical image.
    Zombie segments traced (as the reference does) and skipped (closed form) give the same image."""
    g = T.golden()["render_cornellbox_512_1thread"][mtl]["plane_on" if plane else "plane_off"]
    m = dict(T.load_mesh("cornellbox"))
    if mtl == "without_mtl":
        m["material_ids"] = np.full(len(m["faces"]), 0xFFFFFFFF, np.uint32)   # .mtl not found: materialID = -1
    om = T.oracle_mesh(m)
    ob = O.BVH.build(om)
    nodes, _ = ob.arrays()
    pl = O.plane_from_bbox(nodes[0]["bmin"], nodes[0]["bmax"]) if plane else None
    fr = O.camera_frame((0, 0, 20), (0, 0, 0), width=512, height=512)
    imgs = []
    for skip in ((0, 1) if not plane else (1,)):
        img, cnt, info = ob.render_pass(fr, 512, 512, plane=pl, rng_mode=0, skip_zombies=skip, shader=0, nthreads=1)
        imgs.append(img)
        assert T.fnv(img) == g["fnv"], (mtl, plane, skip)
        assert int((img.reshape(-1, 3).sum(1) != 0).sum()) == g["nonzero"]
        assert abs(float(img.astype(np.float64).sum()) - g["sum"]) < 1e-2
        assert (cnt == 1).all()
    if not plane and mtl == "without_mtl":
        # SURVEY App. A.5 (measured there with 8 RNG streams): ~1.745 M Trace calls, ~55 % of them zombies
        _, _, info = ob.render_pass(fr, 512, 512, rng_mode=0, skip_zombies=1, shader=0, nthreads=1)
        total = info["trace_calls"] + info["zombies"]
        assert abs(total - 1_745_149) < 2000 and abs(info["zombies"] / total - 0.547) < 0.005


def test_rng_stream_is_xorshift128():
    """randomreal (render.cc:137-168): Marsaglia xorshift128 with seeds 123456789+tid, 362436069, 521288629, 88675123."""
    def xs(x, y, z, w, n):
        out = []
        for _ in range(n):
            t = (x ^ (x << 11)) & 0xFFFFFFFF
            x, y, z = y, z, w
            w = (w ^ (w >> 19)) ^ (t ^ (t >> 8))
            out.append(w / 4294967296.0)
        return np.array(out)
    for tid in (0, 3):
        assert np.array_equal(O.rng_stream(0, 0, 20, reference_tid=tid), xs(123456789 + tid, 362436069, 521288629, 88675123, 20))
    a, b = O.rng_stream(5, 0, 8), O.rng_stream(5, 1, 8)
    assert not np.array_equal(a, b) and np.array_equal(a, O.rng_stream(5, 0, 8))
    assert ((a >= 0) & (a < 1)).all()


def test_occlusion_oracle_definition():
    """occluded == closest-hit Traverse returns t < tmax."""
    om, ob = T.oracle_scene("sphere40")
    nodes, _ = ob.arrays()
    rng = np.random.default_rng(1)
    rays = T.random_rays(rng, 20000, nodes[0]["bmin"], nodes[0]["bmax"])
    o = ob.trace(rays)
    t = np.where(o["mask"], o["hits"]["t"], 3.0)
    tmax = t * rng.choice([0.5, 1.0, 2.0], len(t))
    want = o["mask"] & (o["hits"]["t"] < tmax)
    assert np.array_equal(ob.occluded(rays, tmax), want)
    assert want.any() and (~want).any()


def test_dump_load_roundtrip_and_format(tmp_path):
    """BVHAccel::Dump / Load byte format (bvh_accel.cc:484-544): u64 nnodes, nodes, u64 nindices, indices."""
    om, ob = T.oracle_scene("sphere40")
    p = str(tmp_path / "bvh.bin")
    assert ob.dump(p)
    raw = open(p, "rb").read()
    nodes, idx = ob.arrays()
    assert len(raw) == 8 + 64 * len(nodes) + 8 + 4 * len(idx)
    assert int.from_bytes(raw[:8], "little") == len(nodes)
    assert raw[8:8 + 64 * len(nodes)] == nodes.tobytes()
    assert int.from_bytes(raw[8 + 64 * len(nodes):16 + 64 * len(nodes)], "little") == len(idx)
    ob2 = O.BVH.load(p, om)
    n2, i2 = ob2.arrays()
    assert n2.tobytes() == nodes.tobytes() and i2.tobytes() == idx.tobytes()
