"""The oracle restatement against the UNMODIFIED reference compiled into oracle/_ref/libmallie_ref.so
(oracle/Makefile + oracle/ref_harness.cc), on fresh seeded inputs -- beyond the committed golden vectors.
CPU only; skipped where the reference library has not been built (it needs /root/reference at build time,
the prebuilt .so travels with the repo snapshot).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import orabind as O
from oracle import refbind as R
from tests import common as T

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libmallie_ref.so not built")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def random_soup(rng, n, spread=1.0, size=0.15):
    """Triangle soup with float-exact coordinates, some degenerate and some duplicated triangles."""
    c = rng.uniform(-spread, spread, (n, 1, 3))
    v = (c + rng.normal(0, size, (n, 3, 3))).astype(np.float32).astype(np.float64)
    if n > 20:
        v[5] = v[4]                      # exact duplicate: equal-t tie, last visited wins
        v[7, 1] = v[7, 0]                # degenerate (zero area): det == 0
    return v.reshape(-1, 3), np.arange(3 * n, dtype=np.uint32).reshape(n, 3)


@pytest.mark.parametrize("seed,n", [(0, 1), (1, 15), (2, 16), (3, 17), (4, 500), (5, 20000)])
def test_builder_bit_identical(seed, n):
    rng = np.random.default_rng(seed)
    v, f = random_soup(rng, n)
    rs = R.RefScene.from_arrays(v, f)
    rs.build()
    rn, ri = rs.bvh()
    ob = O.BVH.build(O.Mesh(v, f))
    on, oi = ob.arrays()
    assert oi.tobytes() == ri.tobytes()
    assert T.mask_leaf_axis(on).tobytes() == T.mask_leaf_axis(rn).tobytes()
    assert ob.stats() == rs.stats()


def test_builder_on_shared_vertex_mesh_and_grid_degeneracies():
    # all centroids on one plane / identical centroids: the object-median fallback (bvh_accel.cc:405-409)
    g = np.arange(12, dtype=np.float64)
    xx, yy = np.meshgrid(g, g, indexing="ij")
    v = np.stack([xx.ravel(), yy.ravel(), np.zeros(xx.size)], 1)
    idx = (np.arange(11)[:, None] * 12 + np.arange(11)[None, :]).ravel()
    f = np.concatenate([np.stack([idx, idx + 12, idx + 1], 1), np.stack([idx + 1, idx + 12, idx + 13], 1)]).astype(np.uint32)
    f = np.concatenate([f, f[:40]])      # 40 exact duplicates -> identical centroids
    rs = R.RefScene.from_arrays(v, f)
    rs.build()
    rn, ri = rs.bvh()
    on, oi = O.BVH.build(O.Mesh(v, f)).arrays()
    assert oi.tobytes() == ri.tobytes() and T.mask_leaf_axis(on).tobytes() == T.mask_leaf_axis(rn).tobytes()


@pytest.mark.parametrize("seed,n", [(10, 40), (11, 3000)])
def test_traverse_bit_identical_on_random_rays(seed, n):
    rng = np.random.default_rng(seed)
    v, f = random_soup(rng, n)
    mats = rng.integers(0, 5, n).astype(np.uint32)
    nrm = rng.normal(size=(n, 3, 3))
    uvs = rng.uniform(size=(n, 3, 2))
    rs = R.RefScene.from_arrays(v, f, mats, nrm, uvs)
    rs.build()
    om = O.Mesh(v, f, mats, nrm, uvs)
    ob = O.BVH.build(om)
    rays = T.random_rays(rng, 50000, v.min(0), v.max(0))
    # plus axis-parallel rays (inf / NaN in the slab test) through vertices
    ax = np.eye(3)[rng.integers(0, 3, 2000)] * rng.choice([-1.0, 1.0], (2000, 1))
    pts = v[rng.integers(0, len(v), 2000)]
    rays = np.concatenate([rays, np.concatenate([pts - 5.0 * ax, ax], 1)])
    with np.errstate(all="ignore"):
        r = rs.trace(rays, full=True)
        o = ob.trace(rays, full=True)
    T.assert_hits_equal(o["hits"], r["hits"], "oracle vs reference")
    m = r["mask"]
    assert np.array_equal(m, o["mask"]) and m.sum() > 1000
    for fld in ("position", "geometricNormal", "normal", "texcoord", "f0", "f1", "f2", "materialID"):
        assert np.ascontiguousarray(o["isects"][fld][m]).tobytes() == np.ascontiguousarray(r["isects"][fld][m]).tobytes(), fld


def test_traverse_bit_identical_on_edge_case_rays():
    """The very ray set the GPU parity test uses (tests/common.py::edge_case_rays: inf / NaN slabs, -0.0 components,
    origins on box planes, rays in a triangle's plane, zero and huge values): oracle == reference here, GPU == oracle
    in tests/test_gpu_parity.py::test_edge_case_rays."""
    m = T.load_mesh("cornellbox")
    rs = R.RefScene.from_arrays(m["vertices"], m["faces"], m["material_ids"], m["normals"], m["uvs"])
    rs.build()
    _, ob = T.oracle_scene("cornellbox")
    rays = T.edge_case_rays("cornellbox")
    with np.errstate(all="ignore"):
        r = rs.trace(rays, full=True)
        o = ob.trace(rays, full=True)
    T.assert_hits_equal(o["hits"], r["hits"], "oracle vs reference, edge cases")
    assert np.array_equal(r["mask"], o["mask"]) and r["mask"].sum() > 100
    rs.close()


def test_camera_frame_bit_identical_incl_quaternions():
    rng = np.random.default_rng(21)
    cases = [((0, 0, 20), (0, 0, 0), (0, 1, 0), 45.0, (0, 0, 0, 0), 512, 512),
             ((5, 40, 150), (5, 40, 0), (0, 1, 0), 45.0, (0, 0, 0, 0), 1920, 1080),
             ((0, 0, 3), (0, 0, 0), (0, 1, 0), 45.0, (0, 0, 0, 1), 3840, 2160)]
    for _ in range(40):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        cases.append((tuple(rng.uniform(-30, 30, 3)), tuple(rng.uniform(-3, 3, 3)), tuple(rng.normal(size=3)),
                      float(rng.uniform(10, 120)), tuple(q), int(rng.integers(16, 2000)), int(rng.integers(16, 2000))))
    for eye, lookat, up, fov, quat, W, H in cases:
        fr = R.camera_frame(eye, lookat, up, fov, quat, W, H)
        fo = O.camera_frame(eye, lookat, up, fov, quat, W, H)
        for a, b in zip(fo, fr):
            assert a.tobytes() == b.tobytes(), (eye, lookat, up, fov, quat, W, H)
        px, py = rng.uniform(0, W, 64), rng.uniform(0, H, 64)
        assert O.generate_rays(fo, px, py).tobytes() == R.camera_generate(eye, lookat, up, fov, quat, W, H, px, py).tobytes()


def test_plane_intersect_bit_identical():
    rng = np.random.default_rng(5)
    abcd = O.plane_from_bbox(np.array([-1.0, -0.731, -1.0]), np.array([1.0, 1.2, 1.0]))
    rays = np.concatenate([rng.uniform(-2, 2, (5000, 3)), rng.normal(size=(5000, 3))], 1)
    rays[:50, 4] = 0.0                                  # parallel to the plane
    t_in = rng.choice([1e30, 0.5, 2.0, np.finfo(np.float64).max], 5000)
    to, po, no, ho = O.plane_intersect(abcd, rays, t_in)
    tr, pr, nr, hr = R.plane_intersect(abcd, rays, t_in)
    assert np.array_equal(ho, hr) and ho.any() and (~ho).any()
    assert to.tobytes() == tr.tobytes()
    assert po[ho].tobytes() == pr[hr].tobytes() and no[ho].tobytes() == nr[hr].tobytes()


def test_dump_is_byte_compatible_with_reference(tmp_path):
    m = T.load_mesh("sphere40")
    rs = R.RefScene.from_arrays(m["vertices"], m["faces"])
    rs.build()
    pr, po = str(tmp_path / "ref.bvh"), str(tmp_path / "ora.bvh")
    assert rs.dump(pr)
    om, ob = T.oracle_scene("sphere40")
    assert ob.dump(po)
    a, b = open(pr, "rb").read(), open(po, "rb").read()
    assert len(a) == len(b)
    nn = int.from_bytes(a[:8], "little")
    na = T.mask_leaf_axis(np.frombuffer(a[8:8 + 64 * nn], O.NODE_DTYPE))
    nb = T.mask_leaf_axis(np.frombuffer(b[8:8 + 64 * nn], O.NODE_DTYPE))
    assert na.tobytes() == nb.tobytes() and a[8 + 64 * nn:] == b[8 + 64 * nn:]
    # the reference loads the oracle's file and traces identically
    rs2 = R.RefScene.from_arrays(m["vertices"], m["faces"])
    assert rs2.load(po)
    fr = O.camera_frame((0.3, 0.2, 3), (0, 0, 0), width=128, height=128)
    rays = O.generate_grid(fr, 128, 128)
    T.assert_hits_equal(rs2.trace(rays)["hits"], ob.trace(rays)["hits"], "load")


@pytest.mark.parametrize("plane", [False, True])
@pytest.mark.parametrize("mesh,eye,lookat", [("sphere40", (0.3, 0.2, 3), (0, 0, 0)), ("cornellbox", (0, 0, 20), (0, 0, 0)),
                                             ("teapot", (5, 40, 150), (5, 40, 0))])
def test_render_one_thread_bit_identical(mesh, eye, lookat, plane):
    """Render() keeps function-static state (render.cc:113-116,615): run the reference in a fresh process.  The cornell
    box and the teapot carry material ids, facevarying normals and uvs (BuildIntersection's interpolation feeds the
    bounce directions)."""
    code = f"""
import sys; sys.path.insert(0, {ROOT!r})
import numpy as np
from oracle import refbind as R, orabind as O
from tests import common as T
m = T.load_mesh({mesh!r})
rs = R.RefScene.from_arrays(m["vertices"], m["faces"], m.get("material_ids"), m.get("normals"), m.get("uvs")); rs.build()
img, cnt, sec = rs.render(160, 120, {eye!r}, {lookat!r}, plane={plane}, nthreads=1)
print("RESULT %016x %d" % (O.fnv1a64(img), int(cnt.sum())))
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout
    line = [l for l in out.splitlines() if l.startswith("RESULT")][0].split()
    om, ob = T.oracle_scene(mesh)
    nodes, _ = ob.arrays()
    pl = O.plane_from_bbox(nodes[0]["bmin"], nodes[0]["bmax"]) if plane else None
    fr = O.camera_frame(eye, lookat, width=160, height=120)
    img, cnt, _ = ob.render_pass(fr, 160, 120, plane=pl, rng_mode=0, skip_zombies=1, shader=0, nthreads=1)
    assert T.fnv(img) == line[1] and int(cnt.sum()) == int(line[2]) == 160 * 120
    assert img.max() > 0


def test_env_camera_rays_bit_identical():
    """Camera::GenerateEnvRay / GenerateStereoEnvRay (camera.cc:242-329): same libm, same bits."""
    rng = np.random.default_rng(3)
    W, H = 200, 100
    px = np.concatenate([rng.uniform(-0.5, W, 500), [0.0, W - 1.0, 17.0]])
    py = np.concatenate([rng.uniform(-0.5, H, 500), [0.0, H - 1.0, H / 2.0]])
    eye = (0.3, -0.2, 2.5)
    fr = O.camera_frame(eye, (0, 0, 0), width=W, height=H)
    for stereo in (False, True):
        want = R.camera_generate_env(eye, (0, 0, 0), (0, 1, 0), 45.0, (0, 0, 0, 0), W, H, px, py, stereo=stereo)
        got = O.generate_env(fr[0], W, H, px, py, stereo=stereo)
        assert got.tobytes() == want.tobytes(), stereo


@pytest.mark.parametrize("stereo", [False, True])
def test_render_panoramic_one_thread_bit_identical(stereo):
    """RenderPanoramic (render.cc:710-763) in a fresh process (function-static RNG initialisation)."""
    code = f"""
import sys; sys.path.insert(0, {ROOT!r})
import numpy as np
from oracle import refbind as R, orabind as O
from tests import common as T
m = T.load_mesh("sphere40")
rs = R.RefScene.from_arrays(m["vertices"], m["faces"]); rs.build()
img, cnt, sec = rs.render_panoramic(96, 48, (0.1, 0.2, 0.3), (0, 0, 1), stereo={stereo}, nthreads=1)
print("RESULT %016x %d %r" % (O.fnv1a64(img), int(cnt.sum()), float(img.sum(dtype=np.float64))))
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout
    line = [l for l in out.splitlines() if l.startswith("RESULT")][0].split()
    om, ob = T.oracle_scene("sphere40")
    fr = O.camera_frame((0.1, 0.2, 0.3), (0, 0, 1), width=96, height=48)
    img, cnt = O.render_panoramic(ob, fr[0], 96, 48, stereo=stereo, rng_mode=0, nthreads=1)
    assert int(cnt.sum()) == int(line[2]) == 96 * 48 * 10
    assert T.fnv(img) == line[1], (float(img.sum(dtype=np.float64)), line[3])
    assert img.max() > 0
