"""CPU tests of the host-side ingestion a Mallie program runs before the hot path: the OBJ / ESON loaders
(MeshLoader::LoadObj / LoadESON, importers/mesh_loader.cc:26-310 over tiny_obj_loader.cc) and the config.json
reader (LoadJSONConfig, main.cc:98-205), through the C ABI.  Expected meshes were produced by the UNMODIFIED
reference loaders (tests/golden/make_loader_golden.py)."""
import json
import os

import numpy as np
import pytest

import mallie_b200 as M
from oracle import orabind as O
from tests import common as T

REF = "/root/reference"


def in_dir(path):
    class _cd:
        def __enter__(self):
            self.old = os.getcwd()
            os.chdir(path)

        def __exit__(self, *a):
            os.chdir(self.old)
    return _cd()


def assert_mesh_equal(got, want):
    for k in ("vertices", "faces", "material_ids", "normals", "uvs"):
        if k in want and want[k] is not None:
            assert got[k] is not None, k
            a, b = np.asarray(want[k]).reshape(-1), got[k].reshape(-1)
            assert a.shape == b.shape, (k, a.shape, b.shape)
            assert np.array_equal(a, b, equal_nan=True), f"{k} differs from the reference loader"
        else:
            assert got[k] is None, f"{k}: the reference loader leaves it NULL"


def test_obj_loader_tricky_fixture():
    """quads / pentagon fans, negative indices, v/vt/vn forms, usemtl / g / o flushing quirks (faces appended by
    a usemtl that is directly followed by g / o are dropped), CRLF, odd numbers (1e, 1e39 -> inf, +1.0, -.5)."""
    with in_dir(T.GOLDEN):     # mtllib is resolved against the current directory, as in the reference
        got = M.load_mesh("tricky.obj")
    z = np.load(os.path.join(T.GOLDEN, "tricky_mesh.npz"))
    assert_mesh_equal(got, {k: z[k] for k in z.files})
    assert len(got["faces"]) == 15 and np.isinf(got["vertices"]).any()
    assert set(got["material_ids"].tolist()) == {0xFFFFFFFF, 0, 1, 2}


def test_obj_loader_without_mtl_gives_minus_one():
    got = M.load_mesh(os.path.join(T.GOLDEN, "tricky.obj"))      # cwd != fixture dir: tricky.mtl is not found
    assert (got["material_ids"] == 0xFFFFFFFF).all()


def test_obj_loader_missing_file_and_empty(tmp_path):
    with pytest.raises(M.MallieB200Error):
        M.load_mesh(str(tmp_path / "nope.obj"))
    p = tmp_path / "empty.obj"
    p.write_text("# nothing\n\nv 1 2 3\n")
    got = M.load_mesh(str(p))
    assert got["vertices"].shape == (0, 3) and got["faces"].shape == (0, 3)


def test_eson_loader_small_fixture():
    got = M.load_mesh(os.path.join(T.GOLDEN, "small.eson"))
    z = np.load(os.path.join(T.GOLDEN, "small_eson_mesh.npz"))
    assert_mesh_equal(got, {k: z[k] for k in z.files})      # normals / uvs stay NULL even though the file has uvs
    with pytest.raises(M.MallieB200Error):
        M.load_mesh(os.path.join(T.GOLDEN, "tricky.obj"), kind="eson")


def test_scene_scale_and_fit():
    path = os.path.join(T.GOLDEN, "small.eson")
    base = M.load_mesh(path)["vertices"]
    assert np.array_equal(M.load_mesh(path, scene_scale=2.5)["vertices"], base * 2.5)
    fit = M.load_mesh(path, scene_fit=True)["vertices"]
    bmin, bmax = base.min(0), base.max(0)
    want = (((base - bmin) * (1.0 / (bmax - bmin))) - 0.5) * 2.0        # scene.cc:141-158, same operation order
    assert np.array_equal(fit, want) and abs(fit).max() <= 1.0


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference assets are only mounted in the authoring container")
def test_shipped_assets_match_reference_loader_hashes():
    pins = json.load(open(os.path.join(T.GOLDEN, "loader_golden.json")))
    with in_dir(REF):
        cases = {"cornellbox_obj": ("cornellbox_suzanne.obj", 1.0), "teapot_obj": ("teapot.obj", 1.0),
                 "cornellbox_eson": ("cornellbox_suzanne.eson", 1.0), "cornellbox_obj_x2.5": ("cornellbox_suzanne.obj", 2.5)}
        for name, (fn, scale) in cases.items():
            got = M.load_mesh(fn, scene_scale=scale)
            assert [len(got["vertices"]), len(got["faces"])] == pins[name]["shape"]
            for k in ("vertices", "faces", "material_ids", "normals", "uvs"):
                if k in pins[name]:
                    assert "%016x" % O.fnv1a64(np.ascontiguousarray(got[k])) == pins[name][k], (name, k)
                else:
                    assert got[k] is None


def test_committed_golden_meshes_are_what_the_loader_makes_of_obj_text(tmp_path):
    """Round trip without the reference assets: write the committed golden cornell-box mesh as OBJ text
    (float-exact %.9g), load it, and get the same vertices / faces back (per-shape dedupe collapses nothing here
    because every face group is written with its own vertices)."""
    m = T.load_mesh("cornellbox")
    v, f = m["vertices"], m["faces"]
    p = tmp_path / "rt.obj"
    with open(p, "w") as fp:
        for x in v:
            fp.write("v %.9g %.9g %.9g\n" % tuple(x))
        for a in f:
            fp.write("f %d %d %d\n" % tuple(int(i) + 1 for i in a))
    got = M.load_mesh(str(p))
    # one face group: vertices are renumbered in order of first use
    order, seen = [], {}
    for i in f.reshape(-1):
        if int(i) not in seen:
            seen[int(i)] = len(order)
            order.append(int(i))
    assert np.array_equal(got["vertices"], v[order])
    assert np.array_equal(got["faces"], np.vectorize(seen.get)(f).astype(np.uint32))


# ------------------------------------------------------------------------------------------------ config.json
def test_config_defaults_and_shipped_keys():
    c = M.load_config(text="{}")
    assert (c.fov, c.width, c.height, c.num_passes, c.scene_scale) == (45.0, 512, 512, 10, 1.0)   # render.h:33-48
    assert list(c.eye) == [0, 0, -5] and list(c.up) == [0, 1, 0] and not c.plane and c.max_path_length == 16
    shipped = """{
        "0obj_filename" : "cornellbox_suzanne.obj", "eson_filename" : "cornellbox_suzanne.eson",
        "0magicavoxel_filename" : "castle.vox", "material_filename" : "teapot.material.json",
        "resolution" : [512, 512], "scene_scale" : 1.0, "num_passes" : 1000, "plane" : true,
        "eye" : [0, 0, 20], "lookat" : [0, 0, 0], "up" : [0, 1, 0], "dummy" : 0 }"""
    c = M.load_config(text=shipped)
    assert c.obj_filename == b"" and c.eson_filename == b"cornellbox_suzanne.eson"      # "0..." keys are unknown keys
    assert c.material_filename == b"teapot.material.json" and c.num_passes == 1000 and c.plane == 1
    assert list(c.eye) == [0, 0, 20] and (c.width, c.height) == (512, 512)


def test_config_type_rules_and_quirks(tmp_path):
    c = M.load_config(text='{"num_passes": 7, "num_photons": 99, "resolution": [640.9, 480.2], "fov": 60,'
                           ' "eye": [1, 2], "up": [0, "x", 1], "scene_scale": "2", "plane": 1, "scene_fit": true}')
    assert c.num_passes == 99                      # num_photons overwrites num_passes (main.cc:192-195)
    assert (c.width, c.height) == (640, 480)       # truncated
    assert c.fov == 45.0                           # not a key of the reference
    assert list(c.eye) == [0, 0, -5]               # wrong length: ignored
    assert list(c.up) == [0, 0, 1]                 # non-number element reads as 0
    assert c.scene_scale == 1.0 and c.plane == 0 and c.scene_fit == 1     # wrong JSON type: ignored
    for bad in ("[1,2]", "{", '{"a":1,"a":2}', "", '{"a": tru}'):
        with pytest.raises(M.MallieB200Error):
            M.load_config(text=bad)
    with pytest.raises(M.MallieB200Error):
        M.load_config(path=str(tmp_path / "missing.json"))
    p = tmp_path / "c.json"
    p.write_text('{"obj_filename": "~/x.obj", "shader": "primary_shadow", "light": [1,2,3], "max_path_length": 5, "gpus": 8}')
    c = M.load_config(path=str(p))
    assert c.obj_filename.decode() == os.path.expanduser("~/x.obj")
    assert c.shader == M.SHADER_PRIMARY_SHADOW and list(c.light) == [1, 2, 3] and c.max_path_length == 5 and c.num_gpus == 8


def test_scene_init_matches_the_reference_scene_init(tmp_path):
    """The reference's own Scene::Init (scene.cc:66-250, through oracle/_ref: load, scene_fit / scene_scale, default
    build, BoundingBox) against load_mesh + the host builder on the committed fixtures and on a flat mesh, whose zero
    extent takes the `invExtent = extent` branch of scene_fit (scene.cc:134-137)."""
    from oracle import refbind as R
    if not R.available():
        pytest.skip("oracle/_ref is built where /root/reference is mounted")
    flat = tmp_path / "flat.obj"
    with open(flat, "w") as fp:
        for i in range(6):
            for j in range(6):
                fp.write("v %d %g 7.25\n" % (i, 0.5 * j))
        for i in range(5):
            for j in range(5):
                a = i * 6 + j + 1
                fp.write("f %d %d %d\nf %d %d %d\n" % (a, a + 1, a + 6, a + 1, a + 7, a + 6))
    with in_dir(T.GOLDEN):
        for path in ("tricky.obj", "small.eson", str(flat)):
            for scale, fit in ((1.0, False), (2.5, False), (1.0, True), (0.3, True)):
                rs = R.RefScene.init(path, scale, fit)
                want = rs.mesh()
                got = M.load_mesh(path, scene_scale=scale, scene_fit=fit)
                assert got["vertices"].tobytes() == want["vertices"].tobytes(), (path, scale, fit)
                assert np.array_equal(got["faces"], want["faces"])
                hb = M.HostBVH.build(got["vertices"], got["faces"])
                nodes, idx = hb.arrays()
                assert T.tree_fingerprint(nodes, idx) == T.tree_fingerprint(*rs.bvh())
                assert np.array_equal(nodes[0]["bmin"], rs.bounds[0]) and np.array_equal(nodes[0]["bmax"], rs.bounds[1])
                hb.close()
                rs.close()
