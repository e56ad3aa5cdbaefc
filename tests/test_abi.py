"""The C-ABI library without a GPU: it loads, exports every symbol include/mallie_b200.h declares, its
host-side logic (BVH build / dump / load, camera frame, plane, band arithmetic, argument checking) matches
the oracle bit for bit, and every compute entry fails LOUDLY (no CPU fallback) when there is no device.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import mallie_b200 as M
from mallie_b200 import capi, tiles
from oracle import orabind as O
from tests import common as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mallie_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    declared = header_symbols()
    assert len(declared) >= 30
    assert sorted(capi.EXPORTS) == declared, set(capi.EXPORTS) ^ set(declared)
    lib = capi.lib()
    for s in declared:
        assert hasattr(lib, s), s
    # and the dynamic symbol table agrees (nothing resolved lazily from somewhere else)
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (mb200_[a-z0-9_]+)", out))
    assert set(declared) <= exported


def test_library_does_not_link_the_oracle():
    out = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "mallie_ref" not in out
    nm = subprocess.run(["nm", "-D", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "ora_" not in nm and "ref_scene" not in nm


def test_struct_layouts_match_the_header():
    assert C.sizeof(capi.CameraFrame) == 96
    assert C.sizeof(capi.Counters) == 32 and C.sizeof(capi.RenderStats) == 64
    assert C.sizeof(capi.BuildOptions) == 24 and C.sizeof(capi.BuildStats) == 12
    # mb200_render_params: 6 ints, frame (96, 8-aligned), int, float[4], int, u32, int, int, double[3], 5 ints
    assert C.sizeof(capi.RenderParams) == 208
    assert capi.RAY_DTYPE.itemsize == 48 and capi.HIT_DTYPE.itemsize == 32
    assert capi.ISECT_DTYPE.itemsize == 184 and capi.NODE_DTYPE.itemsize == 64
    assert capi.ISECT_DTYPE.fields["position"][1] == 48 and capi.ISECT_DTYPE.fields["normal"][1] == 96
    assert capi.ISECT_DTYPE.fields["texcoord"][1] == 168
    assert capi.lib().mb200_version().decode().startswith("mallie_b200")


def test_header_is_plain_c_and_ctypes_mirrors_agree(tmp_path):
    """include/mallie_b200.h compiles as C (gcc -std=c99) and its struct sizes are the ctypes mirrors' sizes."""
    import subprocess
    names = {"mb200_ray": capi.RAY_DTYPE.itemsize, "mb200_hit": capi.HIT_DTYPE.itemsize,
             "mb200_isect": capi.ISECT_DTYPE.itemsize, "mb200_bvh_node": capi.NODE_DTYPE.itemsize,
             "mb200_build_options": C.sizeof(capi.BuildOptions), "mb200_build_stats": C.sizeof(capi.BuildStats),
             "mb200_counters": C.sizeof(capi.Counters), "mb200_camera_frame": C.sizeof(capi.CameraFrame),
             "mb200_render_params": C.sizeof(capi.RenderParams), "mb200_render_stats": C.sizeof(capi.RenderStats),
             "mb200_config": C.sizeof(capi.Config)}
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "mallie_b200.h"\nint main(void){\n' +
                   "".join('printf("%s %%zu\\n", sizeof(%s));\n' % (n, n) for n in names) +
                   'printf("pixel_step %zu\\n", offsetof(mb200_render_params, pixel_step));return 0;}\n')
    exe = tmp_path / "sizes"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, "-o", str(exe), str(src)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for n, want in names.items():
        assert int(got[n]) == want, (n, got[n], want)
    assert int(got["pixel_step"]) == capi.RenderParams.pixel_step.offset


@pytest.mark.parametrize("mesh,entry", [("cornellbox", "cornellbox_512"), ("teapot", "teapot_1080p"),
                                        ("sphere40", "sphere40_256"), ("sphere500", "sphere500_1080p")])
def test_host_builder_is_the_reference_builder(mesh, entry):
    g = T.golden()[entry]
    m = T.load_mesh(mesh)
    hb = M.HostBVH.build(m["vertices"], m["faces"])
    nodes, idx = hb.arrays()
    assert hb.stats() == g["stats"]
    assert T.fnv(idx) == g["indices_fnv"] and T.fnv(T.mask_leaf_axis(nodes)) == g["nodes_fnv"]
    hb.close()


def test_host_builder_options_and_degenerate_inputs():
    rng = np.random.default_rng(4)
    v = rng.uniform(-1, 1, (900, 3)).astype(np.float32).astype(np.float64)
    f = np.arange(900, dtype=np.uint32).reshape(300, 3)
    for kw, okw in ((dict(min_leaf=4), dict(min_leaf=4)), (dict(max_depth=3), dict(max_depth=3)),
                    (dict(bin_size=8, cost_taabb=1.0), dict(bin_size=8, cost_taabb=1.0))):
        hb = M.HostBVH.build(v, f, **kw)
        on, oi = O.BVH.build(O.Mesh(v, f), **okw).arrays()
        n, i = hb.arrays()
        assert i.tobytes() == oi.tobytes() and T.mask_leaf_axis(n).tobytes() == T.mask_leaf_axis(on).tobytes(), kw
    # empty mesh: no nodes, no indices (Build on zero faces)
    hb = M.HostBVH.build(np.zeros((0, 3)), np.zeros((0, 3), np.uint32))
    n, i = hb.arrays()
    assert len(i) == 0 and len(n) <= 1
    # single triangle: one leaf root
    hb = M.HostBVH.build(v[:3], f[:1])
    n, i = hb.arrays()
    assert len(n) == 1 and n[0]["flag"] == 1 and tuple(n[0]["data"]) == (1, 0)
    # face index out of range is an error, not a crash
    bad = f.copy()
    bad[10, 1] = 5000
    with pytest.raises(M.MallieB200Error):
        M.HostBVH.build(v, bad)


def test_dump_load_byte_compatible(tmp_path):
    m = T.load_mesh("sphere40")
    hb = M.HostBVH.build(m["vertices"], m["faces"])
    p1, p2 = str(tmp_path / "a.bvh"), str(tmp_path / "b.bvh")
    hb.dump(p1)
    om, ob = T.oracle_scene("sphere40")
    ob.dump(p2)
    a, b = open(p1, "rb").read(), open(p2, "rb").read()
    nn = int.from_bytes(a[:8], "little")
    assert len(a) == len(b) == 8 + 64 * nn + 8 + 4 * len(m["faces"])
    assert T.mask_leaf_axis(np.frombuffer(a[8:8 + 64 * nn], capi.NODE_DTYPE)).tobytes() == \
        T.mask_leaf_axis(np.frombuffer(b[8:8 + 64 * nn], capi.NODE_DTYPE)).tobytes()
    assert a[8 + 64 * nn:] == b[8 + 64 * nn:]
    hb2 = M.HostBVH.load(p2)                      # a file written by the oracle (== the reference's format)
    n1, i1 = hb.arrays()
    n2, i2 = hb2.arrays()
    assert i1.tobytes() == i2.tobytes() and T.mask_leaf_axis(n1).tobytes() == T.mask_leaf_axis(n2).tobytes()
    with pytest.raises(M.MallieB200Error):
        M.HostBVH.load(str(tmp_path / "missing.bvh"))
    open(str(tmp_path / "trunc.bvh"), "wb").write(a[:100])
    with pytest.raises(M.MallieB200Error):
        M.HostBVH.load(str(tmp_path / "trunc.bvh"))


def test_camera_frame_and_plane_match_oracle():
    rng = np.random.default_rng(9)
    cases = [((0, 0, 20), (0, 0, 0), (0, 1, 0), 45.0, (0, 0, 0, 0), 512, 512)]
    for _ in range(50):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        cases.append((tuple(rng.uniform(-30, 30, 3)), tuple(rng.uniform(-3, 3, 3)), tuple(rng.normal(size=3)),
                      float(rng.uniform(10, 120)), tuple(q), int(rng.integers(16, 4000)), int(rng.integers(16, 4000))))
    for eye, lookat, up, fov, quat, W, H in cases:
        fg = M.camera_frame(eye, lookat, up, fov, quat, W, H)
        fo = O.camera_frame(eye, lookat, up, fov, quat, W, H)
        for a, b in zip(fg.arrays(), fo):
            assert a.tobytes() == b.tobytes()
    for g in ("cornellbox_512", "teapot_1080p", "sphere40_256", "sphere500_1080p"):
        e = T.golden()[g]
        fg = M.camera_frame(e["eye"], e["lookat"], width=e["width"], height=e["height"])
        for a, b in zip(fg.arrays(), T.golden_frame(e)):
            assert a.tobytes() == b.tobytes()
    for _ in range(20):
        lo = rng.uniform(-5, 0, 3)
        hi = lo + rng.uniform(0.1, 9, 3)
        assert M.plane_from_bounds(lo, hi).tobytes() == O.plane_from_bbox(lo, hi).tobytes()


def test_band_arithmetic_matches_tile_scheduler():
    p = capi.RenderParams()
    for H in (1, 7, 118, 1080, 2160):
        for rows in (4, 8, 16):
            for G in (1, 2, 3, 8):
                total = 0
                for r in range(G):
                    capi.lib().mb200_render_params_default(C.byref(p), 64, H)
                    p.band_rows, p.band_count, p.band_index = rows, G, r
                    n = int(capi.lib().mb200_band_local_rows(C.byref(p)))
                    assert n == len(tiles.band_rows_of_rank(H, rows, G, r))
                    total += n
                assert total == H
    capi.lib().mb200_render_params_default(C.byref(p), 640, 480)
    assert (p.width, p.height, p.x0, p.y0, p.x1, p.y1) == (640, 480, 0, 0, 640, 480)
    assert p.max_path_length == 16 and p.jitter == 1 and p.shader == M.SHADER_PATHTRACE and p.band_rows == 0


def test_gather_permutation_is_a_bijection():
    for H, rows, G in ((1080, 8, 8), (118, 4, 3), (2160, 16, 8), (5, 8, 2)):
        perm = tiles.gather_permutation(H, rows, G)
        pad = tiles.max_local_rows(H, rows, G)
        assert len(set(perm.tolist())) == H and perm.max() < G * pad
        for r in range(G):
            ys = tiles.band_rows_of_rank(H, rows, G, r)
            assert np.array_equal(perm[ys], r * pad + np.arange(len(ys)))


def test_argument_errors_are_reported_not_fatal():
    L = capi.lib()
    h = C.c_void_p()
    assert L.mb200_bvh_build(None, None, 0, None, 0, None) == -1
    assert b"null" in L.mb200_last_error()
    assert L.mb200_bvh_load(C.byref(h), b"/nonexistent/file") == -5
    assert L.mb200_trace_closest(None, None, 0, None, None) == -1
    assert L.mb200_render_pass(None, None, None, None, None) == -1
    assert L.mb200_scene_bounds(None, None, None) == -1
    assert L.mb200_band_local_rows(None) == 0
    assert L.mb200_scene_device_bytes(None) == 0 and L.mb200_scene_device(None) == -1
    L.mb200_scene_destroy(None)
    L.mb200_bvh_destroy(None)


@pytest.mark.skipif(M.device_count() > 0, reason="only meaningful on a box without a GPU")
def test_no_gpu_means_loud_failure_not_fallback():
    m = T.load_mesh("sphere40")
    with pytest.raises(M.MallieB200Error) as e:
        M.Scene(m["vertices"], m["faces"])
    assert "error -3" in str(e.value)            # MB200_ERR_NO_DEVICE
    with pytest.raises(M.MallieB200Error) as e:
        M.Scene.build(m["vertices"], m["faces"])   # device-side build + layout: same answer, no host substitute
    assert "error -3" in str(e.value)
    with pytest.raises(M.MallieB200Error) as e:
        M.HostBVH.build_device(m["vertices"], m["faces"])
    assert "error -2" in str(e.value)            # MB200_ERR_CUDA
    assert capi.launches_issued() == 0


def test_device_builder_rejects_bad_arguments_before_touching_the_gpu():
    m = T.load_mesh("sphere40")
    v, f = m["vertices"], m["faces"]
    for call in (lambda: M.HostBVH.build_device(v, f, min_leaf=1), lambda: M.Scene.build(v, f, min_leaf=1),
                 lambda: M.HostBVH.build_device(v, f, bin_size=1), lambda: M.Scene.build(v, f, bin_size=70000),
                 lambda: M.HostBVH.build_device(v, np.array([[0, 1, 10 ** 6]], np.uint32)),
                 lambda: M.Scene.build(v, np.array([[0, 1, 10 ** 6]], np.uint32))):
        with pytest.raises(M.MallieB200Error) as e:
            call()
        assert "error -1" in str(e.value)        # MB200_ERR_INVALID_ARG
    L = capi.lib()
    assert L.mb200_scene_build(None, 0, None, 0, None, 0, None, None, None, None, None) == -1
    assert L.mb200_bvh_build_device(None, 0, None, 0, None, 0, None) == -1
    assert L.mb200_scene_clone(None, None, 0) == -1 and L.mb200_scene_layout(None, None, None, None) == -1


@pytest.mark.parametrize("mesh", ["cornellbox", "teapot", "sphere40"])
def test_device_layout_is_the_reference_tree(mesh):
    """The host-side re-layout (device/layout.h: pair nodes with both children's boxes, leaf-order triangle records)
    walked in lock-step with the reference-layout tree it was made from."""
    m = T.load_mesh(mesh)
    hb = M.HostBVH.build(m["vertices"], m["faces"])
    nodes, idx = hb.arrays()
    info, pairs, tris = capi.device_layout(m["vertices"], m["faces"], nodes, idx, m["material_ids"])
    assert info["empty"] == 0 and info["tri_record_bytes"] == 48 and info["num_tri_records"] == len(idx)
    assert info["num_pair_nodes"] == int((nodes["flag"] == 0).sum()) and info["depth"] == hb.stats()["maxTreeDepth"]
    assert pairs.dtype.itemsize == 128 and (info["root_cnt"] == 0xFFFFFFFF) == (nodes[0]["flag"] == 0)
    v, f = m["vertices"], m["faces"]
    mats = m["material_ids"] if m["material_ids"] is not None else np.full(len(f), 0xFFFFFFFF, np.uint32)
    # triangle records: leaf (indices_) order, float-exact vertices, faceID + materialID
    assert np.array_equal(tris["face"], idx)
    for k, name in enumerate(("p0", "p1", "p2")):
        assert np.array_equal(tris[name].astype(np.float64), v[f[idx, k]])
    assert np.array_equal(tris["mat"], mats[idx])
    # lock-step walk
    stack, seen = [(0, info["root_ref"], info["root_cnt"])], 0
    while stack:
        ref_node, ref, cnt = stack.pop()
        nd = nodes[ref_node]
        if nd["flag"] == 1:
            assert cnt == nd["data"][0] and ref == nd["data"][1]
            continue
        assert cnt == 0xFFFFFFFF
        pn = pairs[ref]
        seen += 1
        assert pn["axis"] == nd["axis"]
        for c in range(2):
            child = nodes[nd["data"][c]]
            assert np.array_equal(pn["box"][c][:3], child["bmin"]) and np.array_equal(pn["box"][c][3:], child["bmax"])
            stack.append((int(nd["data"][c]), int(pn["ref"][c]), int(pn["cnt"][c])))
    assert seen == info["num_pair_nodes"]
    hb.close()


def test_device_layout_f64_records_empty_and_malformed_trees():
    rng = np.random.default_rng(2)
    v = rng.uniform(-1, 1, (300, 3))                       # not float-representable: 80-byte records with edges
    f = np.arange(300, dtype=np.uint32).reshape(100, 3)
    hb = M.HostBVH.build(v, f)
    nodes, idx = hb.arrays()
    info, pairs, tris = capi.device_layout(v, f, nodes, idx)
    assert info["tri_record_bytes"] == 80 and np.array_equal(tris["face"], idx) and (tris["mat"] == 0xFFFFFFFF).all()
    assert np.array_equal(tris["p0"], v[f[idx, 0]])
    assert np.array_equal(tris["e1"], v[f[idx, 1]] - v[f[idx, 0]]) and np.array_equal(tris["e2"], v[f[idx, 2]] - v[f[idx, 0]])
    info0, p0, t0 = capi.device_layout(np.zeros((0, 3)), np.zeros((0, 3), np.uint32), np.zeros(0, capi.NODE_DTYPE),
                                       np.zeros(0, np.uint32))
    assert info0["empty"] == 1 and len(p0) == 0 and len(t0) == 0
    for breakage in ("cycle", "child_oob", "leaf_oob", "axis", "index_oob"):
        bad, bidx = nodes.copy(), idx.copy()
        b = int(np.nonzero(bad["flag"] == 0)[0][0])
        leaf = int(np.nonzero(bad["flag"] == 1)[0][0])
        if breakage == "cycle":
            bad["data"][b][1] = b
        elif breakage == "child_oob":
            bad["data"][b][0] = len(bad) + 5
        elif breakage == "leaf_oob":
            bad["data"][leaf][1] = len(idx)
        elif breakage == "axis":
            bad["axis"][b] = 3
        else:
            bidx[0] = len(f)
        with pytest.raises(M.MallieB200Error):
            capi.device_layout(v, f, bad, bidx)
    hb.close()


def test_comm_and_probe_entries_reject_bad_arguments_without_a_gpu():
    """The NCCL gather entries (include/mallie_b200.h: mb200_comm_*, mb200_gather_framebuffer,
    mb200_render_frame_gathered) and mb200_probe_peaks validate their arguments before touching CUDA / NCCL."""
    L = capi.lib()
    h = C.c_void_p()
    ident = (C.c_ubyte * 128)()
    assert L.mb200_comm_init(C.byref(h), None, 2, 0, ident) == -1 and not h.value
    assert L.mb200_comm_init(None, None, 2, 0, ident) == -1
    assert L.mb200_comm_adopt(C.byref(h), None, None) == -1
    assert L.mb200_comm_unique_id(None) == -1
    assert L.mb200_comm_size(None) == 0 and L.mb200_comm_rank(None) == -1 and L.mb200_comm_exchange_path(None) == -1
    L.mb200_comm_destroy(None)
    assert L.mb200_gather_framebuffer(None, 8, 8, 3, 4, None, None) == -1
    p = capi.RenderParams()
    L.mb200_render_params_default(C.byref(p), 16, 16)
    assert L.mb200_render_frame_gathered(None, C.byref(p), 1, 4, None, None, None) == -1
    assert b"null" in L.mb200_last_error()
    pk = capi.Peaks()
    assert L.mb200_probe_peaks(0, None) == -1
    if capi.device_count() == 0:
        assert L.mb200_probe_peaks(0, C.byref(pk)) == -3          # MB200_ERR_NO_DEVICE: there is no CPU path
