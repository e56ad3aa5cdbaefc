#!/usr/bin/env python
"""Generates tests/golden/* from the UNMODIFIED reference (oracle/_ref/libmallie_ref.so).

Run in the authoring container only (needs /root/reference for the assets and the
reference library built by `make -C oracle ref`):

    python tests/golden/make_golden.py

Outputs (committed):
  cornellbox_mesh.npz, teapot_mesh.npz   the Mesh the reference's own OBJ loader produces
                                         (MeshLoader::LoadObj, importers/mesh_loader.cc:26)
  golden.json                            per-scene pins: BVH statistics + FNV-1a-64 of nodes / indices,
                                         camera frames, hit counts and FNV-1a-64 hashes of faceID and
                                         (t,u,v) for un-jittered primary rays, spot
                                         hit records, deterministic OMP_NUM_THREADS=1 render hashes
  *_hits_sample.npz                      ~6000 evenly spaced hit records of each ray set, in full
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import orabind as O  # noqa: E402  (only for its FNV helper)
from oracle import refbind as R  # noqa: E402
from mallie_b200.procedural import bumpy_sphere  # noqa: E402

SAMPLE_TARGET = 6000  # ~6k full records per ray set


def mask_leaf_axis(nodes):
    n = nodes.copy()
    n["axis"][n["flag"] == 1] = 0  # uninitialised in the reference (bvh_accel.cc:343-360)
    return n


def scene_entry(name, rs, eye, lookat, W, H, out, spots=()):
    rs.build()
    nodes, idx = rs.bvh()
    fr = R.camera_frame(eye, lookat, (0, 1, 0), 45.0, (0, 0, 0, 0), W, H)
    rays = R.camera_grid(eye, lookat, (0, 1, 0), 45.0, (0, 0, 0, 0), W, H)
    tr = rs.trace(rays, full=True, row=W)
    h, m = tr["hits"], tr["mask"]
    e = dict(
        eye=list(eye), lookat=list(lookat), width=W, height=H,
        stats=rs.stats(), num_nodes=int(len(nodes)), num_indices=int(len(idx)),
        nodes_fnv="%016x" % O.fnv1a64(mask_leaf_axis(nodes)), indices_fnv="%016x" % O.fnv1a64(idx),
        frame=dict(origin=[float.hex(x) for x in fr[0]], corner=[float.hex(x) for x in fr[1]],
                   du=[float.hex(x) for x in fr[2]], dv=[float.hex(x) for x in fr[3]]),
        rays_fnv="%016x" % O.fnv1a64(rays),
        hits=int(m.sum()),
        faceid_fnv="%016x" % O.fnv1a64(h["faceID"]),
        tuv_fnv="%016x" % O.fnv1a64(np.stack([h["t"], h["u"], h["v"]], 1)[m]),
        isect_fnv={f: "%016x" % O.fnv1a64(np.ascontiguousarray(tr["isects"][f][m]))
                   for f in ("position", "geometricNormal", "normal", "texcoord", "materialID", "f0", "f1", "f2")},
        spots=[],
    )
    for (x, y) in spots:
        r = h[y * W + x]
        e["spots"].append(dict(x=x, y=y, faceID=int(r["faceID"]), t=float.hex(float(r["t"])),
                               u=float.hex(float(r["u"])), v=float.hex(float(r["v"]))))
    sel = np.arange(0, len(h), max(1, (len(h) // SAMPLE_TARGET) | 1))
    np.savez_compressed(os.path.join(HERE, f"{name}_hits_sample.npz"), index=sel.astype(np.uint32),
                        hits=h[sel], mask=m[sel], isects=tr["isects"][sel])
    out[name] = e
    return e


def render_hash(obj, plane, cwd):
    """One deterministic Render() (OMP_NUM_THREADS=1) in a fresh process (Render keeps static state).
    cwd decides whether the .mtl next to the .obj is found: found -> materialIDs >= 0 (throughput
    halves per bounce); not found -> materialID == -1 everywhere (the SURVEY App. B hashes)."""
    code = f"""
import os, sys; sys.path.insert(0, {ROOT!r}); os.chdir({cwd!r})
import numpy as np
from oracle import refbind as R, orabind as O
rs = R.RefScene.from_file({obj!r}); rs.build()
img, cnt, sec = rs.render(512, 512, (0,0,20), (0,0,0), plane={plane}, nthreads=1)
print("RESULT %016x %.3f %d" % (O.fnv1a64(img), img.astype(np.float64).sum(), int((img.reshape(-1,3).sum(1)!=0).sum())))
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout
    line = [l for l in out.splitlines() if l.startswith("RESULT")][0].split()
    return dict(fnv=line[1], sum=float(line[2]), nonzero=int(line[3]))


def save_mesh(name, rs):
    m = rs.mesh()
    v32 = m["vertices"].astype(np.float32)
    assert np.array_equal(v32.astype(np.float64), m["vertices"]), "OBJ positions are float-exact"
    kw = dict(vertices=v32, faces=m["faces"], material_ids=m["material_ids"])
    if m["normals"] is not None:
        kw["normals"] = m["normals"]
    if m["uvs"] is not None:
        kw["uvs"] = m["uvs"]
    np.savez_compressed(os.path.join(HERE, f"{name}_mesh.npz"), **kw)


def main():
    out = {}
    cwd = os.getcwd()
    os.chdir(REF)  # .mtl lookup is CWD-relative
    rs = R.RefScene.from_file(os.path.join(REF, "cornellbox_suzanne.obj"))
    save_mesh("cornellbox", rs)
    scene_entry("cornellbox_512", rs, (0, 0, 20), (0, 0, 0), 512, 512, out, spots=[(256, 256), (100, 400)])
    rs = R.RefScene.from_file(os.path.join(REF, "teapot.obj"))
    save_mesh("teapot", rs)
    scene_entry("teapot_1080p", rs, (5, 40, 150), (5, 40, 0), 1920, 1080, out, spots=[(960, 540), (640, 540)])
    os.chdir(cwd)
    v, f = bumpy_sphere(500)
    rs = R.RefScene.from_arrays(v, f)
    e = scene_entry("sphere500_1080p", rs, (0, 0, 3), (0, 0, 0), 1920, 1080, out, spots=[(960, 540), (700, 300)])
    e["vertices_fnv"] = "%016x" % O.fnv1a64(v)
    e["faces_fnv"] = "%016x" % O.fnv1a64(f)
    v, f = bumpy_sphere(40)
    rs = R.RefScene.from_arrays(v, f)
    scene_entry("sphere40_256", rs, (0.3, 0.2, 3), (0, 0, 0), 256, 256, out, spots=[(128, 128)])
    obj = os.path.join(REF, "cornellbox_suzanne.obj")
    out["render_cornellbox_512_1thread"] = dict(
        with_mtl=dict(plane_off=render_hash(obj, False, REF), plane_on=render_hash(obj, True, REF)),
        without_mtl=dict(plane_off=render_hash(obj, False, "/tmp"), plane_on=render_hash(obj, True, "/tmp")))
    with open(os.path.join(HERE, "golden.json"), "w") as fp:
        json.dump(out, fp, indent=1, sort_keys=True)
    print(json.dumps({k: {kk: vv for kk, vv in v.items() if kk in ("hits", "faceid_fnv", "tuv_fnv", "stats")}
                      if isinstance(v, dict) and "hits" in v else v for k, v in out.items()}, indent=1))


if __name__ == "__main__":
    main()
