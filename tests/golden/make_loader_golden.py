#!/usr/bin/env python
"""Golden meshes for the OBJ / ESON loaders, generated from the UNMODIFIED reference loaders
(MeshLoader::LoadObj / LoadESON through oracle/_ref/libmallie_ref.so).  Authoring container only:

    python tests/golden/make_loader_golden.py

Inputs (hand-written, committed): tricky.obj + tricky.mtl.  small.eson is written by this script.
Outputs (committed): tricky_mesh.npz, small_eson_mesh.npz, small.eson, loader_golden.json (hashes of what the
reference loaders make of the shipped assets, checked only where /root/reference is mounted)."""
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import orabind as O  # noqa: E402  (FNV helper only)
from oracle import refbind as R  # noqa: E402


def write_eson(path, fields):
    """LTE ESON container (importers/eson.cc): i64 total, then tag | key\\0 | payload."""
    body = b""
    for key, (tag, val) in fields.items():
        body += bytes([tag]) + key.encode() + b"\0"
        if tag == 2:
            body += struct.pack("<q", val)
        elif tag == 1:
            body += struct.pack("<d", val)
        else:
            body += struct.pack("<q", len(val)) + val
    open(path, "wb").write(struct.pack("<q", 8 + len(body)) + body)


def mesh_arrays(rs):
    m = rs.mesh()
    return {k: v for k, v in m.items() if v is not None}


def main():
    os.chdir(HERE)  # mtllib is resolved against the current directory (tiny_obj_loader.cc:604-616)
    rs = R.RefScene.from_file("tricky.obj")
    np.savez_compressed("tricky_mesh.npz", **mesh_arrays(rs))
    rs.close()

    rng = np.random.default_rng(7)
    nv, nf = 37, 50
    verts = rng.normal(size=(nv, 3)).astype(np.float32)
    faces = rng.integers(0, nv, size=(nf, 3)).astype(np.int32)
    mats = rng.integers(0, 5, size=nf).astype(np.uint16)
    write_eson("small.eson", {
        "faces": (6, faces.tobytes()), "name": (4, b"small"), "num_faces": (2, nf), "scale": (1, 1.5),
        "material_ids": (6, mats.tobytes()), "num_vertices": (2, nv), "vertices": (6, verts.tobytes()),
        "facevarying_uvs": (6, rng.normal(size=(nf, 6)).astype(np.float32).tobytes())})
    rs = R.RefScene.from_file("small.eson")
    np.savez_compressed("small_eson_mesh.npz", **mesh_arrays(rs))
    rs.close()

    pins = {}
    os.chdir("/root/reference")
    for name, fn in (("cornellbox_obj", "cornellbox_suzanne.obj"), ("teapot_obj", "teapot.obj"),
                     ("cornellbox_eson", "cornellbox_suzanne.eson"), ("cornellbox_obj_x2.5", "cornellbox_suzanne.obj")):
        rs = R.RefScene.from_file(fn, scene_scale=2.5 if name.endswith("x2.5") else 1.0)
        m = mesh_arrays(rs)
        pins[name] = {k: "%016x" % O.fnv1a64(np.ascontiguousarray(v)) for k, v in m.items()}
        pins[name]["shape"] = [int(len(m["vertices"])), int(len(m["faces"]))]
        rs.close()
    os.chdir(HERE)
    json.dump(pins, open("loader_golden.json", "w"), indent=1, sort_keys=True)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
