"""Shared helpers for the test-suite: golden fixtures, scenes, comparisons."""
import functools
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

from mallie_b200.procedural import bumpy_sphere  # noqa: E402
from oracle import orabind as O  # noqa: E402


@functools.lru_cache(maxsize=None)
def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as fp:
        return json.load(fp)


@functools.lru_cache(maxsize=None)
def load_mesh(name):
    """name in {cornellbox, teapot, sphere40, sphere500}: dict(vertices f64, faces, material_ids, normals, uvs)."""
    if name.startswith("sphere"):
        v, f = bumpy_sphere(int(name[len("sphere"):]))
        return dict(vertices=v, faces=f, material_ids=None, normals=None, uvs=None)
    z = np.load(os.path.join(GOLDEN, f"{name}_mesh.npz"))
    return dict(vertices=z["vertices"].astype(np.float64), faces=z["faces"],
                material_ids=z["material_ids"] if "material_ids" in z else None,
                normals=z["normals"] if "normals" in z else None, uvs=z["uvs"] if "uvs" in z else None)


def sample(name):
    z = np.load(os.path.join(GOLDEN, f"{name}_hits_sample.npz"))
    return z["index"], z["hits"], z["mask"], z["isects"]


def oracle_mesh(m):
    return O.Mesh(m["vertices"], m["faces"], m["material_ids"], m["normals"], m["uvs"])


@functools.lru_cache(maxsize=None)
def oracle_scene(name):
    om = oracle_mesh(load_mesh(name))
    return om, O.BVH.build(om)


def hexf(xs):
    return np.array([float.fromhex(x) for x in xs])


def golden_frame(entry):
    f = entry["frame"]
    return hexf(f["origin"]), hexf(f["corner"]), hexf(f["du"]), hexf(f["dv"])


def mask_leaf_axis(nodes):
    n = nodes.copy()
    n["axis"][n["flag"] == 1] = 0
    return n


def fnv(a):
    return "%016x" % O.fnv1a64(np.ascontiguousarray(a))


def assert_hits_equal(got, want, what=""):
    """Bit-exact comparison of 32-byte hit records; materialID only where something was hit."""
    hit = want["faceID"] != 0xFFFFFFFF
    for f in ("t", "u", "v", "faceID"):
        a, b = np.ascontiguousarray(got[f]), np.ascontiguousarray(want[f])
        if a.tobytes() != b.tobytes():
            bad = np.nonzero(a.view(np.uint64 if a.dtype.itemsize == 8 else np.uint32) !=
                             b.view(np.uint64 if b.dtype.itemsize == 8 else np.uint32))[0]
            raise AssertionError(f"{what}: field {f} differs on {bad.size} rays, first {bad[:5]}: "
                                 f"{a[bad[:5]]} vs {b[bad[:5]]}")
    assert np.array_equal(got["materialID"][hit], want["materialID"][hit]), f"{what}: materialID differs"


def random_rays(rng, n, bmin, bmax):
    """Incoherent rays: origins on a shell around the box, directions towards random points inside it."""
    c = 0.5 * (np.asarray(bmin) + np.asarray(bmax))
    ext = 0.5 * (np.asarray(bmax) - np.asarray(bmin))
    r = float(np.linalg.norm(ext)) * 2.0 + 1e-3
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    org = c + d * r
    tgt = c + rng.uniform(-1.0, 1.0, size=(n, 3)) * ext * 1.2
    dr = tgt - org
    dr /= np.linalg.norm(dr, axis=1, keepdims=True)
    return np.concatenate([org, dr], axis=1)


# ---- builder edge cases shared by tests/golden/make_build_golden.py (reference fingerprints), the CPU tests of the
# host builder / oracle and the GPU tests of the device builder
def soup(n, seed, scale=1.0, offset=0.0, tri=0.05):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1, 1, (n, 1, 3))
    v = ((c + rng.uniform(-tri, tri, (n, 3, 3))) * scale + offset).reshape(-1, 3)
    return v.astype(np.float32).astype(np.float64), np.arange(3 * n, dtype=np.uint32).reshape(n, 3)


def doubled_grid(n=64):
    """A regular grid, every triangle twice: exactly equal bounds / centroid sums, ties everywhere."""
    g = np.arange(n + 1, dtype=np.float64)
    vv = np.stack([np.repeat(g, n + 1), np.tile(g, n + 1), np.zeros((n + 1) ** 2)], axis=1)
    q = np.array([[i * (n + 1) + j, i * (n + 1) + j + 1, (i + 1) * (n + 1) + j] for i in range(n) for j in range(n)], np.uint32)
    return vv, np.concatenate([q, q[::-1]])


def build_cases():
    """name -> (vertices, faces); default BVHBuildOptions (what the reference harness builds with)."""
    cases = {
        "soup_overlapping": soup(5000, 1),
        "soup_huge_triangles": soup(5000, 2, tri=0.8),
        "soup_tiny_extent": soup(3000, 3, scale=1e-12),
        "soup_large_negative": soup(3000, 4, scale=1e6, offset=-3e6),
        "soup_15": soup(15, 8), "soup_16": soup(16, 9), "soup_17": soup(17, 10), "soup_1": soup(1, 11),
        "doubled_grid": doubled_grid(),
        "identical_100": (np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float64), np.tile(np.array([[0, 1, 2]], np.uint32), (100, 1))),
    }
    vv, q = doubled_grid()
    cases["doubled_grid_yz"] = (vv[:, [2, 0, 1]].copy(), q)
    return cases


def build_option_cases():
    """(mesh name, BVHBuildOptions as keyword arguments): non-default options the reference's builder was run with for
    tests/golden/build_golden.json (key "name|k=v,...").  min_leaf=1 makes the reference hang a depth-maxTreeDepth chain
    of empty left leaves under every single-triangle range; only small meshes carry it."""
    opts = [dict(min_leaf=2, bin_size=8), dict(min_leaf=4), dict(max_depth=3), dict(max_depth=5), dict(bin_size=2),
            dict(bin_size=16), dict(bin_size=128, cost_taabb=0.5), dict(bin_size=1024), dict(cost_taabb=0.0),
            dict(cost_taabb=5.0, min_leaf=8)]
    out = [(n, o) for n in ("soup_overlapping", "doubled_grid", "soup_large_negative", "identical_100") for o in opts]
    out += [(n, dict(min_leaf=1)) for n in ("soup_1", "soup_17", "identical_100")]
    out += [("soup_17", dict(min_leaf=1, max_depth=9)), ("soup_overlapping", dict(min_leaf=1, max_depth=14))]
    return out


def build_option_key(name, opt):
    return name + "|" + ",".join(f"{k}={opt[k]}" for k in sorted(opt))


@functools.lru_cache(maxsize=None)
def build_golden():
    with open(os.path.join(GOLDEN, "build_golden.json")) as fp:
        return json.load(fp)


def tree_fingerprint(nodes, idx):
    return dict(num_nodes=int(len(nodes)), nodes_fnv=fnv(mask_leaf_axis(nodes)), indices_fnv=fnv(idx))


def edge_case_rays(mesh="cornellbox"):
    """Axis-parallel directions (1/0 = inf, 0*inf = NaN in the slab test), -0.0 components, origins exactly on box
    planes / corners, rays along mesh edges, through vertices and in a triangle's plane, un-normalised, zero and huge
    values.  Shared by the oracle-vs-reference test (CPU) and the GPU-vs-oracle test."""
    _, ob = oracle_scene(mesh)
    m = load_mesh(mesh)
    v, f = m["vertices"], m["faces"]
    rng = np.random.default_rng(3)
    rays = []
    axes = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], np.float64)
    for vert in v[rng.choice(len(v), 300, replace=False)]:   # through vertices, axis-parallel
        for a in axes:
            rays.append(np.concatenate([vert - 50.0 * a, a]))
            rays.append(np.concatenate([vert - 50.0 * a, np.where(a == 0, -0.0, a)]))   # -0.0 components
    nodes, _ = ob.arrays()
    for nd in nodes[:60]:                                     # origins exactly on box planes / corners
        for a in axes:
            rays.append(np.concatenate([nd["bmin"], a]))
            rays.append(np.concatenate([nd["bmax"], -a]))
            rays.append(np.concatenate([[nd["bmin"][0], 0.3, 40.0], [0.0, 0.0, -1.0]]))
    for tri in f[rng.choice(len(f), 300, replace=False)]:     # along edges and at edge midpoints
        p0, p1, p2 = v[tri[0]], v[tri[1]], v[tri[2]]
        mid = 0.5 * (p0 + p1)
        org = np.array([0.0, 0.0, 20.0])
        for tgt in (p0, mid, (p0 + p1 + p2) / 3.0):
            d = tgt - org
            rays.append(np.concatenate([org, d / np.linalg.norm(d)]))
            rays.append(np.concatenate([org, d]))             # un-normalised direction
        e = p1 - p0
        if np.linalg.norm(e) > 0:
            rays.append(np.concatenate([p0 - e, e]))          # in the triangle's plane (det ~ 0)
    rays.append(np.array([0, 0, 20, 0, 0, 0], np.float64))    # zero direction: inv = inf, nothing hit
    rays.append(np.array([0, 0, 20, 0, 0, -0.0], np.float64))
    rays.append(np.array([1e300, 0, 0, -1, 0, 0], np.float64))
    return np.array(rays)
