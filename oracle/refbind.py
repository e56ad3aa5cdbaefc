"""ctypes binding of oracle/_ref/libmallie_ref.so -- TEST INFRASTRUCTURE.

The library is the UNMODIFIED Mallie reference (compiled from /root/reference
by oracle/Makefile) behind oracle/ref_harness.cc.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (mallie_b200/) never does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libmallie_ref.so")

HIT_DTYPE = np.dtype([("t", "<f8"), ("u", "<f8"), ("v", "<f8"),
                      ("faceID", "<u4"), ("materialID", "<u4")])
# intersection.h:6-24 (184 bytes)
ISECT_DTYPE = np.dtype([("t", "<f8"), ("u", "<f8"), ("v", "<f8"),
                        ("faceID", "<u4"), ("materialID", "<u4"),
                        ("f0", "<u4"), ("f1", "<u4"), ("f2", "<u4"), ("_pad", "<u4"),
                        ("position", "<f8", 3), ("geometricNormal", "<f8", 3),
                        ("normal", "<f8", 3), ("tangent", "<f8", 3),
                        ("binormal", "<f8", 3), ("texcoord", "<f8", 2)])
# bvh_accel.h:10-29 (64 bytes)
NODE_DTYPE = np.dtype([("bmin", "<f8", 3), ("bmax", "<f8", 3), ("flag", "<i4"),
                       ("axis", "<i4"), ("data", "<u4", 2)])

_lib = None


def available():
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("reference library missing: build with `make -C oracle ref` "
                               "(needs /root/reference)")
        L = C.CDLL(_LIB_PATH)
        vp, sz, dbl, i32 = C.c_void_p, C.c_size_t, C.c_double, C.c_int
        L.ref_scene_from_arrays.restype = vp
        L.ref_scene_from_arrays.argtypes = [vp, sz, vp, sz, vp, vp, vp]
        L.ref_scene_from_file.restype = vp
        L.ref_scene_from_file.argtypes = [C.c_char_p, i32, dbl]
        L.ref_scene_init.restype = vp
        L.ref_scene_init.argtypes = [C.c_char_p, i32, dbl, i32, vp]
        L.ref_scene_destroy.argtypes = [vp]
        for f in ("ref_scene_num_vertices", "ref_scene_num_faces", "ref_scene_num_nodes",
                  "ref_scene_num_indices"):
            getattr(L, f).restype = sz
            getattr(L, f).argtypes = [vp]
        for f in ("ref_scene_has_normals", "ref_scene_has_uvs", "ref_scene_has_material_ids"):
            getattr(L, f).restype = i32
            getattr(L, f).argtypes = [vp]
        L.ref_scene_get_mesh.argtypes = [vp] * 6
        L.ref_scene_build.restype = dbl
        L.ref_scene_build.argtypes = [vp]
        L.ref_scene_build_opts.restype = dbl
        L.ref_scene_build_opts.argtypes = [vp, dbl, C.c_int, C.c_int, C.c_int]
        L.ref_scene_get_bvh.argtypes = [vp, vp, vp]
        L.ref_scene_get_stats.argtypes = [vp, vp]
        L.ref_scene_dump.restype = i32
        L.ref_scene_dump.argtypes = [vp, C.c_char_p]
        L.ref_scene_load.restype = i32
        L.ref_scene_load.argtypes = [vp, C.c_char_p]
        L.ref_scene_trace.restype = dbl
        L.ref_scene_trace.argtypes = [vp, vp, sz, vp, vp, vp, i32, i32, i32]
        L.ref_camera_frame.argtypes = [vp, vp, vp, dbl, vp, i32, i32, vp, vp, vp, vp]
        L.ref_camera_generate.argtypes = [vp, vp, vp, dbl, vp, i32, i32, vp, vp, sz, vp]
        L.ref_camera_generate_grid.argtypes = [vp, vp, vp, dbl, vp, i32, i32, vp]
        L.ref_camera_generate_env.argtypes = [vp, vp, vp, dbl, vp, i32, i32, vp, vp, sz, i32, vp]
        L.ref_scene_render_panoramic.restype = dbl
        L.ref_scene_render_panoramic.argtypes = [vp, i32, i32, dbl, vp, vp, vp, vp, i32, i32, vp, vp]
        L.ref_plane_intersect.argtypes = [C.c_float] * 4 + [vp, vp, sz, vp, vp, vp, vp]
        L.ref_scene_render.restype = dbl
        L.ref_scene_render.argtypes = [vp, i32, i32, dbl, vp, vp, vp, vp, i32, i32, i32, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _d3(v, n=3):
    a = np.ascontiguousarray(v, dtype=np.float64)
    assert a.size == n
    return a


class RefScene:
    """The reference's Scene (+ BVHAccel) behind the harness."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("reference scene creation failed")
        self.h = handle

    @classmethod
    def from_arrays(cls, vertices, faces, material_ids=None, normals=None, uvs=None):
        v = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1)
        f = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1)
        m = None if material_ids is None else np.ascontiguousarray(material_ids, dtype=np.uint32)
        n = None if normals is None else np.ascontiguousarray(normals, dtype=np.float64).reshape(-1)
        t = None if uvs is None else np.ascontiguousarray(uvs, dtype=np.float64).reshape(-1)
        return cls(lib().ref_scene_from_arrays(_p(v), v.size // 3, _p(f), f.size // 3, _p(m), _p(n), _p(t)))

    @classmethod
    def from_file(cls, path, scene_scale=1.0):
        kind = 1 if path.endswith(".eson") else 0
        return cls(lib().ref_scene_from_file(path.encode(), kind, float(scene_scale)))

    @classmethod
    def init(cls, path, scene_scale=1.0, scene_fit=False):
        """The reference's own Scene::Init (load + scene_fit / scene_scale + default build).  self.bounds =
        Scene::BoundingBox afterwards."""
        kind = 1 if path.endswith(".eson") else 0
        b = np.zeros(6, np.float64)
        h = lib().ref_scene_init(path.encode(), kind, float(scene_scale), int(bool(scene_fit)), _p(b))
        if not h:
            raise RuntimeError("reference Scene::Init failed for " + path)
        self = cls(h)
        self.bounds = (b[:3].copy(), b[3:].copy())
        return self

    def close(self):
        if self.h:
            lib().ref_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def mesh(self):
        L = lib()
        nv, nf = L.ref_scene_num_vertices(self.h), L.ref_scene_num_faces(self.h)
        v = np.empty((nv, 3), np.float64)
        f = np.empty((nf, 3), np.uint32)
        m = np.empty(nf, np.uint32) if L.ref_scene_has_material_ids(self.h) else None
        n = np.empty((nf, 3, 3), np.float64) if L.ref_scene_has_normals(self.h) else None
        t = np.empty((nf, 3, 2), np.float64) if L.ref_scene_has_uvs(self.h) else None
        L.ref_scene_get_mesh(self.h, _p(v), _p(f), _p(m), _p(n), _p(t))
        return dict(vertices=v, faces=f, material_ids=m, normals=n, uvs=t)

    def build(self, cost_taabb=None, min_leaf=None, max_depth=None, bin_size=None):
        """BVHAccel::Build; with any option given, explicit BVHBuildOptions (defaults 0.2 / 16 / 256 / 64)."""
        if (cost_taabb, min_leaf, max_depth, bin_size) != (None, None, None, None):
            s = lib().ref_scene_build_opts(self.h, 0.2 if cost_taabb is None else cost_taabb, 16 if min_leaf is None else min_leaf,
                                           256 if max_depth is None else max_depth, 64 if bin_size is None else bin_size)
        else:
            s = lib().ref_scene_build(self.h)
        if s < 0:
            raise RuntimeError("reference BVHAccel::Build failed")
        return s

    def bvh(self):
        L = lib()
        nodes = np.zeros(L.ref_scene_num_nodes(self.h), NODE_DTYPE)
        idx = np.zeros(L.ref_scene_num_indices(self.h), np.uint32)
        L.ref_scene_get_bvh(self.h, _p(nodes), _p(idx))
        return nodes, idx

    def stats(self):
        o = np.zeros(3, np.int32)
        lib().ref_scene_get_stats(self.h, _p(o))
        return dict(maxTreeDepth=int(o[0]), numLeafNodes=int(o[1]), numBranchNodes=int(o[2]))

    def dump(self, path):
        return bool(lib().ref_scene_dump(self.h, path.encode()))

    def load(self, path):
        return bool(lib().ref_scene_load(self.h, path.encode()))

    def trace(self, rays, full=False, row=1920, nthreads=0, repeat=1):
        r = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        n = r.shape[0]
        hits = np.zeros(n, HIT_DTYPE)
        isects = np.zeros(n, ISECT_DTYPE) if full else None
        mask = np.zeros(n, np.uint8)
        sec = lib().ref_scene_trace(self.h, _p(r), n, _p(hits), _p(isects), _p(mask), row, nthreads, repeat)
        return dict(hits=hits, isects=isects, mask=mask.astype(bool), seconds=sec)

    def render(self, width, height, eye, lookat, up=(0, 1, 0), quat=(0, 0, 0, 0), fov=45.0,
               plane=False, step=1, nthreads=0, count=None):
        img = np.zeros((height, width, 3), np.float32)
        cnt = np.zeros((height, width), np.int32) if count is None else np.ascontiguousarray(count, np.int32)
        sec = lib().ref_scene_render(self.h, width, height, float(fov), _p(_d3(eye)), _p(_d3(lookat)),
                                     _p(_d3(up)), _p(_d3(quat, 4)), int(plane), step, nthreads,
                                     _p(img), _p(cnt))
        return img, cnt, sec

    def render_panoramic(self, width, height, eye, lookat, up=(0, 1, 0), quat=(0, 0, 0, 0), fov=45.0, stereo=False,
                         nthreads=0, count=None):
        img = np.zeros((height, width, 3), np.float32)
        cnt = np.zeros((height, width), np.int32) if count is None else np.ascontiguousarray(count, np.int32)
        sec = lib().ref_scene_render_panoramic(self.h, width, height, float(fov), _p(_d3(eye)), _p(_d3(lookat)),
                                               _p(_d3(up)), _p(_d3(quat, 4)), int(stereo), nthreads, _p(img), _p(cnt))
        return img, cnt, sec


def camera_generate_env(eye, lookat, up, fov, quat, width, height, px, py, stereo=False):
    px = np.ascontiguousarray(px, np.float64)
    py = np.ascontiguousarray(py, np.float64)
    rays = np.zeros((px.size, 6))
    lib().ref_camera_generate_env(_p(_d3(eye)), _p(_d3(lookat)), _p(_d3(up)), float(fov), _p(_d3(quat, 4)),
                                  width, height, _p(px), _p(py), px.size, int(stereo), _p(rays))
    return rays


def camera_frame(eye, lookat, up, fov, quat, width, height):
    o, c, du, dv = (np.zeros(3) for _ in range(4))
    lib().ref_camera_frame(_p(_d3(eye)), _p(_d3(lookat)), _p(_d3(up)), float(fov), _p(_d3(quat, 4)),
                           width, height, _p(o), _p(c), _p(du), _p(dv))
    return o, c, du, dv


def camera_generate(eye, lookat, up, fov, quat, width, height, px, py):
    px = np.ascontiguousarray(px, np.float64)
    py = np.ascontiguousarray(py, np.float64)
    rays = np.zeros((px.size, 6))
    lib().ref_camera_generate(_p(_d3(eye)), _p(_d3(lookat)), _p(_d3(up)), float(fov), _p(_d3(quat, 4)),
                              width, height, _p(px), _p(py), px.size, _p(rays))
    return rays


def camera_grid(eye, lookat, up, fov, quat, width, height):
    rays = np.zeros((height * width, 6))
    lib().ref_camera_generate_grid(_p(_d3(eye)), _p(_d3(lookat)), _p(_d3(up)), float(fov),
                                   _p(_d3(quat, 4)), width, height, _p(rays))
    return rays


def plane_intersect(abcd, rays, t_in):
    r = np.ascontiguousarray(rays, np.float64).reshape(-1, 6)
    t_in = np.ascontiguousarray(t_in, np.float64)
    n = r.shape[0]
    t_out, pos, nrm, hit = np.zeros(n), np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n, np.uint8)
    lib().ref_plane_intersect(*[float(x) for x in abcd], _p(r), _p(t_in), n, _p(t_out), _p(pos), _p(nrm), _p(hit))
    return t_out, pos, nrm, hit.astype(bool)


def hdr_to_ldr(image, count):
    """The reference's HDRToLDR (main_console.cc:34-43) itself: float[H,W,3] + int[H,W] -> uint8[H,W,3]."""
    img = np.ascontiguousarray(image, np.float32)
    cnt = np.ascontiguousarray(count, np.int32)
    h, w = cnt.shape
    out = np.zeros((h, w, 3), np.uint8)
    fn = lib().ref_hdr_to_ldr
    fn.argtypes = [C.c_void_p] * 2 + [C.c_int] * 2 + [C.c_void_p]
    fn.restype = None
    fn(_p(img), _p(cnt), w, h, _p(out))
    return out


def fnv1a64(data: bytes) -> int:
    """FNV-1a-64 as used for SURVEY App. B goldens (vectorised per byte is slow; use C-ish loop in numpy)."""
    h = 14695981039346656037
    prime = 1099511628211
    mask = (1 << 64) - 1
    for b in data:
        h = ((h ^ b) * prime) & mask
    return h
