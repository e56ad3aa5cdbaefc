"""ctypes binding of oracle/liboracle.so (oracle/mallie_oracle.c) -- TEST INFRASTRUCTURE.

Plain-C CPU restatement of Mallie's hot path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the
product package (mallie_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

HIT_DTYPE = np.dtype([("t", "<f8"), ("u", "<f8"), ("v", "<f8"),
                      ("faceID", "<u4"), ("materialID", "<u4")])
ISECT_DTYPE = np.dtype([("t", "<f8"), ("u", "<f8"), ("v", "<f8"),
                        ("faceID", "<u4"), ("materialID", "<u4"),
                        ("f0", "<u4"), ("f1", "<u4"), ("f2", "<u4"), ("_pad", "<u4"),
                        ("position", "<f8", 3), ("geometricNormal", "<f8", 3),
                        ("normal", "<f8", 3), ("tangent", "<f8", 3),
                        ("binormal", "<f8", 3), ("texcoord", "<f8", 2)])
NODE_DTYPE = np.dtype([("bmin", "<f8", 3), ("bmax", "<f8", 3), ("flag", "<i4"),
                       ("axis", "<i4"), ("data", "<u4", 2)])
assert HIT_DTYPE.itemsize == 32 and ISECT_DTYPE.itemsize == 184 and NODE_DTYPE.itemsize == 64


class _Mesh(C.Structure):
    _fields_ = [("num_vertices", C.c_size_t), ("num_faces", C.c_size_t),
                ("vertices", C.c_void_p), ("faces", C.c_void_p), ("material_ids", C.c_void_p),
                ("fv_normals", C.c_void_p), ("fv_uvs", C.c_void_p)]


class _RenderParams(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int),
                ("origin", C.c_double * 3), ("corner", C.c_double * 3),
                ("du", C.c_double * 3), ("dv", C.c_double * 3),
                ("use_plane", C.c_int), ("plane", C.c_float * 4),
                ("max_path_length", C.c_int), ("rng_mode", C.c_int), ("pass_", C.c_uint32),
                ("skip_zombies", C.c_int), ("shader", C.c_int), ("light", C.c_double * 3), ("camera_mode", C.c_int)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        vp, sz, dbl, i32 = C.c_void_p, C.c_size_t, C.c_double, C.c_int
        L.ora_bvh_build.restype = vp
        L.ora_bvh_build.argtypes = [C.POINTER(_Mesh), dbl, i32, i32, i32]
        L.ora_bvh_from_arrays.restype = vp
        L.ora_bvh_from_arrays.argtypes = [vp, sz, vp, sz]
        L.ora_bvh_free.argtypes = [vp]
        L.ora_bvh_num_nodes.restype = sz
        L.ora_bvh_num_nodes.argtypes = [vp]
        L.ora_bvh_num_indices.restype = sz
        L.ora_bvh_num_indices.argtypes = [vp]
        L.ora_bvh_nodes.restype = vp
        L.ora_bvh_nodes.argtypes = [vp]
        L.ora_bvh_indices.restype = vp
        L.ora_bvh_indices.argtypes = [vp]
        L.ora_bvh_stats.argtypes = [vp, vp]
        L.ora_bvh_dump.restype = i32
        L.ora_bvh_dump.argtypes = [vp, C.c_char_p]
        L.ora_bvh_load.restype = vp
        L.ora_bvh_load.argtypes = [C.c_char_p]
        L.ora_trace_batch.restype = dbl
        L.ora_trace_batch.argtypes = [vp, C.POINTER(_Mesh), vp, sz, vp, vp, vp, vp, i32, i32]
        L.ora_occluded_batch.argtypes = [vp, C.POINTER(_Mesh), vp, vp, sz, vp, i32]
        L.ora_camera_frame.argtypes = [vp, vp, vp, dbl, vp, i32, i32, vp, vp, vp, vp]
        L.ora_generate_grid.argtypes = [vp, vp, vp, vp, i32, i32, vp]
        L.ora_generate_ray.argtypes = [vp, vp, vp, vp, dbl, dbl, vp]
        L.ora_generate_env_ray.argtypes = [vp, i32, i32, dbl, dbl, i32, vp]
        L.ora_plane_intersect.restype = i32
        L.ora_plane_intersect.argtypes = [vp, vp, vp, vp]
        L.ora_plane_from_bbox.argtypes = [vp, vp, vp]
        L.ora_render_pass.argtypes = [vp, C.POINTER(_Mesh), C.POINTER(_RenderParams), i32, i32, i32, i32,
                                      vp, vp, vp, i32]
        L.ora_render_pass_ex.argtypes = [vp, C.POINTER(_Mesh), C.POINTER(_RenderParams), i32, i32, i32, i32,
                                         vp, vp, vp, i32, vp, vp]
        L.ora_render_panoramic.argtypes = [vp, C.POINTER(_Mesh), C.POINTER(_RenderParams), vp, vp, i32]
        L.ora_fnv1a64.restype = C.c_uint64
        L.ora_fnv1a64.argtypes = [vp, sz, C.c_uint64]
        L.ora_rng_seed_pixel.argtypes = [vp, C.c_uint32, C.c_uint32]
        L.ora_rng_seed_reference.argtypes = [vp, i32]
        L.ora_randomreal.restype = dbl
        L.ora_randomreal.argtypes = [vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def fnv1a64(arr, seed=0):
    a = np.ascontiguousarray(arr)
    return int(lib().ora_fnv1a64(_p(a), a.nbytes, seed))


class Mesh:
    """Holds numpy arrays alive + the C view (mesh.h:7-18)."""

    def __init__(self, vertices, faces, material_ids=None, normals=None, uvs=None):
        self.vertices = np.ascontiguousarray(vertices, np.float64).reshape(-1, 3)
        self.faces = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
        self.material_ids = None if material_ids is None else np.ascontiguousarray(material_ids, np.uint32)
        self.normals = None if normals is None else np.ascontiguousarray(normals, np.float64).reshape(-1, 3, 3)
        self.uvs = None if uvs is None else np.ascontiguousarray(uvs, np.float64).reshape(-1, 3, 2)
        self.c = _Mesh(self.vertices.shape[0], self.faces.shape[0], _p(self.vertices), _p(self.faces),
                       _p(self.material_ids), _p(self.normals), _p(self.uvs))


class BVH:
    def __init__(self, handle, mesh):
        if not handle:
            raise RuntimeError("oracle BVH creation failed")
        self.h = handle
        self.mesh = mesh

    @classmethod
    def build(cls, mesh, cost_taabb=0.2, min_leaf=16, max_depth=256, bin_size=64):
        return cls(lib().ora_bvh_build(C.byref(mesh.c), cost_taabb, min_leaf, max_depth, bin_size), mesh)

    @classmethod
    def from_arrays(cls, nodes, indices, mesh):
        n = np.ascontiguousarray(nodes)
        assert n.dtype.itemsize == 64
        i = np.ascontiguousarray(indices, np.uint32)
        return cls(lib().ora_bvh_from_arrays(_p(n), n.shape[0], _p(i), i.shape[0]), mesh)

    @classmethod
    def load(cls, path, mesh):
        return cls(lib().ora_bvh_load(path.encode()), mesh)

    def dump(self, path):
        return bool(lib().ora_bvh_dump(self.h, path.encode()))

    def __del__(self):
        try:
            if self.h:
                lib().ora_bvh_free(self.h)
                self.h = None
        except Exception:
            pass

    def arrays(self):
        L = lib()
        nn, ni = L.ora_bvh_num_nodes(self.h), L.ora_bvh_num_indices(self.h)
        nodes = np.zeros(nn, NODE_DTYPE)
        idx = np.zeros(ni, np.uint32)
        if nn:
            C.memmove(_p(nodes), L.ora_bvh_nodes(self.h), nn * 64)
        if ni:
            C.memmove(_p(idx), L.ora_bvh_indices(self.h), ni * 4)
        return nodes, idx

    def stats(self):
        o = np.zeros(3, np.int32)
        lib().ora_bvh_stats(self.h, _p(o))
        return dict(maxTreeDepth=int(o[0]), numLeafNodes=int(o[1]), numBranchNodes=int(o[2]))

    def trace(self, rays, full=False, row=1920, nthreads=0):
        r = np.ascontiguousarray(rays, np.float64).reshape(-1, 6)
        n = r.shape[0]
        hits = np.zeros(n, HIT_DTYPE)
        isects = np.zeros(n, ISECT_DTYPE) if full else None
        mask = np.zeros(n, np.uint8)
        totals = np.zeros(3, np.uint64)
        sec = lib().ora_trace_batch(self.h, C.byref(self.mesh.c), _p(r), n, _p(hits), _p(isects), _p(mask),
                                    _p(totals), row, nthreads)
        return dict(hits=hits, isects=isects, mask=mask.astype(bool), seconds=sec,
                    n_node=int(totals[0]), n_tri=int(totals[1]), max_stack=int(totals[2]))

    def occluded(self, rays, tmax, nthreads=0):
        r = np.ascontiguousarray(rays, np.float64).reshape(-1, 6)
        t = np.ascontiguousarray(tmax, np.float64)
        out = np.zeros(r.shape[0], np.uint8)
        lib().ora_occluded_batch(self.h, C.byref(self.mesh.c), _p(r), _p(t), r.shape[0], _p(out), nthreads)
        return out.astype(bool)

    def render_pass(self, frame, width, height, plane=None, max_path_length=16, rng_mode=1, pass_index=0,
                    skip_zombies=1, shader=0, light=(0.0, 0.0, 0.0), tile=None, image=None, count=None,
                    nthreads=0, emit_rays=False, camera_mode=0):
        p = _RenderParams()
        p.width, p.height = width, height
        o, c, du, dv = frame
        for k in range(3):
            p.origin[k], p.corner[k], p.du[k], p.dv[k] = o[k], c[k], du[k], dv[k]
            p.light[k] = light[k]
        p.use_plane = 0 if plane is None else 1
        if plane is not None:
            for k in range(4):
                p.plane[k] = plane[k]
        p.max_path_length, p.rng_mode, p.pass_ = max_path_length, rng_mode, pass_index
        p.skip_zombies, p.shader, p.camera_mode = skip_zombies, shader, camera_mode
        if image is None:
            image = np.zeros((height, width, 3), np.float32)
        if count is None:
            count = np.zeros((height, width), np.int32)
        x0, y0, x1, y1 = tile if tile is not None else (0, 0, width, height)
        rc = np.zeros(7, np.uint64)
        prim = np.zeros((height * width, 6)) if emit_rays else None
        shad = np.zeros((height * width, 7)) if emit_rays else None
        lib().ora_render_pass_ex(self.h, C.byref(self.mesh.c), C.byref(p), x0, y0, x1, y1, _p(image), _p(count),
                                 _p(rc), nthreads, _p(prim), _p(shad))
        info = dict(trace_calls=int(rc[0]), zombies=int(rc[1]), shadow_rays=int(rc[2]), n_node=int(rc[3]),
                    n_tri=int(rc[4]), shadow_n_node=int(rc[5]), shadow_n_tri=int(rc[6]))
        if emit_rays:
            info["primary_rays"] = prim
            ok = ~np.isnan(shad[:, 0])
            info["shadow_rays_buf"] = np.ascontiguousarray(shad[ok, :6])
            info["shadow_tmax"] = np.ascontiguousarray(shad[ok, 6])
        return image, count, info


def render_panoramic(bvh, origin, width, height, stereo=False, max_path_length=16, rng_mode=1, pass_index=0,
                     nthreads=0):
    """RenderPanoramic (render.cc:710-763): 10 samples of PathTraceEnv per pixel, count += 10."""
    p = _RenderParams()
    p.width, p.height = width, height
    for k in range(3):
        p.origin[k] = origin[k]
    p.max_path_length, p.rng_mode, p.pass_, p.skip_zombies = max_path_length, rng_mode, pass_index, 1
    p.shader, p.camera_mode = 2, (2 if stereo else 1)
    image = np.zeros((height, width, 3), np.float32)
    count = np.zeros((height, width), np.int32)
    lib().ora_render_panoramic(bvh.h, C.byref(bvh.mesh.c), C.byref(p), _p(image), _p(count), nthreads)
    return image, count


def camera_frame(eye, lookat, up=(0, 1, 0), fov=45.0, quat=(0, 0, 0, 0), width=512, height=512):
    e, l, u = (np.ascontiguousarray(x, np.float64) for x in (eye, lookat, up))
    q = np.ascontiguousarray(quat, np.float64)
    o, c, du, dv = (np.zeros(3) for _ in range(4))
    lib().ora_camera_frame(_p(e), _p(l), _p(u), float(fov), _p(q), width, height, _p(o), _p(c), _p(du), _p(dv))
    return o, c, du, dv


def generate_env(origin, width, height, px, py, stereo=False):
    """Camera::GenerateEnvRay / GenerateStereoEnvRay for arrays of pixel coordinates."""
    o = np.ascontiguousarray(origin, np.float64)
    px, py = np.asarray(px, np.float64).reshape(-1), np.asarray(py, np.float64).reshape(-1)
    rays = np.zeros((px.size, 6))
    tmp = np.zeros(6)
    for i in range(px.size):
        lib().ora_generate_env_ray(_p(o), width, height, float(px[i]), float(py[i]), int(stereo), _p(tmp))
        rays[i] = tmp
    return rays


def generate_grid(frame, width, height):
    o, c, du, dv = (np.ascontiguousarray(x, np.float64) for x in frame)
    rays = np.zeros((height * width, 6))
    lib().ora_generate_grid(_p(o), _p(c), _p(du), _p(dv), width, height, _p(rays))
    return rays


def generate_rays(frame, px, py):
    o, c, du, dv = (np.ascontiguousarray(x, np.float64) for x in frame)
    px = np.asarray(px, np.float64).reshape(-1)
    py = np.asarray(py, np.float64).reshape(-1)
    rays = np.zeros((px.size, 6))
    tmp = np.zeros(6)
    for i in range(px.size):
        lib().ora_generate_ray(_p(o), _p(c), _p(du), _p(dv), float(px[i]), float(py[i]), _p(tmp))
        rays[i] = tmp
    return rays


def plane_from_bbox(bmin, bmax):
    a = np.zeros(4, np.float32)
    lib().ora_plane_from_bbox(_p(np.ascontiguousarray(bmin, np.float64)), _p(np.ascontiguousarray(bmax, np.float64)),
                              _p(a))
    return a


def plane_intersect(abcd, rays, t_in):
    r = np.ascontiguousarray(rays, np.float64).reshape(-1, 6)
    abcd = np.ascontiguousarray(abcd, np.float32)
    n = r.shape[0]
    t_out, pos, nrm, hit = np.zeros(n), np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n, bool)
    is_ = np.zeros(1, ISECT_DTYPE)
    for i in range(n):
        is_[:] = 0
        is_["t"] = t_in[i]
        h = lib().ora_plane_intersect(_p(abcd), _p(r[i, :3].copy()), _p(r[i, 3:].copy()), _p(is_))
        hit[i] = bool(h)
        t_out[i] = is_["t"][0]
        pos[i] = is_["position"][0]
        nrm[i] = is_["normal"][0]
    return t_out, pos, nrm, hit


def hdr_to_ldr(image, count):
    """HDRToLDR (main_console.cc:34-43): float[H,W,3] + int[H,W] -> uint8[H,W,3]."""
    img, cnt = np.ascontiguousarray(image, np.float32), np.ascontiguousarray(count, np.int32)
    h, w = cnt.shape
    out = np.zeros((h, w, 3), np.uint8)
    fn = lib().ora_hdr_to_ldr
    fn.argtypes, fn.restype = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p], None
    fn(_p(img), _p(cnt), w, h, _p(out))
    return out


def display_bgra(image, count):
    """Display (main_sdl.cc:420-477): float[H,W,3] + int[H,W] -> uint8[H,W,4] BGRA, gamma 2.2."""
    img, cnt = np.ascontiguousarray(image, np.float32), np.ascontiguousarray(count, np.int32)
    h, w = cnt.shape
    out = np.zeros((h, w, 4), np.uint8)
    fn = lib().ora_display_bgra
    fn.argtypes, fn.restype = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p], None
    fn(_p(img), _p(cnt), w, h, _p(out))
    return out


def rng_stream(pixel, pass_index, n, reference_tid=None):
    st = np.zeros(4, np.uint32)
    if reference_tid is None:
        lib().ora_rng_seed_pixel(_p(st), pixel, pass_index)
    else:
        lib().ora_rng_seed_reference(_p(st), reference_tid)
    return np.array([lib().ora_randomreal(_p(st)) for _ in range(n)])
