"""ctypes binding of the mallie_b200 C ABI (include/mallie_b200.h).

This is the Python face of the product: tests and bench.py drive the CUDA path
through exactly the entry points a Mallie `#ifdef ENABLE_B200` build would call
(INTEGRATION.md).  There is NO fallback: if libmallie_b200.so is missing or no
B200 is visible the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmallie_b200.so")

RAY_DTYPE = np.dtype([("org", "<f8", 3), ("dir", "<f8", 3)])
HIT_DTYPE = np.dtype([("t", "<f8"), ("u", "<f8"), ("v", "<f8"), ("faceID", "<u4"), ("materialID", "<u4")])
ISECT_DTYPE = np.dtype([("t", "<f8"), ("u", "<f8"), ("v", "<f8"), ("faceID", "<u4"), ("materialID", "<u4"),
                        ("f0", "<u4"), ("f1", "<u4"), ("f2", "<u4"), ("_pad", "<u4"),
                        ("position", "<f8", 3), ("geometricNormal", "<f8", 3), ("normal", "<f8", 3),
                        ("tangent", "<f8", 3), ("binormal", "<f8", 3), ("texcoord", "<f8", 2)])
NODE_DTYPE = np.dtype([("bmin", "<f8", 3), ("bmax", "<f8", 3), ("flag", "<i4"), ("axis", "<i4"),
                       ("data", "<u4", 2)])
assert RAY_DTYPE.itemsize == 48 and HIT_DTYPE.itemsize == 32
assert ISECT_DTYPE.itemsize == 184 and NODE_DTYPE.itemsize == 64

SHADER_PATHTRACE, SHADER_PRIMARY_SHADOW, SHADER_PRIMARY_ONLY, SHADER_PATHTRACE_ENV = 0, 1, 2, 3
LDR_RGB8_LINEAR, LDR_BGRA8_GAMMA22 = 0, 1
CAMERA_PINHOLE, CAMERA_ENV, CAMERA_ENV_STEREO = 0, 1, 2

# Every symbol include/mallie_b200.h declares (tests/test_abi.py checks the header against this list).
EXPORTS = [
    "mb200_last_error", "mb200_version", "mb200_device_count", "mb200_launches_issued",
    "mb200_build_options_default", "mb200_bvh_build", "mb200_bvh_build_device", "mb200_scene_build", "mb200_scene_layout", "mb200_scene_clone", "mb200_bvh_load", "mb200_bvh_dump",
    "mb200_bvh_num_nodes", "mb200_bvh_num_indices", "mb200_bvh_nodes", "mb200_bvh_indices",
    "mb200_bvh_stats", "mb200_bvh_destroy",
    "mb200_bvh_device_layout",
    "mb200_scene_create", "mb200_scene_destroy", "mb200_scene_bounds", "mb200_scene_device_bytes",
    "mb200_scene_stream", "mb200_scene_device", "mb200_scene_uses_f32_vertices", "mb200_scene_synchronize",
    "mb200_scene_timing", "mb200_scene_kernel_times", "mb200_probe_peaks",
    "mb200_trace_closest", "mb200_trace_closest_full", "mb200_trace_occluded", "mb200_trace_closest_async",
    "mb200_camera_frame_build", "mb200_generate_rays", "mb200_generate_rays_env", "mb200_generate_rays_grid",
    "mb200_render_params_default", "mb200_plane_from_bounds", "mb200_render_pass", "mb200_render_accumulate",
    "mb200_render_frame", "mb200_render_frame_multi", "mb200_band_local_rows", "mb200_resolve_ldr", "mb200_render_frame_ldr",
    "mb200_comm_unique_id", "mb200_comm_init", "mb200_comm_adopt", "mb200_comm_size", "mb200_comm_rank", "mb200_comm_exchange_path",
    "mb200_comm_destroy", "mb200_gather_framebuffer", "mb200_render_frame_gathered",
    "mb200_mesh_load_obj", "mb200_mesh_load_eson", "mb200_mesh_transform", "mb200_mesh_num_vertices",
    "mb200_mesh_num_faces", "mb200_mesh_vertices", "mb200_mesh_faces", "mb200_mesh_material_ids",
    "mb200_mesh_fv_normals", "mb200_mesh_fv_uvs", "mb200_mesh_destroy", "mb200_config_default", "mb200_config_load",
]


class BuildOptions(C.Structure):
    _fields_ = [("cost_taabb", C.c_double), ("min_leaf_primitives", C.c_int), ("max_tree_depth", C.c_int),
                ("bin_size", C.c_int)]


class BuildStats(C.Structure):
    _fields_ = [("max_tree_depth", C.c_int), ("num_leaf_nodes", C.c_int), ("num_branch_nodes", C.c_int)]


class Counters(C.Structure):
    _fields_ = [("nodes_tested", C.c_uint64), ("tris_tested", C.c_uint64), ("rays", C.c_uint64),
                ("max_stack", C.c_uint64)]


class CameraFrame(C.Structure):
    _fields_ = [("origin", C.c_double * 3), ("corner", C.c_double * 3), ("du", C.c_double * 3),
                ("dv", C.c_double * 3)]

    def arrays(self):
        return tuple(np.array(list(getattr(self, k))) for k in ("origin", "corner", "du", "dv"))


class RenderParams(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("x0", C.c_int), ("y0", C.c_int), ("x1", C.c_int),
                ("y1", C.c_int), ("frame", CameraFrame), ("use_plane", C.c_int), ("plane", C.c_float * 4),
                ("max_path_length", C.c_int), ("pass_", C.c_uint32), ("jitter", C.c_int), ("shader", C.c_int),
                ("light", C.c_double * 3),
                ("band_rows", C.c_int), ("band_count", C.c_int), ("band_index", C.c_int), ("band_compact", C.c_int),
                ("pixel_step", C.c_int), ("camera_mode", C.c_int)]


class RenderStats(C.Structure):
    _fields_ = [("primary_rays", C.c_uint64), ("bounce_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("zombie_segments", C.c_uint64), ("camera_nodes_tested", C.c_uint64), ("camera_tris_tested", C.c_uint64),
                ("shadow_nodes_tested", C.c_uint64), ("shadow_tris_tested", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class Config(C.Structure):
    """mb200_config: struct RenderConfig (render.h:11-49) as a POD."""
    _fields_ = [("fov", C.c_double), ("width", C.c_int), ("height", C.c_int), ("eye", C.c_double * 3),
                ("lookat", C.c_double * 3), ("up", C.c_double * 3), ("quat", C.c_double * 4),
                ("scene_scale", C.c_double), ("scene_fit", C.c_int), ("plane", C.c_int), ("num_passes", C.c_int),
                ("num_photons", C.c_int), ("obj_filename", C.c_char * 1024), ("eson_filename", C.c_char * 1024),
                ("magicavoxel_filename", C.c_char * 1024), ("material_filename", C.c_char * 1024),
                ("max_path_length", C.c_int), ("shader", C.c_int), ("light", C.c_double * 3), ("device", C.c_int),
                ("num_gpus", C.c_int)]


class KernelTimes(C.Structure):
    _fields_ = [(k + "_ms", C.c_double) for k in ("camera_trace", "shadow_trace", "bounce_trace", "shade", "resolve",
                                                  "query_trace")] + \
               [(k + "_launches", C.c_uint64) for k in ("camera_trace", "shadow_trace", "bounce_trace", "shade",
                                                        "resolve", "query_trace")] + [("trace_union_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Peaks(C.Structure):
    _fields_ = [("fp64_lane_ops_per_s", C.c_double), ("l2_read_bytes_per_s", C.c_double),
                ("hbm_copy_bytes_per_s", C.c_double), ("sm_count", C.c_int)]


class LayoutInfo(C.Structure):
    _fields_ = [("num_pair_nodes", C.c_uint64), ("num_tri_records", C.c_uint64), ("tri_record_bytes", C.c_uint32),
                ("root_ref", C.c_uint32), ("root_cnt", C.c_uint32), ("depth", C.c_int32), ("empty", C.c_int32)]


PAIR_DTYPE = np.dtype([("box", "<f8", (2, 6)), ("ref", "<u4", (2,)), ("cnt", "<u4", (2,)), ("axis", "<u4"),
                       ("pad", "<u4", (3,))])
TRI32_DTYPE = np.dtype([("p0", "<f4", (3,)), ("face", "<u4"), ("p1", "<f4", (3,)), ("mat", "<u4"), ("p2", "<f4", (3,)),
                        ("pad", "<u4")])
TRI64_DTYPE = np.dtype([("p0", "<f8", (3,)), ("e1", "<f8", (3,)), ("e2", "<f8", (3,)), ("face", "<u4"), ("mat", "<u4")])


class MallieB200Error(RuntimeError):
    pass


_lib = None


def lib():
    """Loads libmallie_b200.so.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MallieB200Error(
                f"{LIB_PATH} is missing: build it with `make -C mallie_b200/csrc` "
                "(or python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp, sz, i32, dbl = C.c_void_p, C.c_size_t, C.c_int, C.c_double
        L.mb200_last_error.restype = C.c_char_p
        L.mb200_version.restype = C.c_char_p
        L.mb200_device_count.restype = i32
        L.mb200_launches_issued.restype = i32
        L.mb200_build_options_default.argtypes = [C.POINTER(BuildOptions)]
        L.mb200_bvh_build.argtypes = [C.POINTER(vp), vp, sz, vp, sz, C.POINTER(BuildOptions)]
        L.mb200_bvh_build_device.argtypes = [C.POINTER(vp), C.c_int, vp, sz, vp, sz, C.POINTER(BuildOptions)]
        L.mb200_scene_build.argtypes = [C.POINTER(vp), C.c_int, vp, sz, vp, sz, vp, vp, vp, C.POINTER(BuildOptions), C.POINTER(vp)]
        L.mb200_scene_clone.argtypes = [C.POINTER(vp), vp, C.c_int]
        L.mb200_scene_layout.argtypes = [vp, C.POINTER(LayoutInfo), vp, vp]
        L.mb200_bvh_load.argtypes = [C.POINTER(vp), C.c_char_p]
        L.mb200_bvh_dump.argtypes = [vp, C.c_char_p]
        L.mb200_bvh_num_nodes.restype = sz
        L.mb200_bvh_num_nodes.argtypes = [vp]
        L.mb200_bvh_num_indices.restype = sz
        L.mb200_bvh_num_indices.argtypes = [vp]
        L.mb200_bvh_nodes.restype = vp
        L.mb200_bvh_nodes.argtypes = [vp]
        L.mb200_bvh_indices.restype = vp
        L.mb200_bvh_indices.argtypes = [vp]
        L.mb200_bvh_stats.argtypes = [vp, C.POINTER(BuildStats)]
        L.mb200_bvh_destroy.argtypes = [vp]
        L.mb200_bvh_device_layout.argtypes = [vp, sz, vp, sz, vp, vp, sz, vp, sz, C.POINTER(LayoutInfo), vp, vp]
        L.mb200_scene_create.argtypes = [C.POINTER(vp), i32, vp, sz, vp, sz, vp, vp, vp, vp, sz, vp, sz]
        L.mb200_scene_destroy.argtypes = [vp]
        L.mb200_scene_bounds.argtypes = [vp, vp, vp]
        L.mb200_scene_device_bytes.restype = sz
        L.mb200_scene_device_bytes.argtypes = [vp]
        L.mb200_scene_stream.restype = vp
        L.mb200_scene_stream.argtypes = [vp]
        L.mb200_scene_device.argtypes = [vp]
        L.mb200_scene_uses_f32_vertices.argtypes = [vp]
        L.mb200_scene_synchronize.argtypes = [vp]
        L.mb200_scene_timing.argtypes = [vp, i32]
        L.mb200_scene_kernel_times.argtypes = [vp, C.POINTER(KernelTimes)]
        L.mb200_probe_peaks.argtypes = [i32, C.POINTER(Peaks)]
        L.mb200_trace_closest.argtypes = [vp, vp, sz, vp, C.POINTER(Counters)]
        L.mb200_trace_closest_full.argtypes = [vp, vp, sz, vp, vp]
        L.mb200_trace_occluded.argtypes = [vp, vp, vp, sz, vp, C.POINTER(Counters)]
        L.mb200_trace_closest_async.argtypes = [vp, vp, sz, vp]
        L.mb200_camera_frame_build.argtypes = [C.POINTER(CameraFrame), vp, vp, vp, dbl, vp, i32, i32]
        L.mb200_generate_rays.argtypes = [vp, C.POINTER(CameraFrame), vp, vp, sz, vp]
        L.mb200_generate_rays_grid.argtypes = [vp, C.POINTER(CameraFrame), i32, i32, i32, i32, vp]
        L.mb200_generate_rays_env.argtypes = [vp, vp, i32, i32, vp, vp, sz, i32, vp]
        L.mb200_render_params_default.argtypes = [C.POINTER(RenderParams), i32, i32]
        L.mb200_plane_from_bounds.argtypes = [vp, vp, vp]
        L.mb200_render_pass.argtypes = [vp, C.POINTER(RenderParams), vp, vp, C.POINTER(RenderStats)]
        L.mb200_render_accumulate.argtypes = [vp, C.POINTER(RenderParams), i32, vp, vp, C.POINTER(RenderStats)]
        L.mb200_render_frame.argtypes = [vp, C.POINTER(RenderParams), i32, vp, vp, C.POINTER(RenderStats)]
        L.mb200_band_local_rows.argtypes = [C.POINTER(RenderParams)]
        L.mb200_render_frame_multi.argtypes = [C.POINTER(vp), i32, C.POINTER(RenderParams), i32, i32, vp, vp,
                                               C.POINTER(RenderStats)]
        L.mb200_resolve_ldr.argtypes = [vp, vp, vp, i32, i32, i32, vp]
        L.mb200_render_frame_ldr.argtypes = [vp, C.POINTER(RenderParams), i32, i32, vp, C.POINTER(RenderStats)]
        L.mb200_comm_unique_id.argtypes = [vp]
        L.mb200_comm_init.argtypes = [C.POINTER(vp), vp, i32, i32, vp]
        L.mb200_comm_adopt.argtypes = [C.POINTER(vp), vp, vp]
        L.mb200_comm_size.argtypes = [vp]
        L.mb200_comm_rank.argtypes = [vp]
        L.mb200_comm_exchange_path.argtypes = [vp]
        L.mb200_comm_destroy.argtypes = [vp]
        L.mb200_gather_framebuffer.argtypes = [vp, i32, i32, i32, i32, vp, vp]
        L.mb200_render_frame_gathered.argtypes = [vp, C.POINTER(RenderParams), i32, i32, vp, vp, C.POINTER(RenderStats)]
        L.mb200_mesh_load_obj.argtypes = [C.POINTER(vp), C.c_char_p]
        L.mb200_mesh_load_eson.argtypes = [C.POINTER(vp), C.c_char_p]
        L.mb200_mesh_transform.argtypes = [vp, dbl, i32]
        for fn in ("num_vertices", "num_faces"):
            getattr(L, "mb200_mesh_" + fn).restype = sz
            getattr(L, "mb200_mesh_" + fn).argtypes = [vp]
        for fn in ("vertices", "faces", "material_ids", "fv_normals", "fv_uvs"):
            getattr(L, "mb200_mesh_" + fn).restype = vp
            getattr(L, "mb200_mesh_" + fn).argtypes = [vp]
        L.mb200_mesh_destroy.argtypes = [vp]
        L.mb200_config_default.argtypes = [C.POINTER(Config)]
        L.mb200_config_load.argtypes = [C.POINTER(Config), C.c_char_p, C.c_char_p]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise MallieB200Error(f"mallie_b200 error {rc}: {lib().mb200_last_error().decode()}")


def _p(a):
    """numpy array -> void*; int -> device pointer passthrough; None -> NULL."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def device_count():
    return int(lib().mb200_device_count())


def launches_issued():
    return int(lib().mb200_launches_issued())


def probe_peaks(device=0):
    """mb200_probe_peaks: measured FP64-pipe, L2-read and HBM-copy ceilings of the GPU (diagnostics)."""
    pk = Peaks()
    check(lib().mb200_probe_peaks(device, C.byref(pk)))
    return {k: getattr(pk, k) for k, _ in pk._fields_}


def render_frame_multi(scenes, params, num_passes, band_rows=8, image=None, count=None, stats=True):
    """mb200_render_frame_multi: one frame over several single-GPU Scene replicas from one host thread."""
    if image is None:
        image = np.zeros((params.height, params.width, 3), np.float32)
    if count is None:
        count = np.zeros((params.height, params.width), np.int32)
    arr = (C.c_void_p * len(scenes))(*[s.h for s in scenes])
    st = RenderStats()
    check(lib().mb200_render_frame_multi(arr, len(scenes), C.byref(params), num_passes, band_rows, _p(image), _p(count),
                                         C.byref(st) if stats else None))
    return image, count, (st.as_dict() if stats else None)


class Comm:
    """mb200_comm: the NCCL communicator of a frame split over one process per GPU (one Scene replica each)."""

    @staticmethod
    def _load_order():
        # the library binds NCCL with dlopen("libnccl.so.2"); if this process is going to use PyTorch as well, torch's
        # bundled copy has to be the one in the process, so it is loaded first (see csrc/device/gather.cu)
        try:
            import torch  # noqa: F401
        except ImportError:
            pass

    def __init__(self, scene, nranks, rank, unique_id):
        self._load_order()
        self.h = None
        self.scene = scene
        h = C.c_void_p()
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        check(lib().mb200_comm_init(C.byref(h), scene.h, nranks, rank, buf))
        self.h, self.nranks, self.rank = h, nranks, rank

    @staticmethod
    def unique_id():
        Comm._load_order()
        buf = (C.c_ubyte * 128)()
        check(lib().mb200_comm_unique_id(buf))
        return bytes(buf)

    def gather(self, width, height, band_rows, d_bands, image, channels=3):
        """d_bands: device address (int) of this rank's compact band buffer; image: device address, numpy array or None."""
        check(lib().mb200_gather_framebuffer(self.h, width, height, channels, band_rows, _p(d_bands), _p(image)))

    def render_frame(self, params, num_passes, band_rows, image=None, count=None, stats=False):
        """mb200_render_frame_gathered: image / count may be device addresses (ints), numpy arrays or None."""
        st = RenderStats()
        check(lib().mb200_render_frame_gathered(self.h, C.byref(params), num_passes, band_rows, _p(image), _p(count),
                                                C.byref(st) if stats else None))
        return st.as_dict() if stats else None

    def exchange_path(self):
        """1: the last gathered frame went through the peer-memory exchange kernel, 0: through ncclAllGather."""
        return int(lib().mb200_comm_exchange_path(self.h))

    def close(self):
        if self.h:
            lib().mb200_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_layout(vertices, faces, nodes, indices, material_ids=None):
    """mb200_bvh_device_layout: (info dict, pair nodes, triangle records) -- what scene creation would upload."""
    v = np.ascontiguousarray(vertices, np.float64).reshape(-1, 3)
    f = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
    n = np.ascontiguousarray(nodes)
    i = np.ascontiguousarray(indices, np.uint32)
    m = None if material_ids is None else np.ascontiguousarray(material_ids, np.uint32)
    info = LayoutInfo()
    args = (_p(v), v.shape[0], _p(f), f.shape[0], _p(m), _p(n), n.shape[0], _p(i), i.shape[0])
    check(lib().mb200_bvh_device_layout(*args, C.byref(info), None, None))
    pairs = np.zeros(info.num_pair_nodes, PAIR_DTYPE)
    tris = np.zeros(info.num_tri_records, TRI32_DTYPE if info.tri_record_bytes == 48 else TRI64_DTYPE)
    check(lib().mb200_bvh_device_layout(*args, C.byref(info), _p(pairs), _p(tris)))
    return {k: int(getattr(info, k)) for k, _ in info._fields_}, pairs, tris


def load_config(path=None, text=None):
    """LoadJSONConfig (main.cc:98-205) -> Config; raises MallieB200Error when the file is unreadable / malformed."""
    cfg = Config()
    lib().mb200_config_default(C.byref(cfg))
    check(lib().mb200_config_load(C.byref(cfg), path.encode() if path is not None else None,
                                  text.encode() if text is not None else None))
    return cfg


def load_mesh(path, kind=None, scene_scale=1.0, scene_fit=False):
    """MeshLoader::LoadObj / LoadESON + Scene::Init's scale / fit.  Returns dict(vertices, faces, material_ids,
    normals, uvs) of numpy arrays (normals / uvs None when the loader leaves them NULL)."""
    L = lib()
    kind = kind or ("eson" if path.endswith(".eson") else "obj")
    h = C.c_void_p()
    check((L.mb200_mesh_load_eson if kind == "eson" else L.mb200_mesh_load_obj)(C.byref(h), path.encode()))
    try:
        check(L.mb200_mesh_transform(h, float(scene_scale), int(bool(scene_fit))))
        nv, nf = L.mb200_mesh_num_vertices(h), L.mb200_mesh_num_faces(h)

        def grab(fn, dtype, count):
            ptr = getattr(L, "mb200_mesh_" + fn)(h)
            if not ptr or count == 0:
                return None if fn.startswith("fv_") else np.zeros(0, dtype)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), (count,)).copy()

        return dict(vertices=grab("vertices", np.float64, 3 * nv).reshape(-1, 3),
                    faces=grab("faces", np.uint32, 3 * nf).reshape(-1, 3),
                    material_ids=grab("material_ids", np.uint32, nf),
                    normals=grab("fv_normals", np.float64, 9 * nf), uvs=grab("fv_uvs", np.float64, 6 * nf))
    finally:
        L.mb200_mesh_destroy(h)


# ----------------------------------------------------------------------------------------------
class HostBVH:
    """mb200_bvh: the host-built, reference-layout tree (BVHAccel::Build / Dump / Load)."""

    def __init__(self, handle):
        self.h = handle

    @classmethod
    def build(cls, vertices, faces, cost_taabb=0.2, min_leaf=16, max_depth=256, bin_size=64):
        v = np.ascontiguousarray(vertices, np.float64).reshape(-1, 3)
        f = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
        opt = BuildOptions(cost_taabb, min_leaf, max_depth, bin_size)
        h = C.c_void_p()
        check(lib().mb200_bvh_build(C.byref(h), _p(v), v.shape[0], _p(f), f.shape[0], C.byref(opt)))
        return cls(h)

    @classmethod
    def build_device(cls, vertices, faces, device=0, cost_taabb=0.2, min_leaf=16, max_depth=256, bin_size=64):
        """The same tree, grown level by level on GPU `device` (mb200_bvh_build_device)."""
        v = np.ascontiguousarray(vertices, np.float64).reshape(-1, 3)
        f = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
        opt = BuildOptions(cost_taabb, min_leaf, max_depth, bin_size)
        h = C.c_void_p()
        check(lib().mb200_bvh_build_device(C.byref(h), device, _p(v), v.shape[0], _p(f), f.shape[0], C.byref(opt)))
        return cls(h)

    @classmethod
    def load(cls, path):
        h = C.c_void_p()
        check(lib().mb200_bvh_load(C.byref(h), path.encode()))
        return cls(h)

    def dump(self, path):
        check(lib().mb200_bvh_dump(self.h, path.encode()))

    def arrays(self):
        L = lib()
        nn, ni = L.mb200_bvh_num_nodes(self.h), L.mb200_bvh_num_indices(self.h)
        nodes = np.zeros(nn, NODE_DTYPE)
        idx = np.zeros(ni, np.uint32)
        if nn:
            C.memmove(_p(nodes), L.mb200_bvh_nodes(self.h), nn * 64)
        if ni:
            C.memmove(_p(idx), L.mb200_bvh_indices(self.h), ni * 4)
        return nodes, idx

    def stats(self):
        s = BuildStats()
        check(lib().mb200_bvh_stats(self.h, C.byref(s)))
        return dict(maxTreeDepth=s.max_tree_depth, numLeafNodes=s.num_leaf_nodes, numBranchNodes=s.num_branch_nodes)

    def close(self):
        if self.h:
            lib().mb200_bvh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def camera_frame(eye, lookat, up=(0, 1, 0), fov=45.0, quat=(0, 0, 0, 0), width=512, height=512):
    f = CameraFrame()
    e, l, u = (np.ascontiguousarray(x, np.float64) for x in (eye, lookat, up))
    q = np.ascontiguousarray(quat, np.float64)
    check(lib().mb200_camera_frame_build(C.byref(f), _p(e), _p(l), _p(u), float(fov), _p(q), width, height))
    return f


def plane_from_bounds(bmin, bmax):
    out = np.zeros(4, np.float32)
    lib().mb200_plane_from_bounds(_p(np.ascontiguousarray(bmin, np.float64)),
                                  _p(np.ascontiguousarray(bmax, np.float64)), _p(out))
    return out


class Scene:
    """mb200_scene: the device-resident scene (what Scene::Init builds, scene.cc:224-230)."""

    def __init__(self, vertices, faces, material_ids=None, normals=None, uvs=None, nodes=None, indices=None,
                 device=0):
        self.h = None
        self.vertices = np.ascontiguousarray(vertices, np.float64).reshape(-1, 3)
        self.faces = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
        m = None if material_ids is None else np.ascontiguousarray(material_ids, np.uint32)
        n = None if normals is None else np.ascontiguousarray(normals, np.float64).reshape(-1)
        t = None if uvs is None else np.ascontiguousarray(uvs, np.float64).reshape(-1)
        if nodes is None:
            bvh = HostBVH.build(self.vertices, self.faces)
            nodes, indices = bvh.arrays()
            bvh.close()
        self.nodes = np.ascontiguousarray(nodes)
        self.indices = np.ascontiguousarray(indices, np.uint32)
        assert self.nodes.dtype.itemsize == 64
        h = C.c_void_p()
        check(lib().mb200_scene_create(C.byref(h), device, _p(self.vertices), self.vertices.shape[0], _p(self.faces),
                                       self.faces.shape[0], _p(m), _p(n), _p(t), _p(self.nodes),
                                       self.nodes.shape[0], _p(self.indices), self.indices.shape[0]))
        self.h = h

    @classmethod
    def build(cls, vertices, faces, material_ids=None, normals=None, uvs=None, device=0, want_bvh=True,
              cost_taabb=0.2, min_leaf=16, max_depth=256, bin_size=64):
        """mb200_scene_build: BVH build + traversal layout entirely on the GPU.  With want_bvh the reference-layout
        tree is downloaded too (self.nodes / self.indices)."""
        self = cls.__new__(cls)
        self.h = None
        self.vertices = np.ascontiguousarray(vertices, np.float64).reshape(-1, 3)
        self.faces = np.ascontiguousarray(faces, np.uint32).reshape(-1, 3)
        m = None if material_ids is None else np.ascontiguousarray(material_ids, np.uint32)
        n = None if normals is None else np.ascontiguousarray(normals, np.float64).reshape(-1)
        t = None if uvs is None else np.ascontiguousarray(uvs, np.float64).reshape(-1)
        opt = BuildOptions(cost_taabb, min_leaf, max_depth, bin_size)
        h, b = C.c_void_p(), C.c_void_p()
        check(lib().mb200_scene_build(C.byref(h), device, _p(self.vertices), self.vertices.shape[0], _p(self.faces),
                                      self.faces.shape[0], _p(m), _p(n), _p(t), C.byref(opt),
                                      C.byref(b) if want_bvh else None))
        self.h = h
        self.nodes = self.indices = None
        if want_bvh:
            bvh = HostBVH(b)
            self.nodes, self.indices = bvh.arrays()
            self.build_stats = bvh.stats()
            bvh.close()
        return self

    def clone(self, device):
        """mb200_scene_clone: a replica on GPU `device`, copied device to device."""
        other = Scene.__new__(Scene)
        other.h = None
        other.vertices, other.faces = self.vertices, self.faces
        other.nodes, other.indices = getattr(self, "nodes", None), getattr(self, "indices", None)
        h = C.c_void_p()
        check(lib().mb200_scene_clone(C.byref(h), self.h, device))
        other.h = h
        return other

    def layout(self):
        """mb200_scene_layout: (info dict, pair nodes, triangle records) as resident on the device."""
        info = LayoutInfo()
        check(lib().mb200_scene_layout(self.h, C.byref(info), None, None))
        pairs = np.zeros(info.num_pair_nodes, PAIR_DTYPE)
        tris = np.zeros(info.num_tri_records, TRI32_DTYPE if info.tri_record_bytes == 48 else TRI64_DTYPE)
        check(lib().mb200_scene_layout(self.h, C.byref(info), _p(pairs), _p(tris)))
        return {k: int(getattr(info, k)) for k, _ in info._fields_}, pairs, tris

    def close(self):
        if self.h:
            lib().mb200_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- info
    def bounds(self):
        a, b = np.zeros(3), np.zeros(3)
        check(lib().mb200_scene_bounds(self.h, _p(a), _p(b)))
        return a, b

    def device_bytes(self):
        return int(lib().mb200_scene_device_bytes(self.h))

    def stream(self):
        return int(lib().mb200_scene_stream(self.h) or 0)

    def uses_f32_vertices(self):
        return bool(lib().mb200_scene_uses_f32_vertices(self.h))

    def synchronize(self):
        check(lib().mb200_scene_synchronize(self.h))

    def timing(self, enable=True):
        """Bracket every kernel this scene launches with CUDA events (mb200_scene_timing)."""
        check(lib().mb200_scene_timing(self.h, int(enable)))

    def kernel_times(self):
        """ms and launch counts per kernel class since the last call (synchronises the stream)."""
        kt = KernelTimes()
        check(lib().mb200_scene_kernel_times(self.h, C.byref(kt)))
        return kt.as_dict()

    # -- queries (numpy in / numpy out, or raw device pointers as ints)
    @staticmethod
    def _rays(rays):
        r = np.ascontiguousarray(rays, np.float64).reshape(-1, 6)
        return r, r.shape[0]

    def trace_closest(self, rays, counters=False):
        r, n = self._rays(rays)
        hits = np.zeros(n, HIT_DTYPE)
        c = Counters()
        check(lib().mb200_trace_closest(self.h, _p(r), n, _p(hits), C.byref(c) if counters else None))
        if counters:
            return hits, dict(nodes_tested=int(c.nodes_tested), tris_tested=int(c.tris_tested), rays=int(c.rays),
                              max_stack=int(c.max_stack))
        return hits

    def trace_closest_full(self, rays):
        r, n = self._rays(rays)
        isects = np.zeros(n, ISECT_DTYPE)
        mask = np.zeros(n, np.uint8)
        check(lib().mb200_trace_closest_full(self.h, _p(r), n, _p(isects), _p(mask)))
        return isects, mask.astype(bool)

    def trace_occluded(self, rays, tmax, counters=False):
        r, n = self._rays(rays)
        t = np.ascontiguousarray(tmax, np.float64)
        assert t.shape[0] == n
        occ = np.zeros(n, np.uint8)
        c = Counters()
        check(lib().mb200_trace_occluded(self.h, _p(r), _p(t), n, _p(occ), C.byref(c) if counters else None))
        if counters:
            return occ.astype(bool), dict(nodes_tested=int(c.nodes_tested), tris_tested=int(c.tris_tested),
                                          rays=int(c.rays), max_stack=int(c.max_stack))
        return occ.astype(bool)

    def trace_closest_device(self, d_rays, n, d_hits):
        """Enqueue-only on the scene's stream; d_rays/d_hits are device addresses (ints)."""
        check(lib().mb200_trace_closest_async(self.h, _p(int(d_rays)), n, _p(int(d_hits))))

    def generate_rays(self, frame, px, py):
        px = np.ascontiguousarray(px, np.float64).reshape(-1)
        py = np.ascontiguousarray(py, np.float64).reshape(-1)
        rays = np.zeros((px.size, 6))
        check(lib().mb200_generate_rays(self.h, C.byref(frame), _p(px), _p(py), px.size, _p(rays)))
        return rays

    def generate_rays_env(self, origin, width, height, px, py, stereo=False):
        """Camera::GenerateEnvRay / GenerateStereoEnvRay for arrays of pixel coordinates."""
        o = np.ascontiguousarray(origin, np.float64)
        px, py = np.ascontiguousarray(px, np.float64).reshape(-1), np.ascontiguousarray(py, np.float64).reshape(-1)
        rays = np.zeros((px.size, 6))
        check(lib().mb200_generate_rays_env(self.h, _p(o), width, height, _p(px), _p(py), px.size, int(stereo), _p(rays)))
        return rays

    def generate_rays_grid(self, frame, x0, y0, x1, y1, out=None):
        if out is None:
            out = np.zeros(((y1 - y0) * (x1 - x0), 6))
        check(lib().mb200_generate_rays_grid(self.h, C.byref(frame), x0, y0, x1, y1, _p(out)))
        return out

    # -- frame
    def render_params(self, frame, width, height, tile=None, plane=None, max_path_length=16, pass_index=0,
                      jitter=True, shader=SHADER_PATHTRACE, light=(0.0, 20.0, 0.0), bands=None, compact=False, step=1,
                      camera_mode=0):
        """bands = (band_rows, band_count, band_index) enables the multi-GPU row-band interleave; step is
        Render()'s coarse-preview step (render.cc:657-698)."""
        p = RenderParams()
        lib().mb200_render_params_default(C.byref(p), width, height)
        p.frame = frame
        if tile is not None:
            p.x0, p.y0, p.x1, p.y1 = tile
        if plane is not None:
            p.use_plane = 1
            for k in range(4):
                p.plane[k] = float(plane[k])
        p.max_path_length, p.pass_, p.jitter, p.shader = max_path_length, pass_index, int(jitter), shader
        for k in range(3):
            p.light[k] = float(light[k])
        if bands is not None:
            p.band_rows, p.band_count, p.band_index = bands
            p.band_compact = int(compact)
        p.pixel_step = int(step)
        p.camera_mode = int(camera_mode)
        return p

    @staticmethod
    def band_local_rows(params):
        return int(lib().mb200_band_local_rows(C.byref(params)))

    def _out_buffers(self, params, image, count):
        rows = self.band_local_rows(params) if (params.band_rows > 0 and params.band_compact) else params.height
        if image is None:
            image = np.zeros((rows, params.width, 3), np.float32)
        if count is None:
            count = np.zeros((rows, params.width), np.int32)
        return image, count

    def render_frame(self, params, num_passes, image=None, count=None, stats=True):
        """image = sum of num_passes samples, count = num_passes (both overwritten)."""
        image, count = self._out_buffers(params, image, count)
        st = RenderStats()
        check(lib().mb200_render_frame(self.h, C.byref(params), num_passes, _p(image), _p(count),
                                       C.byref(st) if stats else None))
        return image, count, (st.as_dict() if stats else None)

    def render_pass(self, params, image=None, count=None, stats=True):
        """image/count: numpy arrays (host) or ints (device addresses)."""
        if image is None:
            image = np.zeros((params.height, params.width, 3), np.float32)
        if count is None:
            count = np.zeros((params.height, params.width), np.int32)
        st = RenderStats()
        check(lib().mb200_render_pass(self.h, C.byref(params), _p(image), _p(count), C.byref(st) if stats else None))
        return image, count, (st.as_dict() if stats else None)

    def resolve_ldr(self, image, count, width, height, mode=0, out=None):
        """mb200_resolve_ldr: mode 0 = HDRToLDR (RGB8), 1 = Display (BGRA8, gamma 2.2).  image / count / out: numpy arrays
        or device addresses (ints)."""
        if out is None:
            out = np.zeros((height, width, 3 if mode == 0 else 4), np.uint8)
        check(lib().mb200_resolve_ldr(self.h, _p(image), _p(count), width, height, mode, _p(out)))
        return out

    def render_frame_ldr(self, params, num_passes, mode=0, out=None, stats=True):
        """mb200_render_frame_ldr: the frame as 8-bit pixels; the float frame stays on the GPU."""
        if out is None:
            out = np.zeros((params.height, params.width, 3 if mode == 0 else 4), np.uint8)
        st = RenderStats()
        check(lib().mb200_render_frame_ldr(self.h, C.byref(params), num_passes, mode, _p(out), C.byref(st) if stats else None))
        return out, (st.as_dict() if stats else None)

    def render_accumulate(self, params, num_passes, image=None, count=None, stats=True):
        if image is None:
            image = np.zeros((params.height, params.width, 3), np.float32)
        if count is None:
            count = np.zeros((params.height, params.width), np.int32)
        st = RenderStats()
        check(lib().mb200_render_accumulate(self.h, C.byref(params), num_passes, _p(image), _p(count),
                                            C.byref(st) if stats else None))
        return image, count, (st.as_dict() if stats else None)
