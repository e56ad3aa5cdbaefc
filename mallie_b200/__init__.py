"""mallie_b200 -- B200-native implementation of Mallie's render hot path.

The product is the C-ABI shared library `libmallie_b200.so` (include/mallie_b200.h):
host C++ (BVH build, camera frame, Scene/Render mirror of the reference API) plus
hand-written sm_100a CUDA kernels (ray generation, BVH traversal, ray/triangle
intersection, shading).  This package is only its ctypes face for tests and bench.py.
"""
from . import capi  # noqa: F401
from .capi import (HostBVH, Scene, MallieB200Error, camera_frame, plane_from_bounds, device_count,  # noqa: F401
                   load_mesh, load_config, Config, render_frame_multi, Comm,
                   SHADER_PATHTRACE, SHADER_PRIMARY_SHADOW, SHADER_PRIMARY_ONLY, SHADER_PATHTRACE_ENV,
                   CAMERA_PINHOLE, CAMERA_ENV, CAMERA_ENV_STEREO)
