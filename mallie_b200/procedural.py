"""Procedural benchmark meshes.

Input generation only -- not part of the hot path.  The "bumpy sphere" is the
~1 M / ~10 M triangle scene BASELINE.json's HBM-roofline runs are quoted on:
a UV sphere with a 5 % sinusoidal radius perturbation, vertex positions computed
in double and rounded to float (so every coordinate is exactly float-representable,
as with Mallie's OBJ loader, importers/tiny_obj_loader.cc:696-697).
"""
import numpy as np


def bumpy_sphere(n):
    """N=500 -> exactly 1 000 000 triangles / 501 501 vertices; N=1581 -> 9 998 244 triangles.

    Returns (vertices float64 [nv,3] holding float32-exact values, faces uint32 [nf,3]).
    """
    nu, nv = 2 * n, n
    i = np.arange(nu + 1, dtype=np.float64)
    j = np.arange(nv + 1, dtype=np.float64)
    u = (2.0 * np.pi * i / nu)[None, :]
    v = (np.pi * j / nv)[:, None]
    r = (1.0 + 0.05 * np.sin(16.0 * u) * np.sin(12.0 * v)).astype(np.float32).astype(np.float64)
    x = (r * np.sin(v) * np.cos(u)).astype(np.float32)
    y = (r * np.cos(v) * np.ones_like(u)).astype(np.float32)
    z = (r * np.sin(v) * np.sin(u)).astype(np.float32)
    verts = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float64)
    jj, ii = np.meshgrid(np.arange(nv, dtype=np.uint32), np.arange(nu, dtype=np.uint32), indexing="ij")
    a = (jj * (nu + 1) + ii).reshape(-1)
    b = a + 1
    c = a + nu + 1
    d = c + 1
    faces = np.empty((a.size, 2, 3), np.uint32)
    faces[:, 0, 0], faces[:, 0, 1], faces[:, 0, 2] = a, c, b
    faces[:, 1, 0], faces[:, 1, 1], faces[:, 1, 2] = b, c, d
    return verts, faces.reshape(-1, 3)
