#!/usr/bin/env python
"""Randomised parity run (development aid, GPU): random triangle soups and grids at random scales / offsets, random
BVHBuildOptions, built on the device, traced with rays of several kinds (incoherent, axis-parallel through vertices, from
inside, along edges, un-normalised); closest hits, traversal counters and occlusion (tmax at t scaled by 0.9 / 1 / 1.1 and
at +-2 ulp) against the oracle, the device-built tree against the oracle's.  Prints one line per scene and a summary.

    python tools/fuzz_parity.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mallie_b200 as M  # noqa: E402
from oracle import orabind as O  # noqa: E402
from tests import common as T  # noqa: E402


def scene(rng):
    kind = rng.integers(0, 4)
    scale = 10.0 ** rng.uniform(-6, 6)
    offset = scale * rng.uniform(-100, 100) if rng.random() < 0.5 else 0.0
    n = int(2 ** rng.uniform(0, 13))
    if kind == 0:
        v, f = T.soup(n, int(rng.integers(1 << 30)), scale, offset, tri=float(10.0 ** rng.uniform(-3, 0)))
    elif kind == 1:                                   # grid with shared vertices (edges / vertices hit exactly), maybe doubled
        g = int(rng.integers(2, 40))
        a = np.arange(g + 1, dtype=np.float64)
        z = rng.normal(0, 0.3, (g + 1) ** 2) * (rng.random() < 0.7)
        v = np.stack([np.repeat(a, g + 1), np.tile(a, g + 1), z], axis=1) * scale + offset
        v = v.astype(np.float32).astype(np.float64)
        q = np.array([[i * (g + 1) + j, i * (g + 1) + j + 1, (i + 1) * (g + 1) + j] for i in range(g) for j in range(g)], np.uint32)
        q2 = np.array([[i * (g + 1) + j + 1, (i + 1) * (g + 1) + j + 1, (i + 1) * (g + 1) + j] for i in range(g) for j in range(g)], np.uint32)
        f = np.concatenate([q, q2] + ([q[::-1]] if rng.random() < 0.3 else []))
    elif kind == 2:                                   # degenerate: needles, zero-area and repeated triangles mixed in
        v, f = T.soup(max(n, 8), int(rng.integers(1 << 30)), scale, offset)
        k = len(f) // 4
        v = v.copy()
        v[f[:k, 2]] = v[f[:k, 1]]                     # zero area
        v[f[k:2 * k, 2]] = v[f[k:2 * k, 0]] + (v[f[k:2 * k, 1]] - v[f[k:2 * k, 0]]) * 0.5   # collinear
        f = np.concatenate([f, f[:k]])
    else:                                             # double-precision vertices (80-byte records)
        v, f = T.soup(n, int(rng.integers(1 << 30)), scale, offset)
        v = v * (1.0 + 1e-9)
    opt = {}
    if rng.random() < 0.5:
        opt = dict(min_leaf=int(rng.integers(2, 33)), bin_size=int(rng.choice([2, 4, 16, 64, 256])), max_depth=int(rng.choice([3, 8, 32, 256])),
                   cost_taabb=float(rng.choice([0.0, 0.2, 1.0, 5.0])))
    return v, f, opt, (kind, n, scale, offset)


def rays_for(rng, v, f, n):
    lo, hi = v.min(0), v.max(0)
    ext = np.maximum(hi - lo, 1e-30)
    out = [T.random_rays(rng, n, lo, hi)]
    axes = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], np.float64)
    vi = v[rng.integers(0, len(v), n // 8)]
    a = axes[rng.integers(0, 6, len(vi))]
    out.append(np.concatenate([vi - a * 2.0 * np.linalg.norm(ext), np.where(rng.random((len(vi), 3)) < 0.5, a, np.where(a == 0, -0.0, a))], axis=1))
    tri = f[rng.integers(0, len(f), n // 8)]
    p0, p1, p2 = v[tri[:, 0]], v[tri[:, 1]], v[tri[:, 2]]
    w = rng.random((len(tri), 1))
    tgt = np.where(rng.random((len(tri), 1)) < 0.5, p0 + (p1 - p0) * w, (p0 + p1 + p2) / 3.0)    # on an edge / centroid
    org = lo + rng.uniform(-0.5, 1.5, (len(tri), 3)) * ext                                         # inside and outside
    d = tgt - org
    out.append(np.concatenate([org, d * 10.0 ** rng.uniform(-3, 3, (len(tri), 1))], axis=1))     # un-normalised
    r = np.concatenate(out)
    return np.ascontiguousarray(r[np.isfinite(r).all(axis=1) & (np.abs(r[:, 3:]).sum(axis=1) > 0)])


def run(budget, seed, verbose=True):
    rng = np.random.default_rng(seed)
    t_end = time.time() + budget
    scenes = rays_total = 0
    while time.time() < t_end:
        v, f, opt, what = scene(rng)
        tag = f"kind {what[0]} tris {len(f)} scale {what[2]:.2e} offset {what[3]:.2e} opt {opt}"
        ob = O.BVH.build(O.Mesh(v, f), **opt)
        sc = M.Scene.build(v, f, **opt)
        on, oi = ob.arrays()
        assert T.mask_leaf_axis(sc.nodes).tobytes() == T.mask_leaf_axis(on).tobytes() and np.array_equal(sc.indices, oi), "tree: " + tag
        rays = rays_for(rng, v, f, 4096)
        hits, cnt = sc.trace_closest(rays, counters=True)
        o = ob.trace(rays)
        T.assert_hits_equal(hits, o["hits"], tag)
        assert cnt["nodes_tested"] == o["n_node"] and cnt["tris_tested"] == o["n_tri"], "counters: " + tag
        t = o["hits"]["t"]
        for tm in (np.where(o["mask"], t * rng.choice([0.9, 1.0, 1.1], len(rays)), 1e30),
                   np.where(o["mask"], np.nextafter(np.nextafter(t, np.inf), np.inf), 1.0),
                   np.where(o["mask"], np.nextafter(np.nextafter(t, -np.inf), -np.inf), 1.0)):
            assert np.array_equal(sc.trace_occluded(rays, tm), ob.occluded(rays, tm)), "occlusion: " + tag
        scenes += 1
        rays_total += len(rays)
        if verbose:
            print(f"ok {tag}: {int(o['mask'].sum())} hits of {len(rays)} rays", flush=True)
        sc.close()
    return scenes, rays_total


if __name__ == "__main__":
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    scenes, rays_total = run(budget, seed)
    print(f"FUZZ OK: {scenes} scenes, {rays_total} rays, seed {seed}")
