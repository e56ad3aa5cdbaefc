"""numpy model of the level-synchronous BVH build the device runs (mallie_b200/csrc/device/bvh_build_gpu.cu): the same
data-parallel steps -- segmented bounds, per-segment histograms, per-plane SAH costs, the Hoare partition expressed
with prefix sums over "misplaced" flags, level-order node creation, pre-order numbering from subtree sizes -- written
with whole-array operations.  Test infrastructure: tests/test_lsbuild_model.py holds it against the host builder (which
is pinned to the reference's tree), so the ALGORITHM is checked on a box without a GPU; tests/test_gpu_build.py checks
the CUDA implementation itself.  Reference: BVHAccel::Build, bvh_accel.cc:36-482."""
import numpy as np

KP = np.finfo(np.float64).eps * 1024.0           # kEPS pad of the node bounds, bvh_accel.cc:283
NODE_DTYPE = np.dtype([("bmin", "<f8", (3,)), ("bmax", "<f8", (3,)), ("flag", "<i4"), ("axis", "<i4"), ("data", "<u4", (2,))])


def half_area2(lo3, hi3):
    dx, dy, dz = hi3[:, 0] - lo3[:, 0], hi3[:, 1] - lo3[:, 1], hi3[:, 2] - lo3[:, 2]
    return 2.0 * (dx * dy + dy * dz + dz * dx)


def build(vertices, faces, min_leaf=16, max_depth=256, nb=64, taabb=0.2):
    v = np.ascontiguousarray(vertices, np.float64)
    f = np.ascontiguousarray(faces, np.uint32)
    nt = len(f)
    p = v[f]                                                     # [T, vertex, axis]
    lo, hi = p.min(axis=1).T.copy(), p.max(axis=1).T.copy()      # [3, T]: per-triangle bounds
    cs = ((p[:, 0, :] + p[:, 1, :]) + p[:, 2, :]).T.copy()       # p0 + p1 + p2, SAHPred's left side (bvh_accel.cc:257-277)
    idx = np.arange(nt, dtype=np.uint32)
    nodes = [dict()]                                             # creation (level) order
    segs = [(0, nt, 0)]                                          # (l, r, node slot), sorted by l
    level = 0
    while segs:
        S = len(segs)
        sl = np.array([s[0] for s in segs], np.int64)
        sr = np.array([s[1] for s in segs], np.int64)
        sn = [s[2] for s in segs]
        n = sr - sl
        pos = np.concatenate([np.arange(a, b) for a, b in zip(sl, sr)])      # all active positions
        sid = np.repeat(np.arange(S), n)                                        # ... and their segment
        tri = idx[pos]
        starts = np.concatenate([[0], np.cumsum(n)[:-1]])
        # bounds (bvh_accel.cc:285-315) and the leaf test (bvh_accel.cc:341)
        bmin = np.stack([np.minimum.reduceat(lo[a][tri], starts) for a in range(3)], 1) - KP
        bmax = np.stack([np.maximum.reduceat(hi[a][tri], starts) for a in range(3)], 1) + KP
        leaf = (n < min_leaf) | (level >= max_depth)
        # histograms of the bound minima / maxima (ContributeBinBuffer, bvh_accel.cc:82-142)
        ext = bmax - bmin
        scale = np.where(ext > KP, float(nb) / np.where(ext > KP, ext, 1.0), 0.0)
        step = ext * (1.0 / nb)
        hmin = np.zeros((S, 3, nb), np.int64)
        hmax = np.zeros((S, 3, nb), np.int64)
        for a in range(3):
            qlo = np.floor((lo[a][tri] - bmin[sid, a]) * scale[sid, a]).astype(np.uint32).astype(np.int64)
            qhi = np.floor((hi[a][tri] - bmin[sid, a]) * scale[sid, a]).astype(np.uint32).astype(np.int64)
            qlo = np.where(qlo.astype(np.float64) >= nb, nb - 1, qlo)
            qhi = np.where(qhi.astype(np.float64) >= nb, nb - 1, qhi)
            np.add.at(hmin, (sid, a, qlo), 1)
            np.add.at(hmax, (sid, a, qhi), 1)
        # SAH: every plane's cost from its own prefix counts, winner = (cost, plane) minimum, lower plane on ties
        # (FindCutFromBinBuffer, bvh_accel.cc:156-255, keeps the first strict improvement)
        whole = half_area2(bmin, bmax)
        inv_whole = np.where(whole > KP, 1.0 / np.where(whole > KP, whole, 1.0), 0.0)
        t_box, t_tri = taabb, 1.0 - taabb
        best_cost = np.full((S, 3), np.finfo(np.float64).max)
        best_pos = bmin + 0.5 * step
        for a in range(3):
            nl = np.cumsum(hmin[:, a, :nb - 1], axis=1)
            nr = n[:, None] - np.cumsum(hmax[:, a, :nb - 1], axis=1)
            for i in range(nb - 1):
                plane = bmin[:, a] + (i + 0.5) * step[:, a]
                lh, rl = bmax.copy(), bmin.copy()
                lh[:, a] = plane
                rl[:, a] = plane
                al, ar = half_area2(bmin, lh), half_area2(rl, bmax)
                cost = (np.float64(np.float32(2.0) * np.float64(t_box)) + (al * inv_whole) * nl[:, i].astype(np.float64) * t_tri
                        + (ar * inv_whole) * nr[:, i].astype(np.float64) * t_tri)
                upd = cost < best_cost[:, a]
                best_cost[upd, a] = cost[upd]
                best_pos[upd, a] = plane[upd]
        axis = np.zeros(S, np.int64)
        c = best_cost[:, 0].copy()
        m1 = c > best_cost[:, 1]
        axis[m1], c[m1] = 1, best_cost[m1, 1]
        m2 = c > best_cost[:, 2]
        axis[m2], c[m2] = 2, best_cost[m2, 2]
        thresh = best_pos[np.arange(S), axis] * 3.0
        # partition (std::partition, bvh_accel.cc:402): mid = l + #left; the k-th misplaced element from the left swaps
        # with the k-th misplaced element from the right -- ranks from exclusive prefix sums over the flags
        pred = cs[axis[sid], tri] < thresh[sid]
        pred[leaf[sid]] = True                                     # leaves: nothing moves
        nleft = np.add.reduceat(pred.astype(np.int64), starts)
        mid = sl + nleft
        left_side = pos < mid[sid]
        mis_l, mis_r = left_side & ~pred, ~left_side & pred
        rank_l = (np.cumsum(mis_l) - mis_l)
        rank_r = (np.cumsum(mis_r) - mis_r)
        m = np.add.reduceat(mis_l.astype(np.int64), starts)
        assert np.array_equal(m, np.add.reduceat(mis_r.astype(np.int64), starts))
        rlist = pos[mis_r]                                         # right-misplaced positions, in position order
        lpos = pos[mis_l]
        ls = sid[mis_l]
        k = rank_l[mis_l] - rank_l[starts][ls]
        partner = rlist[rank_r[starts][ls] + (m[ls] - 1 - k)]
        a_, b_ = idx[lpos].copy(), idx[partner].copy()
        idx[lpos], idx[partner] = b_, a_
        fallback = (~leaf) & ((nleft == 0) | (nleft == n))         # object median, array left as partitioned (bvh_accel.cc:405-418)
        mid = np.where(fallback, sl + (n >> 1), mid)
        nxt = []
        for s in range(S):
            nd = nodes[sn[s]]
            nd["bmin"], nd["bmax"] = bmin[s], bmax[s]
            if leaf[s]:
                nd["flag"], nd["axis"], nd["d"] = 1, 0, (int(n[s]), int(sl[s]))
            else:
                nd["flag"], nd["axis"] = 0, int(axis[s])
                c0 = len(nodes)
                nodes.append(dict())
                nodes.append(dict())
                nd["c"] = (c0, c0 + 1)
                nxt += [(int(sl[s]), int(mid[s]), c0), (int(mid[s]), int(sr[s]), c0 + 1)]
        segs = nxt
        level += 1
    # pre-order numbers: subtree sizes bottom-up (children are created after their parent), numbers top-down
    size = np.ones(len(nodes), np.int64)
    for i in range(len(nodes) - 1, -1, -1):
        if nodes[i]["flag"] == 0:
            size[i] = 1 + size[nodes[i]["c"][0]] + size[nodes[i]["c"][1]]
    num = np.zeros(len(nodes), np.int64)
    for i, nd in enumerate(nodes):
        if nd["flag"] == 0:
            c0, c1 = nd["c"]
            num[c0], num[c1] = num[i] + 1, num[i] + 1 + size[c0]
    out = np.zeros(len(nodes), NODE_DTYPE)
    for i, nd in enumerate(nodes):
        o = out[num[i]]
        o["bmin"], o["bmax"], o["flag"], o["axis"] = nd["bmin"], nd["bmax"], nd["flag"], nd["axis"]
        o["data"] = nd["d"] if nd["flag"] == 1 else (num[nd["c"][0]], num[nd["c"][1]])
    return out, idx
