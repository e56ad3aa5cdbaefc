#!/usr/bin/env python
"""Randomised multi-process frames (development aid, N GPUs of one box): every rank renders random frames (sizes, band
rows, sample counts, shader) through mb200_render_frame_gathered -- rows that are multiples of 16 bytes go through the
peer-memory exchange kernel, the others through NCCL; the frame size changes almost every time, so the peer mapping is
torn down and rebuilt constantly -- and compares what it receives with the same frame rendered on its own GPU alone.

    python tools/fuzz_gather.py <world> [seconds] [seed]      (spawns one process per GPU)"""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rank_main(rank, world, idfile, budget, seed):
    import mallie_b200 as M
    from tests import common as T
    m = T.load_mesh("sphere40")
    sc = M.Scene(m["vertices"], m["faces"], device=rank)
    if rank == 0:
        uid = M.Comm.unique_id()
        with open(idfile + ".tmp", "wb") as fp:
            fp.write(uid)
        os.rename(idfile + ".tmp", idfile)
    else:
        t0 = time.time()
        while not os.path.exists(idfile):
            assert time.time() - t0 < 120
            time.sleep(0.05)
        uid = open(idfile, "rb").read()
    comm = M.Comm(sc, world, rank, uid)
    rng = np.random.default_rng(seed)             # the same sequence on every rank
    n_iter = max(1, int(budget * 12))             # a fixed count: every rank must make the same number of collective calls
    paths = {0: 0, 1: 0}
    for it in range(n_iter):
        W = int(rng.integers(5, 120)) * int(rng.choice([4, 4, 1])) + int(rng.integers(0, 2)) * int(rng.choice([0, 1]))
        H = int(rng.integers(3, 260))
        spp = int(rng.integers(1, 5))
        band_rows = int(rng.choice([4, 8, 12, 32]))
        shader = int(rng.choice([M.SHADER_PRIMARY_SHADOW, M.SHADER_PATHTRACE]))
        fg = M.camera_frame((0.2, 0.1, 3.0), (0, 0, 0), width=W, height=H)
        p = sc.render_params(fg, W, H, shader=shader, light=(2, 4, 3), pass_index=int(rng.integers(0, 9)), max_path_length=4)
        want, wcnt, _ = sc.render_frame(p, spp)
        img = np.full((H, W, 3), -1.0, np.float32)
        cnt = np.zeros((H, W), np.int32)
        reps = int(rng.integers(1, 4))            # back-to-back frames of one size: the double buffering
        for _ in range(reps):
            comm.render_frame(p, spp, band_rows, img, cnt)
            assert img.tobytes() == want.tobytes() and np.array_equal(cnt, wcnt), (rank, it, W, H, spp, band_rows, shader)
        paths[comm.exchange_path()] += 1
    print(f"FUZZ GATHER OK rank {rank}: {n_iter} frame sizes, {paths[1]} through the peer-memory kernel, {paths[0]} through NCCL", flush=True)
    comm.close()
    sc.close()


if __name__ == "__main__":
    if len(sys.argv) >= 2 and sys.argv[1] == "--rank":
        rank_main(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], float(sys.argv[5]), int(sys.argv[6]))
        sys.exit(0)
    world = int(sys.argv[1])
    budget = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    idfile = f"/tmp/mb200_fuzz_gather_{os.getpid()}"
    procs = [subprocess.Popen([sys.executable, __file__, "--rank", str(r), str(world), idfile, str(budget), str(seed)]) for r in range(world)]
    rc = [p.wait() for p in procs]
    if os.path.exists(idfile):
        os.remove(idfile)
    sys.exit(max(rc))
