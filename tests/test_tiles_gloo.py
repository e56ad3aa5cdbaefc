"""Multi-process image partition on CPU (gloo, world_size 2 and 3): every rank "renders" its interleaved
row bands into a compact buffer, ONE all-gather re-assembles the framebuffer (mallie_b200/tiles.py).
The pixel values come from the oracle so the assembled image is checked against a real single-process frame.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, BAND = 96, 70, 8          # 70 rows: the last band is ragged and ranks own different row counts


def full_frame():
    sys.path.insert(0, ROOT)
    from oracle import orabind as O
    from tests import common as T
    om, ob = T.oracle_scene("sphere40")
    fr = O.camera_frame((0.3, 0.2, 3), (0, 0, 0), width=W, height=H)
    img, _, _ = ob.render_pass(fr, W, H, rng_mode=1, pass_index=0, shader=1, light=(2.0, 4.0, 3.0))
    return img


def worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    from mallie_b200 import tiles
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.from_numpy(full_frame())
        rows = tiles.band_rows_of_rank(H, BAND, world, rank)
        local = full[torch.from_numpy(rows)].contiguous()       # what mb200_render_frame(band_compact=1) returns
        g = tiles.FramebufferGather(W, H, BAND, world, rank, torch.device("cpu"))
        out = g(local)
        ok = bool(torch.equal(out, full))
        again = g(local)                                         # buffers are reused frame after frame
        ok = ok and bool(torch.equal(again, full))
        # whole-job ray count reduction as bench.py does it
        t = torch.tensor([len(rows) * W], dtype=torch.int64)
        dist.all_reduce(t)
        ok = ok and int(t[0]) == W * H
        q.put((rank, ok, len(rows)))
    finally:
        dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_band_gather_reassembles_frame(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert sum(n for _, _, n in res) == H
