"""Multi-GPU image partition: pixels are independent, so the frame is split by rows.

Rows are cut into bands of `band_rows` scanlines; band b belongs to rank b % world (interleaving balances the
load: expensive rows are spread over all GPUs).  Every rank renders its bands into a compact buffer
(mb200_render_params.band_compact) and ONE collective per frame re-assembles the framebuffer:
all_gather of the padded per-rank buffers followed by a row permutation.  There is no other communication
on the data path; each rank keeps a full replica of the scene.

Works on any torch.distributed backend: NCCL on the GPUs (bench.py), gloo on CPU tensors (tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def band_rows_of_rank(height, band_rows, world, rank):
    """Global row indices owned by `rank`, in the order they are packed in its compact buffer."""
    nbands = (height + band_rows - 1) // band_rows
    rows = []
    for b in range(rank, nbands, world):
        rows.extend(range(b * band_rows, min((b + 1) * band_rows, height)))
    return np.asarray(rows, dtype=np.int64)


def max_local_rows(height, band_rows, world):
    return max(len(band_rows_of_rank(height, band_rows, world, r)) for r in range(world))


def gather_permutation(height, band_rows, world):
    """perm[y] = row of the [world * max_local_rows] gathered buffer that holds global row y."""
    pad = max_local_rows(height, band_rows, world)
    perm = np.empty(height, dtype=np.int64)
    for r in range(world):
        rows = band_rows_of_rank(height, band_rows, world, r)
        perm[rows] = r * pad + np.arange(len(rows))
    return perm


class FramebufferGather:
    """One all-gather per frame: local [rows_local, W, C] -> full [H, W, C] on every rank."""

    def __init__(self, width, height, band_rows, world, rank, device, channels=3, dtype=torch.float32, group=None):
        self.world, self.rank, self.group = world, rank, group
        self.pad_rows = max_local_rows(height, band_rows, world)
        self.local_rows = len(band_rows_of_rank(height, band_rows, world, rank))
        self.send = torch.zeros((self.pad_rows, width, channels), dtype=dtype, device=device)
        self.recv = torch.zeros((world * self.pad_rows, width, channels), dtype=dtype, device=device)
        self.perm = torch.from_numpy(gather_permutation(height, band_rows, world)).to(device)
        self.full = torch.zeros((height, width, channels), dtype=dtype, device=device)

    def __call__(self, local):
        assert local.shape[0] == self.local_rows
        if self.local_rows == self.pad_rows and local.is_contiguous():
            send = local                                  # the render kernel's output IS the send buffer
        else:
            self.send[: self.local_rows].copy_(local)
            send = self.send
        dist.all_gather_into_tensor(self.recv, send, group=self.group)
        torch.index_select(self.recv, 0, self.perm, out=self.full)
        return self.full
