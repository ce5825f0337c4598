"""Row-band sharding of one frame across the GPUs of a box -- host-side plumbing.

Rank r of n owns the bands b with b % n == r (band = `band_h` consecutive rows, interleaving
balances sky rows against sphere-dense rows) and renders them compactly with
``rt_render_bands``.  The only exchange step of the path is the gather of the 8-bit bands to
rank 0, done with ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests);
rank 0 then scatters the rows of every rank back to their image positions.
"""
import numpy as np

from .api import band_row_list, band_rows


class BandGather:
    """Pre-allocates the buffers of the gather so that a step is just gather + index_copy."""

    def __init__(self, W, H, band_h, rank, nranks, device, dist=None):
        import torch
        self.torch, self.dist = torch, dist
        self.W, self.H, self.band_h, self.rank, self.n = W, H, band_h, rank, nranks
        self.rows = [band_row_list(H, band_h, k, nranks) for k in range(nranks)]
        self.max_rows = max(len(r) for r in self.rows)
        # +16: rt_render_bands wants a 16-byte aligned buffer it may overrun by nothing, the pad keeps
        # every rank's tensor the same size for the collective
        self.part = torch.zeros(self.max_rows * W * 3 + 16, dtype=torch.uint8, device=device)
        self.gathered = None
        self.full = None
        if rank == 0:
            self.full = torch.zeros((H, W, 3), dtype=torch.uint8, device=device)
            self.row_idx = [torch.from_numpy(r.astype(np.int64)).to(device) for r in self.rows]
            if nranks > 1:
                self.gathered = [torch.zeros_like(self.part) for _ in range(nranks)]

    @property
    def my_rows(self):
        return self.rows[self.rank]

    def gather(self):
        """Collects every rank's bands on rank 0 and returns the assembled [H, W, 3] frame there
        (None on the other ranks).  With one rank it only reshapes."""
        if self.n > 1:
            self.dist.gather(self.part, self.gathered, dst=0)
        if self.rank != 0:
            return None
        parts = self.gathered if self.n > 1 else [self.part]
        for k in range(self.n):
            idx = self.row_idx[k]
            self.full[idx] = parts[k][: idx.numel() * self.W * 3].view(idx.numel(), self.W, 3)
        return self.full
