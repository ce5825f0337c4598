"""Row-band sharding of one frame across the GPUs of a box -- host-side plumbing.

Rank r of n owns the bands b with b % n == r (band = `band_h` consecutive rows, interleaving
balances sky rows against sphere-dense rows) and renders them compactly with
``rt_render_bands``.  The only exchange step of the path is the gather of the 8-bit bands to
rank 0, done with ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests);
rank 0 then scatters the rows of every rank back to their image positions.

PeerFrame is the B200-native form of that step: rank 0 owns the assembled frame, every other rank maps it
through CUDA IPC and its render kernels store their finished rows straight into it over NVLink
(rt_render_bands_frame) -- the transfer overlaps the compute pixel by pixel, and what is left of the gather is
one completion flag per rank (rt_peer_signal / rt_peer_wait) plus an acknowledge word for back-pressure.
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


class _DevArray:
    """Exposes a raw device allocation to torch (zero-copy) through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerFrame:
    """One assembled [H, W, 3] frame in rank 0's HBM that all ranks of the box render into.

    Per frame:  rank r > 0: wait(ack >= seq-1) -> render rows into rank 0's frame -> signal(flag[r] = seq)
                rank 0    : render its own rows -> wait(flag[1..n) >= seq) -> (consumer) -> release(): ack = seq
    Handles travel once through `dist` (any backend; object broadcast).  All waits / signals are stream ordered."""

    def __init__(self, renderer, W, H, band_h, rank, nranks, dist=None):
        self.r, self.W, self.H, self.band_h, self.rank, self.n, self.dist = renderer, W, H, band_h, rank, nranks, dist
        self.seq = 0
        self.nbytes = H * W * 3 + 16
        self._owned = []
        if rank == 0:
            self.frame_ptr = renderer.dev_alloc(self.nbytes)
            self.ctl_ptr = renderer.dev_alloc(4 * 128)          # [0..64) flags, [64] ack, [65] error word
            self._owned = [self.frame_ptr, self.ctl_ptr]
            handles = [renderer.ipc_export(self.frame_ptr), renderer.ipc_export(self.ctl_ptr)] if nranks > 1 else None
        else:
            handles = None
        if nranks > 1:
            box = [handles]
            dist.broadcast_object_list(box, src=0)
            if rank != 0:
                self.frame_ptr = renderer.ipc_open(box[0][0])
                self.ctl_ptr = renderer.ipc_open(box[0][1])
                self.err_ptr = renderer.dev_alloc(4)            # local error word of this rank's waits
                self._owned = [self.err_ptr]
        if rank == 0:
            self.err_ptr = self.ctl_ptr + 4 * 65

    def render(self, depth, stream_ptr=None, want_stats=False):
        """Enqueues this rank's share of the next frame (and, on rank 0, the wait for everybody else's)."""
        self.seq += 1
        r = self.r
        if self.rank != 0:
            r.peer_wait(self.ctl_ptr + 4 * 64, 1, self.seq - 1, self.err_ptr, stream_ptr)     # frame seq-1 consumed?
        st = r.render_bands_frame(self.W, self.H, depth, self.band_h, self.rank, self.n, self.frame_ptr, stream_ptr, want_stats)
        if self.n > 1:
            if self.rank != 0:
                r.peer_signal(self.ctl_ptr + 4 * self.rank, self.seq, stream_ptr)
            else:
                r.peer_wait(self.ctl_ptr + 4, self.n - 1, self.seq, self.err_ptr, stream_ptr)
        return st

    def release(self, stream_ptr=None):
        """Rank 0, after whatever consumes the frame has been enqueued: lets the other ranks start the next frame."""
        if self.rank == 0 and self.n > 1:
            self.r.peer_signal(self.ctl_ptr + 4 * 64, self.seq, stream_ptr)

    def frame(self):
        """Rank 0: the assembled frame as a torch uint8 [H, W, 3] view of the device memory."""
        import torch
        assert self.rank == 0
        return torch.as_tensor(_DevArray(self.frame_ptr, (self.H, self.W, 3), "|u1"), device="cuda")

    def error(self):
        """After a synchronise: 0, or 1 + index of the flag a wait timed out on."""
        import torch
        return int(torch.as_tensor(_DevArray(self.err_ptr, (1,), "<u4"), device="cuda").item())

    def close(self):
        if self.rank != 0 and self.n > 1:
            self.r.ipc_close(self.frame_ptr); self.r.ipc_close(self.ctl_ptr)
        for p in self._owned:
            self.r.dev_free(p)
        self._owned = []
