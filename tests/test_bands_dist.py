"""The N > 1 host path on CPU: world_size-2 (and 3) gloo groups, every rank contributes the rows
the band rule gives it, rank 0 must reassemble the oracle's frame exactly.  The per-rank "render" is
the oracle here (test infrastructure); on GPUs the same BandGather carries rt_render_bands output."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, scene_path


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, band_h, out_path):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle_py
    import rtb200
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = rtb200.load_scene(scene_path("medium"))
    frame = oracle_py.render(scene, W, H, 3, nthreads=1)["rgb"]
    g = rtb200.BandGather(W, H, band_h, rank, world, torch.device("cpu"), dist)
    rows = g.my_rows
    assert len(rows) == rtb200.band_rows(H, band_h, rank, world)
    g.part[: len(rows) * W * 3] = torch.from_numpy(np.ascontiguousarray(frame[rows]).reshape(-1))
    full = g.gather()
    if rank == 0:
        np.save(out_path, full.numpy())
        np.save(out_path + ".ref.npy", frame)
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,W,H,band_h", [(2, 64, 45, 16), (2, 33, 70, 4), (3, 40, 50, 8)])
def test_band_gather_reassembles_the_frame(tmp_path, world, W, H, band_h):
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(world, _free_port(), W, H, band_h, out), nprocs=world, join=True)
    full = np.load(out)
    ref = np.load(out + ".ref.npy")
    assert full.shape == (H, W, 3) and np.array_equal(full, ref)


def test_single_rank_is_a_reshape(rt):
    g = rt.BandGather(8, 5, 16, 0, 1, torch.device("cpu"))
    g.part[: 8 * 5 * 3] = torch.arange(120, dtype=torch.uint8)
    assert torch.equal(g.gather().reshape(-1), torch.arange(120, dtype=torch.uint8))
