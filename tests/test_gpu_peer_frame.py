"""Frame assembly over peer memory (rt_render_bands_frame + CUDA IPC + completion flags): two PROCESSES, one rank
each, render their interleaved bands into rank 0's frame.  Both ranks use cuda:0 here (the round-end GPU tier has
one GPU; IPC between processes works on one device too) -- on a multi-GPU box the same code runs one rank per GPU
over NVLink (bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, scene_path
from test_bands_dist import _free_port

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, W, H, D, band_h, frames, out_path):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import rtb200
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    scene = rtb200.load_scene(scene_path("complex"))
    r = rtb200.Renderer(0)
    r.upload(scene)
    pf = rtb200.PeerFrame(r, W, H, band_h, rank, world, dist)
    outs = []
    for _ in range(frames):
        pf.render(D)
        if rank == 0:
            torch.cuda.synchronize()
            outs.append(pf.frame().cpu().numpy().copy())
            pf.release()
    torch.cuda.synchronize()
    assert pf.error() == 0
    if rank == 0:
        ref, _ = r.render(W, H, D)
        np.save(out_path, np.stack(outs + [ref]))
    dist.barrier()
    pf.close()
    r.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,W,H,D,band_h", [(2, 320, 180, 5, 16), (3, 200, 100, 3, 8)])
def test_ranks_assemble_the_frame_in_rank0_memory(tmp_path, world, W, H, D, band_h):
    out = str(tmp_path / "frames.npy")
    mp.spawn(_worker, args=(world, _free_port(), W, H, D, band_h, 3, out), nprocs=world, join=True)
    a = np.load(out)
    for k in range(3):
        assert np.array_equal(a[k], a[-1]), k


def test_single_rank_frame_mode_equals_rt_render(rt, scenes):
    with rt.Renderer(0) as r:
        r.upload(scenes["medium"])
        for (W, H, D, band_h) in [(333, 77, 4, 16), (64, 64, 0, 16)]:
            pf = rt.PeerFrame(r, W, H, band_h, 0, 1)
            pf.render(D)
            torch.cuda.synchronize()
            got = pf.frame().cpu().numpy()
            ref, _ = r.render(W, H, D)
            assert np.array_equal(got, ref)
            pf.close()
