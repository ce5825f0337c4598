"""Device time of every rank's share of an N-way split, for several band heights (one GPU renders each share in turn):
how evenly do interleaved bands spread the frame?   usage: probe_ranks.py [N] [band heights, comma separated]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rtb200
sc = rtb200.load_scene(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", "complex.txt"))
W, H, D = 1920, 1080, 5
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
bands = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [16, 8, 4]
r = rtb200.Renderer(0)
r.upload(sc)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
buf = torch.empty(H * W * 3 + 16, dtype=torch.uint8, device="cuda:0")
for bh in bands:
    res = []
    for rank in range(n):
        for _ in range(6): r.render_bands_device(W, H, D, bh, rank, n, buf.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        ms = []
        for k in range(25):
            flush.fill_(k & 255)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); r.render_bands_device(W, H, D, bh, rank, n, buf.data_ptr(), stream.cuda_stream); e1.record(stream)
            torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
        res.append(float(np.median(ms)))
    print("N=%d band_h=%2d: per-rank ms %s  max %.4f mean %.4f (max/mean %.3f)" % (n, bh, " ".join("%.4f" % v for v in res), max(res), np.mean(res), max(res) / np.mean(res)), flush=True)
