"""A/B of the whole-frame kernel against the wavefront kernels: production launch sequence (no stats), CUDA events per
frame on the launching stream, L2 flushed between frames; full frame and one rank's share of an n-way band split.
usage: probe_frame.py [scene] [W H D]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtb200
name = sys.argv[1] if len(sys.argv) > 1 else "complex"
W, H, D = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (1920, 1080, 5)
sc = rtb200.load_scene(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", name + ".txt"))
r = rtb200.Renderer(0)
r.upload(sc)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
for n in (1, 2, 4, 8):
    rows = rtb200.band_rows(H, 16, 0, n)
    buf = torch.empty(rows * W * 3 + 16, dtype=torch.uint8, device="cuda:0")
    for fk in (0, 1):
        r.set_option("frame_kernel", fk)
        for _ in range(5):
            r.render_bands_device(W, H, D, 16, 0, n, buf.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        ms = []
        for k in range(40):
            flush.fill_(k & 255)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            r.render_bands_device(W, H, D, 16, 0, n, buf.data_ptr(), stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        print("%s %dx%d d%d rank 0 of %d, frame_kernel=%d: median %.4f ms  min %.4f ms" % (name, W, H, D, n, fk, ms[len(ms) // 2], ms[0]), flush=True)
