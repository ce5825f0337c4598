"""Frame time against max_depth (production launch sequence, CUDA events, L2 flushed): what each reflection level adds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtb200
sc = rtb200.load_scene(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", "complex.txt"))
W, H = 1920, 1080
r = rtb200.Renderer(0)
r.upload(sc)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
buf = torch.empty(H * W * 3 + 16, dtype=torch.uint8, device="cuda:0")
for n in (1, 8):
    prev = 0
    for D in ([int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else (1, 2, 3, 4, 5, 6, 8)):
        for _ in range(5): r.render_bands_device(W, H, D, 16, 0, n, buf.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        ms = []
        for k in range(40):
            flush.fill_(k & 255)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); r.render_bands_device(W, H, D, 16, 0, n, buf.data_ptr(), stream.cuda_stream); e1.record(stream)
            torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
        ms.sort(); m = ms[len(ms) // 2]
        print("rank 0 of %d, depth %d: %.4f ms (+%.4f)" % (n, D, m, m - prev), flush=True); prev = m
