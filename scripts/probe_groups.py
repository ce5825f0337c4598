"""Experiment: one frame as G interleaved band groups rendered CONCURRENTLY on G streams of one GPU
(G contexts, each with its own queues): do the latency-bound phases of one group hide behind the others?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtb200
sc = rtb200.load_scene(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", "complex.txt"))
W, H, D = 1920, 1080, 5
dev = torch.device("cuda", 0)
for G in (1, 2, 3, 4, 6):
    for band_h in (16, 64, 270):
        rs = [rtb200.Renderer(0) for _ in range(G)]
        for r in rs: r.upload(sc)
        streams = [torch.cuda.Stream(device=dev) for _ in range(G)]
        bufs = [torch.empty(rtb200.band_rows(H, band_h, g, G) * W * 3 + 16, dtype=torch.uint8, device=dev) for g in range(G)]
        main = torch.cuda.Stream(device=dev)
        def frame():
            start = torch.cuda.Event(enable_timing=True); end = torch.cuda.Event(enable_timing=True)
            start.record(main)
            for g in range(G):
                streams[g].wait_event(start)
                rs[g].render_bands_device(W, H, D, band_h, g, G, bufs[g].data_ptr(), streams[g].cuda_stream)
                e = torch.cuda.Event(); e.record(streams[g]); main.wait_event(e)
            end.record(main)
            return start, end
        for _ in range(5): frame()
        torch.cuda.synchronize()
        ts = []
        for _ in range(30):
            s, e = frame(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        ts.sort()
        print("G=%d band_h=%3d: median %.3f ms  min %.3f ms" % (G, band_h, ts[len(ts)//2], ts[0]))
        for r in rs: r.close()
