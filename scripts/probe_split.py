"""Does splitting ONE GPU's frame over k contexts on k streams (interleaved bands, each with 1/k of the resident grid) hide
the latency-bound tails?  Wall clock per frame into a pinned host frame, single context vs rt_create_multi(k) on one device.
usage: RT_GRID_DIV=k probe_split.py k"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rtb200
k = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sc = rtb200.load_scene(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", "complex.txt"))
W, H, D = 1920, 1080, 5
r = rtb200.Renderer(0)
r.upload(sc)
frame = r.pinned_frame(W, H)
def timeit(fn, n=60):
    for _ in range(10): fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e3
one = timeit(lambda: r.render(W, H, D, out=frame, want_stats=False))
ref = frame.copy()
_, st = r.render(W, H, D, out=frame)
print("single context: %.4f ms wall per frame (device %.4f)" % (one, st.ms_device))
for bh in (16, 64):
    with rtb200.MultiRenderer(k) as m:
        m.upload(sc)
        t = timeit(lambda: m.render(W, H, D, band_h=bh, out=frame, want_stats=False))
        _, st = m.render(W, H, D, band_h=bh, out=frame)
        print("%d contexts, band %d: %.4f ms wall per frame (slowest rank's device time %.4f), identical=%s" % (k, bh, t, st.ms_device, np.array_equal(ref, frame)))
