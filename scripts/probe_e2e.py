"""Where the end-to-end step goes: upload alone, render (no copy), render + D2H."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtb200
sc = rtb200.load_scene(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", "complex.txt"))
W, H, D = 1920, 1080, 5
r = rtb200.Renderer(0)
r.upload(sc)
host = r.pinned_frame(W, H)
buf = torch.empty(H * W * 3 + 16, dtype=torch.uint8, device="cuda:0")
def timeit(f, n=200):
    for _ in range(10): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("upload only            %.4f ms" % timeit(lambda: r.upload(sc)))
print("render to device + sync %.4f ms" % timeit(lambda: (r.render_bands_device(W, H, D, 16, 0, 1, buf.data_ptr(), None), torch.cuda.synchronize())))
print("rt_render (with D2H)    %.4f ms" % timeit(lambda: r.render(W, H, D, out=host, want_stats=False)))
print("upload + rt_render      %.4f ms" % timeit(lambda: (r.upload(sc), r.render(W, H, D, out=host, want_stats=False))))
