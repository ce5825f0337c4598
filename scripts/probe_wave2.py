"""Frame time vs wave_levels WITHOUT stats (so that PDL stays on): CUDA events around rt_render_bands."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtb200
name = sys.argv[1] if len(sys.argv) > 1 else "complex"
W, H, D = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (1920, 1080, 5)
nr = int(sys.argv[5]) if len(sys.argv) > 5 else 1
sc = rtb200.load_scene(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", name + ".txt"))
r = rtb200.Renderer(0)
r.upload(sc)
st = torch.cuda.Stream()
buf = torch.empty(rtb200.band_rows(H, 16, 0, nr) * W * 3 + 16, dtype=torch.uint8, device="cuda:0")
for wl in (1, 2, 3, 4, 5):
    if wl > max(D, 1): break
    r.set_option("wave_levels", wl)
    for _ in range(5): r.render_bands_device(W, H, D, 16, 0, nr, buf.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    ts = []
    for _ in range(40):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); r.render_bands_device(W, H, D, 16, 0, nr, buf.data_ptr(), st.cuda_stream); b.record(st)
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    print("%s %dx%d d%d rank 0 of %d, wave_levels %d: median %.3f ms min %.3f" % (name, W, H, D, nr, wl, ts[len(ts)//2], ts[0]))
