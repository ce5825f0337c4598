"""Per-rank latency floor: time ONE rank's share of the frame (rank 0 of n) on one GPU, with phase times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtb200
sc = rtb200.load_scene(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", "complex.txt"))
W, H, D = 1920, 1080, 5
r = rtb200.Renderer(0)
r.upload(sc)
for n in (1, 2, 4, 8):
    rows = rtb200.band_rows(H, 16, 0, n)
    buf = torch.empty(rows * W * 3 + 16, dtype=torch.uint8, device="cuda:0")
    for wl in (1, 2):
        r.set_option("wave_levels", wl)
        ms = []
        for _ in range(30):
            st = r.render_bands_device(W, H, D, 16, 0, n, buf.data_ptr(), None, want_stats=True)
            ms.append((st.ms_device, st.ms_closest0, st.ms_shadow0, st.ms_level0))
        ms.sort(); m = ms[len(ms) // 2]
        print("n=%d wave_levels=%d: rank-0 share %.3f ms (closest0 %.3f shadow0 %.3f level0 %.3f) alive %s" % (n, wl, m[0], m[1], m[2], m[3], [int(x) for x in st.alive[:D]]))
