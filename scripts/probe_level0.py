"""Level-0 kernel times of a stats render (event marks after k_closest0 / k_shadow / k_shade; no PDL): medians over frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rtb200
sc = rtb200.load_scene(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", "complex.txt"))
W, H = 1920, 1080
D = int(sys.argv[1]) if len(sys.argv) > 1 else 1
r = rtb200.Renderer(0)
r.upload(sc)
r.set_option("level_timing", 1)               # event marks between the level-0 kernels (turns PDL off)
buf = torch.empty(H * W * 3 + 16, dtype=torch.uint8, device="cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
v = []
for k in range(25):
    flush.fill_(k & 255); torch.cuda.synchronize()
    st = r.render_bands_device(W, H, D, 16, 0, 1, buf.data_ptr(), None, want_stats=True)
    if k >= 5: v.append((st.ms_closest0, st.ms_shadow0, st.ms_level0 - st.ms_closest0 - st.ms_shadow0, st.ms_device))
m = np.median(np.array(v), axis=0)
print("depth %d: closest0 %.4f shadow0 %.4f shade0 %.4f frame %.4f ms" % (D, *m))
