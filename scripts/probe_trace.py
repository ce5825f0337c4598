"""Phase trace of the whole-frame kernel (RT_FRAME_TRACE=1): usage probe_trace.py [scene W H D n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtb200
name = sys.argv[1] if len(sys.argv) > 1 else "complex"
W, H, D = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (1920, 1080, 5)
n = int(sys.argv[5]) if len(sys.argv) > 5 else 1
sc = rtb200.load_scene(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", name + ".txt"))
r = rtb200.Renderer(0)
r.upload(sc)
rows = rtb200.band_rows(H, 16, 0, n)
buf = torch.empty(rows * W * 3 + 16, dtype=torch.uint8, device="cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
for k in range(6):
    flush.fill_(k)
    torch.cuda.synchronize()
    st = r.render_bands_device(W, H, D, 16, 0, n, buf.data_ptr(), None, want_stats=False)
torch.cuda.synchronize()
st = r.render_bands_device(W, H, D, 16, 0, n, buf.data_ptr(), None, want_stats=True)
print("ms_device %.4f alive %s hits %d" % (st.ms_device, [int(x) for x in st.alive[:D]], st.hits))
