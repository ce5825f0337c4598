"""Frame time vs the number of wavefront levels (rt_set_option wave_levels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
name = sys.argv[1] if len(sys.argv) > 1 else "complex"
W, H, D = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (1920, 1080, 5)
sc = rtb200.load_scene(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", name + ".txt"))
r = rtb200.Renderer(0)
r.upload(sc)
for wl in (1, 2, 3, 4, 5, 6, 8, 10):
    if wl > max(D, 1): break
    r.set_option("wave_levels", wl)
    ms = []
    for _ in range(20):
        _, st = r.render(W, H, D)
        ms.append(st.ms_device)
    ms = sorted(ms)[len(ms) // 2]
    print("%s %dx%d d%d wave_levels %d: frame %.3f ms launches %d" % (name, W, H, D, wl, ms, st.kernel_launches))
