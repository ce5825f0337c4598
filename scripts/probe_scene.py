"""Times one configuration through the C ABI (device-resident scene): per-phase CUDA-event times.
usage: probe_scene.py {simple|medium|complex|synth:N:seed[:rmin:rmax]} W H depth [reps] [check]
`check` additionally compares hit indices / shadow masks with the exact FP64 kernel (mode 1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import rtb200

name, W, H, D = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
check = len(sys.argv) > 6 and sys.argv[6] == "check"
if name.startswith("synth"):
    import gen_scene
    p = name.split(":")
    extra = [float(x) for x in p[3:5]]
    sc = rtb200.Scene(*gen_scene.generate(int(p[1]), int(p[2]), *extra))
else:
    sc = rtb200.load_scene(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", name + ".txt"))
r = rtb200.Renderer(0, accel=int(os.environ["RT_ACCEL"]) if "RT_ACCEL" in os.environ else None)
r.upload(sc)
ms = []
for _ in range(reps):
    rgb, st = r.render(W, H, D)
    ms.append((st.ms_device, st.ms_closest0, st.ms_shadow0, st.ms_level0))
ms = sorted(ms)[len(ms) // 2]
rays = st.closest_queries + st.shadow_queries
print("%s %dx%d d%d: frame %.3f ms (closest0 %.3f shadow0 %.3f level0 %.3f)  rays %d  %.1f Mrays/s  fp64 %d viol %d launches %d alive %s cand/walk %.1f (%d walks, %d fallbacks)" % (
    name, W, H, D, ms[0], ms[1], ms[2], ms[3], rays, rays / ms[0] / 1e3, st.fp64_intersections, st.filter_violations, st.kernel_launches,
    [int(x) for x in st.alive[:D]], st.bundle_candidates / max(1, st.bundle_walks), st.bundle_walks, st.bundle_fallbacks))
if check:
    rgb, hit, mask, st = r.render_debug(W, H, D)
    with rtb200.Renderer(0, mode="exact") as e:
        e.upload(sc)
        rgb2, hit2, mask2, st2 = e.render_debug(W, H, D)
    ok, pct, mx = rtb200.compare_rgb(rgb2, rgb, 0.5)
    print("check vs exact FP64 kernel: hit %s mask %s rgb ok=%s pct=%.5f max=%d  exact %.1f ms" % (
        np.array_equal(hit, hit2), np.array_equal(mask, mask2), ok, pct, mx, st2.ms_device))
