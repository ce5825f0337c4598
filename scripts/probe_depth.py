"""Marginal cost of each reflection level: complex.txt 1080p at depth 1..6."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
sc = rtb200.load_scene(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", "complex.txt"))
r = rtb200.Renderer(0)
r.upload(sc)
for depth in (1, 2, 3, 4, 5, 6, 10):
    ms = []
    for _ in range(10):
        _, st = r.render(1920, 1080, depth)
        ms.append(st.ms_device)
    ms = sorted(ms)[len(ms) // 2]
    print("depth %d frame %.3f ms rays %d alive %s launches %d" % (depth, ms, st.closest_queries + st.shadow_queries, [int(x) for x in st.alive[:depth]], st.kernel_launches))
