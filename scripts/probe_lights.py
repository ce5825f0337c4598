"""Marginal cost of each phase of the level-0 kernel: complex.txt with 0..5 lights, depth 1."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rtb200
sc = rtb200.load_scene(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", "complex.txt"))
r = rtb200.Renderer(0)
for depth in (1, 5):
    for L in range(0, 6):
        s2 = rtb200.Scene(sc.spheres, sc.lights[:L], sc.ambient, sc.camera)
        r.upload(s2)
        ms = []
        for _ in range(8):
            _, st = r.render(1920, 1080, depth)
            ms.append((st.ms_level0, st.ms_device))
        ms = sorted(ms)[len(ms) // 2]
        print("depth %d L=%d level0 %.3f ms frame %.3f ms rays %d fp64 %d" % (depth, L, ms[0], ms[1], st.closest_queries + st.shadow_queries, st.fp64_intersections))
