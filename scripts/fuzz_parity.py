"""Long-running parity fuzz (not part of the test suite): random adversarial scenes vs the oracle, all kernel families.
usage: fuzz_parity.py [first_seed] [count]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "scripts")):
    sys.path.insert(0, p)
import numpy as np
import rtb200, oracle_py
from test_gpu_parity import _random_scene
first, count = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (100, 200)
rs = {"fast": rtb200.Renderer(0), "bvh": rtb200.Renderer(0, mode="bvh")}
bad = 0
for seed in range(first, first + count):
    sc = _random_scene(rtb200, seed)
    W, H, D = [(96, 54, 4), (61, 47, 6), (130, 40, 3), (40, 90, 8)][seed % 4]
    o = oracle_py.render(sc, W, H, D, want_idx=True)
    for m, r in rs.items():
        r.upload(sc)
        rgb, hit, mask, st = r.render_debug(W, H, D)
        ok_rgb, pct, mx = rtb200.compare_rgb(o["rgb"], rgb, 0.5)
        good = np.array_equal(hit, o["hit_idx"]) and np.array_equal(mask, o["shadow_mask"]) and ok_rgb and mx <= 2 and st.filter_violations == 0 \
            and st.closest_queries == o["counters"]["closest_queries"] and st.occluded == o["counters"]["occluded"]
        if not good:
            bad += 1
            print("MISMATCH seed %d mode %s: hit %s mask %s rgb %s (%.4f%%, max %d) viol %d" % (
                seed, m, np.array_equal(hit, o["hit_idx"]), np.array_equal(mask, o["shadow_mask"]), ok_rgb, pct, mx, st.filter_violations))
print("fuzz: %d scenes x %d modes, %d mismatches" % (count, len(rs), bad))
sys.exit(1 if bad else 0)
