"""Re-runs ONE scene of scripts/fuzz_parity.py and prints what differs (pixels, levels, values).  usage: repro_fuzz.py seed [big|mid]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "scripts")):
    sys.path.insert(0, p)
import numpy as np
import rtb200, oracle_py
from test_gpu_parity import _random_scene
seed = int(sys.argv[1]); MID = len(sys.argv) > 2 and sys.argv[2] == "mid"; BIG = len(sys.argv) > 2 and sys.argv[2] in ("big", "mid")
g = np.random.default_rng(seed)
if BIG:
    import gen_scene
    sc = rtb200.Scene(*gen_scene.generate(int(g.integers(60, 300) if MID else g.integers(300, 4000)), seed, 0.05, float(g.uniform(0.3, 1.5))))
    W, H, D = int(g.integers(16, 120)), int(g.integers(16, 70)), int(g.integers(1, 7))
else:
    sc = _random_scene(rtb200, seed)
    if seed % 2: W, H, D = [(96, 54, 4), (61, 47, 6), (130, 40, 3), (40, 90, 8)][seed % 4]
    else: W, H, D = int(g.integers(2, 260)), int(g.integers(2, 200)), int(g.integers(0, 9))
    if seed % 3 == 0:
        k = float(10.0 ** g.uniform(-2, 3)); off = g.uniform(-1, 1, 3) * float(10.0 ** g.uniform(0, 4)) * k
        sp = sc.spheres.copy(); sp[:, :3] = sp[:, :3] * k + off; sp[:, 3] *= k
        li = sc.lights.copy(); li[:, :3] = li[:, :3] * k + off
        cm = sc.camera.copy(); cm[:3] = cm[:3] * k + off; cm[3:6] = cm[3:6] * k + off
        sc = rtb200.Scene(sp, li, sc.ambient, cm)
    if seed % 7 == 3:
        nl = int(g.integers(6, 41))
        li = np.column_stack([g.uniform(-8, 8, (nl, 3)) - np.array([0, -4, 6]), g.uniform(0.05, 0.3, (nl, 3)), np.ones(nl)])
        sc = rtb200.Scene(sc.spheres, li, sc.ambient, sc.camera)
    if seed % 37 == 0: sc = rtb200.Scene(sc.spheres, sc.lights[:0], sc.ambient, sc.camera)
    if seed % 41 == 0: sc = rtb200.Scene(sc.spheres[:0], sc.lights, sc.ambient, sc.camera)
print("seed %d: N %d L %d %dx%d d%d lib %s" % (seed, sc.nspheres, sc.nlights, W, H, D, os.environ.get("RTB200_LIB", "default")))
o = oracle_py.render(sc, W, H, D, want_idx=True)
for m, kw in (("fast", {"accel": 1} if BIG else {}), ("bvh", {"mode": "bvh"}), ("exact", {"mode": "exact"})):
    with rtb200.Renderer(0, **kw) as r:
        r.upload(sc)
        frames = [r.render(W, H, D)[0] for _ in range(3)]
        rgb, hit, mask, st = r.render_debug(W, H, D)
        for k, f in enumerate(frames):
            dif = np.argwhere((f != rgb).any(axis=2))
            if len(dif):
                j, i = dif[0]
                print("  %s: render #%d differs from render_debug at %d pixels, first (row %d col %d): %s vs %s; oracle %s, hit levels %s" % (
                    m, k, len(dif), j, i, f[j, i].tolist(), rgb[j, i].tolist(), o["rgb"][j, i].tolist(), hit[j, i].tolist() if D else []))
        bh = np.argwhere(hit != o["hit_idx"]) if D else []
        bm = np.argwhere(mask != o["shadow_mask"]) if D else []
        ok, pct, mx = rtb200.compare_rgb(o["rgb"], rgb, 0.5)
        print("  %s: hit mismatches %d, mask mismatches %d, rgb ok %s max %d, viol %d, closest %d (oracle %d)" % (
            m, len(bh), len(bm), ok, mx, st.filter_violations, st.closest_queries, o["counters"]["closest_queries"]))
        for j, i, k in list(bh)[:4]:
            print("     hit (row %d col %d level %d): got %s want %s" % (j, i, k, hit[j, i].tolist(), o["hit_idx"][j, i].tolist()))
