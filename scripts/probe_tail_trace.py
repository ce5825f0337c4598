"""Phase stamps of the tail kernel (diagnostics build: scripts/build_variant.sh trace -DRT_TAIL_TRACE, RTB200_LIB=...):
per reflection level handled by k_bounce, over the warps that took a chunk: time of closest / exact finish / geometry /
shadow walks / continuation in microseconds (SM clock), live rays and hits per warp, and the wall-clock span of the level.
usage: probe_tail_trace.py [n_ranks]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rtb200
from rtb200 import api
sc = rtb200.load_scene(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", "complex.txt"))
W, H, D = 1920, 1080, 5
r = rtb200.Renderer(0)
r.upload(sc)
lib = r._lib
WARPS, LEVELS, STAMPS = 4096, 6, 12
buf = torch.empty(H * W * 3 + 16, dtype=torch.uint8, device="cuda:0")
for n in (int(a) for a in (sys.argv[1:] or ["1", "8"])):
    for _ in range(6): r.render_bands_device(W, H, D, 16, 0, n, buf.data_ptr(), None)
    torch.cuda.synchronize()
    out = np.zeros(WARPS * LEVELS * STAMPS, dtype=np.uint64)
    # clear, render once, read
    got = lib.rt_debug_tail_trace(out.ctypes.data_as(C.POINTER(C.c_uint64)), out.size)
    assert got == out.size
    t = out.reshape(WARPS, LEVELS, STAMPS)
    clk = (t[:, :, :6] & np.uint64((1 << 48) - 1)).astype(np.int64)
    ext = (t[:, :, :6] >> np.uint64(48)).astype(np.int64)
    g0, g1 = t[:, :, 6].astype(np.int64), t[:, :, 7].astype(np.int64)
    print("== 1/%d of the frame" % n)
    base = g0[g0 > g0.max() - 400000].min()         # (stamps of earlier frames are older)
    for lv in range(LEVELS):
        m = (g0[:, lv] >= base) & (g1[:, lv] > 0) & (clk[:, lv, 0] > 0) & (clk[:, lv, 5] > clk[:, lv, 0])
        if not m.any(): continue
        d = np.diff(clk[m, lv, :], axis=1) / 1965.0
        names = ["closest", "finish", "geom", "shadow", "cont"]
        print(" tail level +%d: %d warps, live rays/warp mean %.1f max %d, hits/warp %.1f, continuing %.1f; span %.1f .. %.1f us after the first stamp" % (
            lv, m.sum(), ext[m, lv, 0].mean(), ext[m, lv, 0].max(), ext[m, lv, 2].mean(), ext[m, lv, 5].mean(),
            (g0[m, lv].min() - base) / 1e3, (g1[m, lv].max() - base) / 1e3))
        for k, nm in enumerate(names):
            print("   %-8s mean %7.2f  p50 %7.2f  p90 %7.2f  max %7.2f us" % (nm, d[:, k].mean(), np.percentile(d[:, k], 50), np.percentile(d[:, k], 90), d[:, k].max()))
        pr = t[m, lv, 8:12].astype(np.float64)
        print("   packed shadow walks (warps that used them: %d): chunk tests %.2f us, drains %.2f us, drain iterations %.1f, chunks %.1f" % (
            (pr[:, 3] > 0).sum(), pr[:, 0].mean() / 1965, pr[:, 1].mean() / 1965, pr[:, 2].mean(), pr[:, 3].mean()))
        tot = (clk[m, lv, 5] - clk[m, lv, 0]) / 1965.0
        print("   %-8s mean %7.2f  p50 %7.2f  p90 %7.2f  max %7.2f us" % ("level", tot.mean(), np.percentile(tot, 50), np.percentile(tot, 90), tot.max()))
