"""ctypes wrapper of oracle/librt_oracle.so and of the oracle/_ref binaries.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_lib = None


class Counters(C.Structure):
    _fields_ = [("closest_queries", C.c_uint64), ("hits", C.c_uint64), ("shadow_queries", C.c_uint64),
                ("occluded", C.c_uint64), ("alive", C.c_uint64 * 32)]

    def as_dict(self):
        return {"closest_queries": int(self.closest_queries), "hits": int(self.hits),
                "shadow_queries": int(self.shadow_queries), "occluded": int(self.occluded),
                "alive": [int(x) for x in self.alive], "rays": int(self.closest_queries + self.shadow_queries)}


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "librt_oracle.so")
        if not os.path.exists(path):
            env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
            subprocess.check_call(["make", "-C", HERE, "oracle"], env=env, stdout=subprocess.DEVNULL)
        _lib = C.CDLL(path)
        vp, i = C.c_void_p, C.c_int
        _lib.rto_render.argtypes = [vp, i, vp, i, vp, vp, vp, C.c_double, i, i, i, vp, vp, vp, vp,
                                    C.POINTER(Counters), i, i]
        _lib.rto_intersect.argtypes = [vp, vp, vp, C.c_double, C.POINTER(C.c_double)]
        _lib.rto_write_ppm.argtypes = [C.c_char_p, vp, i, i]
        _lib.rto_num_threads.restype = i
        _lib.rto_set_sample_offset.argtypes = [C.c_double, C.c_double]
    return _lib


def num_threads():
    return lib().rto_num_threads()


def render_supersampled(scene, W, H, depth):
    """The reference's `ray_cuda -a` rule (src/main_gpu.cu:249-258,327-333) in the serial renderer's FP64: four
    samples per pixel at offsets (0,0) (.5,0) (0,.5) (.5,.5), summed in that order, x 1/4, then quantised.
    Returns dict(rgb [H,W,3], samples = the four per-sample render() dicts in that order)."""
    passes = [render(scene, W, H, depth, want_fb=True, want_idx=True, sample_offset=o)
              for o in ((0.0, 0.0), (0.5, 0.0), (0.0, 0.5), (0.5, 0.5))]
    acc = ((passes[0]["fb"] + passes[1]["fb"]) + passes[2]["fb"]) + passes[3]["fb"]
    rgb = (255.99 * np.minimum(1.0, acc * 0.25)).astype(np.int32).astype(np.uint8)
    return {"rgb": rgb, "samples": passes}


def render(scene, W, H, depth, want_fb=False, want_idx=False, pix_step=1, nthreads=0, sample_offset=(0.0, 0.0)):
    """scene: any object with .spheres [N,10], .lights [L,7], .ambient [3], .camera [7] (float64).
    Returns dict(rgb [H,W,3] uint8 bottom-row-first, fb, hit_idx [H,W,depth], shadow_mask, counters)."""
    sph = np.ascontiguousarray(scene.spheres, dtype=np.float64).reshape(-1, 10)
    lig = np.ascontiguousarray(scene.lights, dtype=np.float64).reshape(-1, 7)
    amb = np.ascontiguousarray(scene.ambient, dtype=np.float64)
    cam = np.ascontiguousarray(scene.camera, dtype=np.float64)
    pos, look = np.ascontiguousarray(cam[:3]), np.ascontiguousarray(cam[3:6])
    rgb = np.zeros((H, W, 3), dtype=np.uint8)
    fb = np.zeros((H, W, 3), dtype=np.float64) if want_fb else None
    d = max(depth, 1)
    hit = np.full((H, W, d), -2, dtype=np.int32) if want_idx else None
    mask = np.zeros((H, W, d), dtype=np.uint32) if want_idx else None
    cnt = Counters()
    lib().rto_set_sample_offset(float(sample_offset[0]), float(sample_offset[1]))
    rc = lib().rto_render(sph.ctypes.data, sph.shape[0], lig.ctypes.data, lig.shape[0], amb.ctypes.data,
                          pos.ctypes.data, look.ctypes.data, float(cam[6]), W, H, depth,
                          fb.ctypes.data if want_fb else None, rgb.ctypes.data,
                          hit.ctypes.data if want_idx else None, mask.ctypes.data if want_idx else None,
                          C.byref(cnt), pix_step, nthreads)
    lib().rto_set_sample_offset(0.0, 0.0)
    if rc != 0:
        raise RuntimeError("rto_render failed: %d" % rc)
    return {"rgb": rgb, "fb": fb, "hit_idx": hit, "shadow_mask": mask, "counters": cnt.as_dict()}


def intersect(origin, direction, center, radius):
    o = np.ascontiguousarray(origin, dtype=np.float64)
    d = np.ascontiguousarray(direction, dtype=np.float64)
    c = np.ascontiguousarray(center, dtype=np.float64)
    t = C.c_double()
    h = lib().rto_intersect(o.ctypes.data, d.ctypes.data, c.ctypes.data, float(radius), C.byref(t))
    return bool(h), t.value


def write_ppm(path, rgb_bottom_first):
    a = np.ascontiguousarray(rgb_bottom_first, dtype=np.uint8)
    H, W, _ = a.shape
    if lib().rto_write_ppm(os.fsencode(path), a.ctypes.data, W, H) != 0:
        raise RuntimeError("rto_write_ppm failed")


# ---- the reference's own binaries (oracle/_ref, built by oracle/Makefile) -------------------
def ref_available():
    return all(os.path.exists(os.path.join(REF_DIR, n)) for n in ("ray_serial", "ray_openmp", "ref_harness"))


def ref_harness(*args, env=None, cwd=None):
    e = dict(os.environ)
    if env:
        e.update(env)
    return subprocess.run([os.path.join(REF_DIR, "ref_harness")] + [str(a) for a in args], check=True,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=e, cwd=cwd).stdout


def ref_trace(scene_path, W, H, depth, out_bin):
    """Runs `ref_harness trace` and parses its dump: (hit_idx [H,W,D], shadow_mask, counters dict)."""
    ref_harness("trace", scene_path, W, H, depth, out_bin)
    raw = np.fromfile(out_bin, dtype=np.uint8)
    hdr = raw[:16].view(np.int32)
    cn = raw[16:48].view(np.uint64)
    n = int(hdr[0]) * int(hdr[1]) * int(hdr[2])
    hit = raw[48:48 + 4 * n].view(np.int32).reshape(int(hdr[1]), int(hdr[0]), int(hdr[2]))
    mask = raw[48 + 4 * n:48 + 8 * n].view(np.uint32).reshape(hit.shape)
    return hit, mask, {"closest_queries": int(cn[0]), "hits": int(cn[1]), "shadow_queries": int(cn[2]),
                       "occluded": int(cn[3])}
