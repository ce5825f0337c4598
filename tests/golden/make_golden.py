#!/usr/bin/env python3
"""Generates tests/golden/* from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Needs /root/reference and oracle/_ref (built by `make -C oracle`).  What it writes:

  scenes/{simple,medium,complex}.txt   the reference's scene INPUTS with comments/blank lines
        stripped; every directive line keeps its tokens verbatim, in file order, so the parsed
        doubles are identical (checked below by rendering both and comparing md5s)
  scenes/quirks.txt                    a hand-made loader torture file (not from the reference)
  golden.json    md5 of the reference's own PPM output per (scene, W, H, depth) -- from the
        unmodified ray_serial at its built-in 1280x720 d10 and from ref_harness elsewhere --
        plus the ray counters of the reference's find_intersection / in_shadow calls
  small_<scene>.npz   160x90 depth-5 hit-index map, shadow mask and RGB from ref_harness
  quirks_dump.txt     `ref_harness dump` of scenes/quirks.txt (what the reference's loader parses)
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import oracle_py  # noqa: E402
import rtb200  # noqa: E402

QUIRKS = """# loader torture file (tests/golden/make_golden.py) -- exercises include/scene_loader.h:38-127
   # indented comment
\t
sphere 0 0 -20 2 1.0 0.0 0.0 0.0 1.0 10 trailing tokens are ignored
   sphere 3 0 -20 2 0.0 1.0 0.0 0.5 1.0 10
sphere 1 2 3
sphere 1 2 3 4 5 6 7 8 9 abc
sphere -3 0 -20 2e0 0.0 0.0 1.0 .25 1. 1e1
sphere 0 -102 -20 100 0.5 0.5 0.5 0.0 1.0 5xyz
light 10 10 -10 1.0 1.0 1.0 0.7
light 1 2 3 4 5
Light 1 2 3 4 5 6 7
ambient 0.2 0.2 0.2
ambient 0.1 0.1
ambient 0.1 0.15 0.2
camera 0 1 2 0 0 -20 50
camera 0 2 5 0 0 -20 60
bogus 1 2 3
sphere 5 5 -30 1 +0.5 0.5 0.5 0 1 7"""


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def normalise_scene(src, dst, name):
    keep = []
    for line in open(src):
        s = line.strip()
        if not s or s.startswith("#"):
            continue
        keep.append(" ".join(s.split()))
    with open(dst, "w") as f:
        f.write("# %s scene INPUT of shininglegend/cs420-ray-tracer (scenes/%s.txt), comments stripped by\n"
                "# tests/golden/make_golden.py; directive lines verbatim and in file order.\n" % (name, name))
        f.write("\n".join(keep) + "\n")


def main():
    assert oracle_py.ref_available(), "run `make -C oracle` first"
    os.makedirs(os.path.join(HERE, "scenes"), exist_ok=True)
    golden = {"generated_by": "tests/golden/make_golden.py", "images": [], "counters": []}
    tmp = tempfile.mkdtemp()
    for name in ("simple", "medium", "complex"):
        src = os.path.join(REF, "scenes", name + ".txt")
        dst = os.path.join(HERE, "scenes", name + ".txt")
        normalise_scene(src, dst, name)
        # unmodified ray_serial, built-in 1280x720 depth 10, on the ORIGINAL file and on the fixture
        for which, path in (("reference-file", src), ("fixture", dst)):
            subprocess.run([os.path.join(oracle_py.REF_DIR, "ray_serial"), path], cwd=tmp, check=True,
                           stdout=subprocess.DEVNULL)
            h = md5(os.path.join(tmp, "output_serial.ppm"))
            golden["images"].append({"scene": name, "W": 1280, "H": 720, "depth": 10, "md5": h,
                                     "by": "ray_serial (unmodified) on " + which})
        for (W, H, D) in ((1920, 1080, 5), (160, 90, 5), (97, 61, 3)):
            out = os.path.join(tmp, "h.ppm")
            oracle_py.ref_harness("render", dst, W, H, D, out, "omp")
            golden["images"].append({"scene": name, "W": W, "H": H, "depth": D, "md5": md5(out), "by": "ref_harness"})
        for (W, H, D) in ((1280, 720, 10), (1920, 1080, 5)):
            hit, mask, cn = oracle_py.ref_trace(dst, W, H, D, os.path.join(tmp, "t.bin"))
            alive = [int((hit[:, :, k] != -2).sum()) for k in range(D)]
            golden["counters"].append({"scene": name, "W": W, "H": H, "depth": D, **cn, "alive": alive,
                                       "hit_idx_sha256": hashlib.sha256(np.ascontiguousarray(hit).tobytes()).hexdigest(),
                                       "shadow_mask_sha256": hashlib.sha256(np.ascontiguousarray(mask).tobytes()).hexdigest()})
        hit, mask, cn = oracle_py.ref_trace(dst, 160, 90, 5, os.path.join(tmp, "t.bin"))
        out = os.path.join(tmp, "s.ppm")
        oracle_py.ref_harness("render", dst, 160, 90, 5, out)
        _, _, _, img = rtb200.read_ppm(out)
        np.savez_compressed(os.path.join(HERE, "small_%s.npz" % name), hit_idx=hit, shadow_mask=mask,
                            rgb_top_first=img.astype(np.uint8), counters=json.dumps(cn))
    # loader quirks
    q = os.path.join(HERE, "scenes", "quirks.txt")
    with open(q, "w") as f:
        f.write(QUIRKS)   # deliberately no trailing newline
    p = subprocess.run([os.path.join(oracle_py.REF_DIR, "ref_harness"), "dump", q], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    with open(os.path.join(HERE, "quirks_dump.txt"), "w") as f:
        f.write("\n".join(l for l in p.stdout.splitlines() if not l.startswith("Loaded scene")) + "\n")
    with open(os.path.join(HERE, "quirks_warnings.txt"), "w") as f:
        f.write(p.stderr)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(golden, f, indent=1)
    print(json.dumps(golden["images"], indent=1))


if __name__ == "__main__":
    main()
