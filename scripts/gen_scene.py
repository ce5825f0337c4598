#!/usr/bin/env python3
"""Procedural sphere scenes for BASELINE.json configs 4 and 5.

Writes a scene in the loader's text grammar (include/scene_loader.h:15-21) with %.6f fields, so
the oracle and the GPU parse the same decimals.  numpy default_rng(seed) = PCG64; draw order is
fixed (x, y, z, r, colour, diffuse-mask, metallic, shininess) so files are reproducible:
    python scripts/gen_scene.py 10000 420 out.txt          # config 4 (radius U(0.08,0.35))
    python scripts/gen_scene.py 100000 421 out.txt         # config 5 (radius U(0.03,0.12))
Camera / ambient / the 4 lights / the ground sphere are those of scenes/complex.txt (lines
196-209 of the reference file), the ground sphere is written LAST.
"""
import sys

import numpy as np

LIGHTS = [
    (0, 12, -10, 1.0, 1.0, 1.0, 0.4),
    (10, 10, -15, 1.0, 0.9, 0.8, 0.3),
    (-10, 10, -15, 0.8, 0.9, 1.0, 0.3),
    (5, 8, -20, 1.0, 1.0, 0.9, 0.25),
]


def generate(n, seed, rmin=None, rmax=None):
    """Returns (spheres [n,10], lights [4,7], ambient [3], camera [7]) as float64 arrays holding
    exactly the values the %.6f text parses to."""
    if rmin is None:
        rmin, rmax = (0.08, 0.35) if n <= 20000 else (0.03, 0.12)
    m = n - 1
    rng = np.random.default_rng(seed)
    x = rng.uniform(-30, 30, m)
    y = rng.uniform(-1.5, 12, m)
    z = rng.uniform(-80, -12, m)
    r = rng.uniform(rmin, rmax, m)
    col = rng.uniform(0.1, 1.0, (m, 3))
    diffuse = rng.random(m) < 0.5
    metallic = np.where(diffuse, 0.0, rng.uniform(0.1, 0.9, m))
    shin = rng.integers(5, 121, m).astype(np.float64)
    sph = np.column_stack([x, y, z, r, col, metallic, 1.0 - metallic, shin])
    ground = np.array([[0, -102, -20, 100, 0.3, 0.3, 0.3, 0.0, 1.0, 5]], dtype=np.float64)
    sph = np.vstack([sph, ground])
    sph = np.array([[float("%.6f" % v) for v in row] for row in sph.tolist()], dtype=np.float64)
    lights = np.array(LIGHTS, dtype=np.float64)
    ambient = np.array([0.1, 0.1, 0.12])
    camera = np.array([0, 3, 12, 0, 0, -20, 65], dtype=np.float64)
    return sph, lights, ambient, camera


def to_text(sph, lights, ambient, camera, header=""):
    out = ["# %s" % header] if header else []
    for s in sph:
        out.append("sphere " + " ".join("%.6f" % v for v in s))
    for li in lights:
        out.append("light " + " ".join("%.6f" % v for v in li))
    out.append("ambient " + " ".join("%.6f" % v for v in ambient))
    out.append("camera " + " ".join("%.6f" % v for v in camera))
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    if len(sys.argv) < 4:
        print(__doc__)
        sys.exit(2)
    n, seed, path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    text = to_text(*generate(n, seed), header="synthetic %d spheres, seed %d (scripts/gen_scene.py)" % (n, seed))
    with open(path, "w") as f:
        f.write(text)
