#!/usr/bin/env python3
"""benchmark.sh-compatible report for the drop-in binary.

Writes the CSV the reference's scripts/benchmark.sh writes -- header
`Implementation,Scene,Threads,Iteration,Time(s),Pixels/s,Speedup` (scripts/benchmark.sh:29) -- and the same
"Average Execution Times" summary table (scripts/benchmark.sh:193-236), with three corrections of the reference
script: OpenMP rows use the binary's `OpenMP time:` line (the script's `head -1` picks `Serial time:`), Pixels/s uses the resolution the binaries really render (1280x720 for every scene, src/main.cpp:95-96;
the script assumes 640x480 / 800x600, scripts/benchmark.sh:91-101) and Speedup is computed against the serial
mean instead of being written as 1.0.  Time(s) is the program's own `... time: X seconds` line, as the
reference's extract_time() greps it (scripts/benchmark.sh:32-34).

    python scripts/benchmark_csv.py [--iterations 3] [--threads 1,2,4,8] [--out benchmark_results]

CUDA rows come from cs420-ray-tracer_b200/ray_cuda (needs a B200); Serial / OpenMP rows from oracle/_ref
(the unmodified reference binaries) when they are present -- as a baseline on this box's host cores.
"""
import argparse
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = ["simple.txt", "medium.txt", "complex.txt"]
W, H = 1280, 720


def run(cmd, env=None, cwd=None, prefer=None):
    """The program's own timing line.  ray_openmp prints BOTH `Serial time:` and `OpenMP time:` (src/main.cpp:161,203);
    the reference script's `head -1` takes the serial one for its OpenMP rows -- `prefer` picks the right line."""
    e = dict(os.environ)
    if env:
        e.update(env)
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=e, cwd=cwd).stdout
    m = [l for l in out.splitlines() if re.search(r"(time:|seconds)", l)]
    if prefer:
        m = [l for l in m if prefer in l] or m
    return float(re.search(r"[0-9]+\.[0-9]+(e-?[0-9]+)?", m[0]).group(0)) if m else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iterations", type=int, default=3)
    ap.add_argument("--threads", default="1,2,4,8")
    ap.add_argument("--out", default=os.path.join(ROOT, "benchmark_results"))
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    path = os.path.join(args.out, "benchmark_%s.csv" % time.strftime("%Y%m%d_%H%M%S"))
    scenes_dir = os.path.join(ROOT, "tests", "golden", "scenes")
    ref = os.path.join(ROOT, "oracle", "_ref")
    cuda = os.path.join(ROOT, "cs420-ray-tracer_b200", "ray_cuda")
    rows = []          # (impl, scene, threads, iteration, seconds)
    tmp = args.out     # binaries write output_*.ppm into cwd
    for scene in SCENES:
        sp = os.path.join(scenes_dir, scene)
        if not args.no_cpu and os.path.exists(os.path.join(ref, "ray_serial")):
            for it in range(1, args.iterations + 1):
                rows.append(("Serial", scene, 1, it, run([os.path.join(ref, "ray_serial"), sp], cwd=tmp)))
            for t in [int(x) for x in args.threads.split(",")]:
                for it in range(1, args.iterations + 1):
                    rows.append(("OpenMP", scene, t, it, run([os.path.join(ref, "ray_openmp"), sp], env={"OMP_NUM_THREADS": str(t)}, cwd=tmp,
                                                            prefer="OpenMP time")))
        if os.path.exists(cuda):
            for it in range(1, args.iterations + 1):
                rows.append(("CUDA", scene, 1, it, run([cuda, sp, "--frames", "20"], cwd=tmp)))
    rows = [r for r in rows if r[4]]
    mean = {}
    for impl, scene, t, _, sec in rows:
        mean.setdefault((impl, scene, t), []).append(sec)
    mean = {k: sum(v) / len(v) for k, v in mean.items()}
    with open(path, "w") as f:
        f.write("Implementation,Scene,Threads,Iteration,Time(s),Pixels/s,Speedup\n")
        for impl, scene, t, it, sec in rows:
            base = mean.get(("Serial", scene, 1))
            f.write("%s,%s,%d,%d,%.6f,%.0f,%s\n" % (impl, scene, t, it, sec, W * H / sec, "%.2f" % (base / sec) if base else ""))
    print("Average Execution Times (seconds):")
    print("-----------------------------------")
    print("%-15s %-12s %-12s %-12s" % ("Implementation", "Simple", "Medium", "Complex"))
    print("-----------------------------------")
    keys = sorted({(k[0], k[2]) for k in mean}, key=lambda k: (["Serial", "OpenMP", "CUDA"].index(k[0]), k[1]))
    for impl, t in keys:
        name = impl if impl != "OpenMP" else "OpenMP (%dt)" % t
        print("%-15s" % name + "".join(" %-12s" % ("%.6f" % mean[(impl, s, t)] if (impl, s, t) in mean else "-") for s in SCENES))
    print("-----------------------------------")
    print("CSV: %s" % path)
    return 0


if __name__ == "__main__":
    sys.exit(main())
