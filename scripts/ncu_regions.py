#!/usr/bin/env python3
"""Aggregates the per-source-line samples / instructions of one kernel in an .ncu-rep by file and line RANGE.
usage: ncu_regions.py report.ncu-rep kernel-regex file:lo-hi=name ..."""
import csv, io, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
regions = []
for a in sys.argv[3:]:
    spec, name = a.split("=")
    f, rng = spec.split(":")
    lo, hi = rng.split("-")
    regions.append((f, int(lo), int(hi), name))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx,
                      "--launch-count", "1"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
fn = None; hdr = None; rows = []
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == "File Path": fn = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "": continue
    rows.append((fn, r))
iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed")
num = lambda x: int(x) if x.isdigit() else 0
tot = sum(num(r[iS]) for _, r in rows); toti = sum(num(r[iI]) for _, r in rows)
agg = collections.OrderedDict((n, [0, 0]) for *_, n in regions); agg["(other)"] = [0, 0]
byfile = collections.Counter()
for f, r in rows:
    ln = num(r[0]); hit = "(other)"
    for rf, lo, hi, n in regions:
        if f == rf and lo <= ln <= hi: hit = n; break
    agg[hit][0] += num(r[iS]); agg[hit][1] += num(r[iI])
    if hit == "(other)": byfile[f] += num(r[iS])
print("total samples %d, warp instructions %d" % (tot, toti))
for n, (s, i) in agg.items():
    print("%-28s %5.1f%% samples %5.1f%% instructions" % (n, 100.0 * s / tot, 100.0 * i / toti))
print("other by file:", dict(byfile.most_common(8)))
