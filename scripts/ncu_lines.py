#!/usr/bin/env python3
"""Per-source-line hot spots of one kernel in an .ncu-rep (cuda,sass correlated view).
usage: ncu_lines.py report.ncu-rep kernel-regex [top] [smp|inst]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
by = sys.argv[4] if len(sys.argv) > 4 else "smp"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx,
                      "--launch-count", "1"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
fn = None; hdr = None; rows = []
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == "File Path": fn = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "": continue   # SASS rows
    rows.append((fn, r))
iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed")
num = lambda x: int(x) if x.isdigit() else 0
tot = sum(num(r[iS]) for _, r in rows); toti = sum(num(r[iI]) for _, r in rows)
print("total samples %d, warp instructions %d" % (tot, toti))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
key = iS if by == "smp" else iI
agg = sorted(rows, key=lambda fr: -num(fr[1][key]))
for f, r in agg[:top]:
    st = sorted(((num(r[i]), hdr[i][6:]) for i in stall_cols if num(r[i]) > 0), reverse=True)[:3]
    print("%5.1f%% smp %5.1f%% inst  %s:%s  %-90s %s" % (100.0 * num(r[iS]) / tot, 100.0 * num(r[iI]) / toti, f, r[0], r[1].strip()[:90],
                                                 " ".join("%s=%d" % (n, v) for v, n in st)))
