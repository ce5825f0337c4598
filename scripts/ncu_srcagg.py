#!/usr/bin/env python3
"""Aggregates a `ncu --page source --csv --print-source cuda,sass` dump (optionally .gz) per source line: warp instructions,
average active lanes, samples; prints the top lines and per-file totals.   usage: ncu_srcagg.py dump.csv[.gz] [top]"""
import csv, gzip, io, sys, collections
fn = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = (gzip.open(fn, "rt") if fn.endswith(".gz") else open(fn)).read()
f = None; hdr = None; rows = []
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == "File Path": f = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "": continue
    rows.append((f, r))
iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed")
num = lambda x: int(x) if x.isdigit() else 0
tot = sum(num(r[iS]) for _, r in rows); toti = sum(num(r[iI]) for _, r in rows)
print("total samples %d, warp instructions %d" % (tot, toti))
byf = collections.defaultdict(lambda: [0, 0, 0])
for f, r in rows:
    byf[f][0] += num(r[iS]); byf[f][1] += num(r[iI]); byf[f][2] += num(r[iT])
for f, (s, i, t) in sorted(byf.items(), key=lambda kv: -kv[1][1]):
    print("  %-22s %5.1f%% smp %5.1f%% inst, %4.1f lanes" % (f, 100.0 * s / tot, 100.0 * i / toti, t / max(i, 1)))
for f, r in sorted(rows, key=lambda fr: -num(fr[1][iI]))[:top]:
    print("%5.1f%% smp %5.1f%% inst %4.1f lanes  %s:%s  %s" % (100.0 * num(r[iS]) / tot, 100.0 * num(r[iI]) / toti, num(r[iT]) / max(num(r[iI]), 1), f, r[0], r[1].strip()[:110]))
