#!/usr/bin/env python3
"""Static SASS opcode listing of the built library (cuobjdump -sass): per kernel the instruction count, code size and the
counts of the opcodes that prove what the kernels are made of (TMA bulk copies, mbarriers, packed FP32, FP64, 128-bit
global stores, local-memory spills ...).   usage: sass_opcodes.py [library.so] > profiles/r02_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cs420-ray-tracer_b200", "librt_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
WATCH = ["UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "LDG", "STG", "STG.E.128", "LDG.E.128",
         "LDL", "STL", "ATOMG", "REDG", "REDUX", "VOTE", "SHFL", "BAR", "MEMBAR", "CCTL", "ACQBULK", "UTMALDG", "HMMA", "UTCMMA", "CALL", "BRA"]
arch = re.findall(r"arch = (sm_\w+)", out)
print("library: %s   arch of every cubin: %s" % (os.path.basename(lib), sorted(set(arch))))
cur = None; cnt = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); cnt[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        op = m.group(1)
        cnt[cur]["_total"] += 1
        cnt[cur][op.split(".")[0]] += 1
        if op.startswith("STG.E.128") or op.startswith("LDG.E.128"): cnt[cur][".".join(op.split(".")[:3])] += 1
try:
    names = subprocess.run(["c++filt"] + list(cnt), stdout=subprocess.PIPE, text=True).stdout.split("\n")
except Exception:
    names = list(cnt)
for (k, c), nm in zip(cnt.items(), names):
    print("\n%s\n  %d instructions, %.1f KB" % (nm, c["_total"], c["_total"] * 16 / 1024.0))
    print("  " + "  ".join("%s %d" % (w, c[w]) for w in WATCH if c[w]))
