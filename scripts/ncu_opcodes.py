#!/usr/bin/env python3
"""EXECUTED instruction mix of every kernel in an .ncu-rep (captured with --set full --import-source on): dynamic
thread-instruction counts per SASS opcode from the source page, the executed FP32 / FP64 flops that follow from them
(FFMA = 2, FFMA2 = 4, FMUL/FADD = 1, FMUL2/FADD2 = 2 per thread instruction; DFMA = 2, DMUL/DADD = 1) and, with the raw
page's duration and counters, the executed fraction of an FP32 peak.
usage: ncu_opcodes.py report.ncu-rep [peak_tflops] [--json out.json]"""
import collections, csv, io, json, subprocess, sys

rep = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 72.4
jout = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
FLOPS32 = {"FFMA": 2, "FFMA2": 4, "FMUL": 1, "FMUL2": 2, "FADD": 1, "FADD2": 2}
FLOPS64 = {"DFMA": 2, "DMUL": 1, "DADD": 1}

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
col = lambda name: hdr.index(name) if name in hdr else None
kernels = []
for r in rows[2:]:
    g = lambda n: float(r[col(n)].replace(",", "")) if col(n) is not None and r[col(n)] not in ("", "n/a") else None
    kernels.append({"id": int(r[0]), "name": r[col("Kernel Name")], "us": g("gpu__time_duration.sum") / 1e3 if (g("gpu__time_duration.sum") or 0) > 1e3 else g("gpu__time_duration.sum"),
                    "fma_pipe_pct": g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"), "issue_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "occupancy_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"), "l1_hit_pct": g("l1tex__t_sector_hit_rate.pct"),
                    "icache_hit_pct": g("sm__icc_request_hit_rate.pct"), "smem_wavefronts": g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
                    "smem_bank_conflicts": g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
                    "dram_read_mb": g("dram__bytes_read.sum"), "dram_write_mb": g("dram__bytes_write.sum"), "regs": g("launch__registers_per_thread")})
unit_ix = None
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
sec = -1; h2 = None
mix = [collections.Counter() for _ in kernels]
nsec = src.count('"Kernel Name"')
per = max(1, nsec // max(1, len(kernels)))          # ncu prints each launch's listing `per` times (one per view)
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "Kernel Name": sec += 1; continue
    if r[0] == "Address": h2 = r; continue
    cur = sec // per
    if h2 is None or sec < 0 or sec % per or cur >= len(kernels) or not r[0].startswith("0x"): continue
    ins = r[h2.index("Source")].strip()
    if ins.startswith("@"): ins = ins.split(None, 1)[1] if " " in ins else ins
    op = ins.split()[0] if ins else "?"
    try: n = int(r[h2.index("Predicated-On Thread Instructions Executed")])
    except ValueError: n = 0
    mix[cur][op] += n
out = []
tot32 = tot64 = tott = 0.0
for k, m in zip(kernels, mix):
    base = collections.Counter()
    for op, n in m.items(): base[op.split(".")[0]] += n
    f32 = sum(n * FLOPS32.get(op, 0) for op, n in base.items()); f64 = sum(n * FLOPS64.get(op, 0) for op, n in base.items())
    total = sum(base.values())
    k.update({"thread_instructions": total, "fp32_flops_executed": f32, "fp64_flops_executed": f64,
              "executed_fp32_frac_of_peak": f32 / (k["us"] * 1e-6) / (peak * 1e12) if k["us"] else None,
              "top_opcodes": [(op, n, round(100.0 * n / max(total, 1), 2)) for op, n in base.most_common(14)],
              "select": {op: base.get(op, 0) for op in ("FFMA2", "FFMA", "FMUL2", "FMUL", "FADD2", "FADD", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "LDG", "STG", "LDL", "STL", "UBLKCP", "SYNCS", "ATOMG", "REDUX", "VOTE", "SHFL", "BRA")}})
    tot32 += f32; tot64 += f64; tott += k["us"] or 0
    out.append(k)
    print("=== [%d] %s  %.1f us, %s regs" % (k["id"], k["name"][:60], k["us"], k["regs"]))
    print("    executed FP32 %.3f GFLOP = %.1f %% of %.1f TFLOP/s; FP64 %.3f GFLOP; FMA pipe %.1f %%, issue %.1f %%, occupancy %.1f %%, L1 hit %.1f %%, icache hit %.1f %%" % (
        f32 * 1e-9, 100 * (k["executed_fp32_frac_of_peak"] or 0), peak, f64 * 1e-9, k["fma_pipe_pct"] or 0, k["issue_pct"] or 0, k["occupancy_pct"] or 0, k["l1_hit_pct"] or 0, k["icache_hit_pct"] or 0))
    print("    shared-memory wavefronts %s (bank conflicts %s), DRAM read %s write %s MB" % (k["smem_wavefronts"], k["smem_bank_conflicts"], k["dram_read_mb"], k["dram_write_mb"]))
    print("    " + "  ".join("%s %.1f%%" % (op, p) for op, n, p in k["top_opcodes"]))
print("=== all %d launches: %.1f us (serialised, cold cache), executed FP32 %.3f GFLOP = %.1f %% of peak over that time; FP64 %.3f GFLOP" % (
    len(kernels), tott, tot32 * 1e-9, 100 * tot32 / (tott * 1e-6) / (peak * 1e12) if tott else 0, tot64 * 1e-9))
if jout:
    json.dump({"source": rep.split("/")[-1], "peak_tflops": peak, "kernels": out, "sum_us": tott, "fp32_flops_executed": tot32, "fp64_flops_executed": tot64,
               "executed_fp32_frac_of_peak": tot32 / (tott * 1e-6) / (peak * 1e12) if tott else None}, open(jout, "w"), indent=1)
