#!/usr/bin/env python3
"""Summarises an .ncu-rep: per-kernel headline metrics, stall reasons, and SASS hot spots by opcode.
usage: ncu_summary.py report.ncu-rep [kernel-regex]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_lsu.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")] or \
        [h for h in hdr if "warp_issue_stalled" in h and h.endswith("pct")]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("=== %s (id %s)" % (name, r[0]))
    for w in want:
        if w in hdr:
            print("  %-70s %s" % (w, r[hdr.index(w)]))
    st = []
    for h in hdr:
        if "issue_stalled" in h and ("pct" in h or "ratio" in h):
            try: st.append((float(r[hdr.index(h)]), h))
            except ValueError: pass
    for v, h in sorted(st, reverse=True)[:10]:
        print("  stall %-66s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_",""), v))
