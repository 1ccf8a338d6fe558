#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics, stall reasons and the hottest source lines (needs -lineinfo)."""
import csv, io, re, subprocess, sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:3]:
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print("kernel:", d.get("Kernel Name", ("?",))[0][:100])
    keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct",
            "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
            "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum"]
    for k in keys:
        if k in d:
            print(f"  {k:75s} {d[k][0]:>18s} {d[k][1]}")
    print("  stalls (warps per issue):")
    for k in sorted(d):
        m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", k)
        if m and float(d[k][0]) > 0.03:
            print(f"     {m.group(1):25s} {float(d[k][0]):.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
fname = "?"
h = None
lines = []
tot_i = tot_w = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Name":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        h = r
        ci, cw = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
        cl, cwt = h.index("stall_long_sb"), h.index("stall_wait")
        continue
    if h is None or len(r) < len(h) or not r[0].isdigit():
        continue
    try:
        ni, nw, nl, nwt = float(r[ci] or 0), float(r[cw] or 0), float(r[cl] or 0), float(r[cwt] or 0)
    except ValueError:
        continue
    if ni == 0 and nw == 0:
        continue
    tot_i += ni; tot_w += nw
    lines.append((ni, nw, nl, nwt, fname, r[0], r[1].strip()[:120]))
print(f"source lines by stall samples (total warp-instructions {tot_i:.3g}, samples {tot_w:.3g})")
for ni, nw, nl, nwt, f, ln, s_ in sorted(lines, key=lambda t: -t[1])[:top]:
    print(f"  {100*ni/max(tot_i,1):5.1f}% inst {100*nw/max(tot_w,1):5.1f}% stall (long_sb {100*nl/max(tot_w,1):4.1f} wait {100*nwt/max(tot_w,1):4.1f})  {f}:{ln:>4s}  {s_}")
