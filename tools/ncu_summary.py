#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics + hottest source lines.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_lines]"""
import csv, io, subprocess, sys, re, collections

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
d = dict(zip(hdr, zip(units, vals)))
for k in KEYS:
    if k in d:
        print("%-75s %12s %s" % (k, d[k][1], d[k][0]))
print("-- stall reasons (warps per issue-active cycle) --")
st = [(float(v[1].replace(",", "")), k) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")]
for v, k in sorted(st, reverse=True)[:8]:
    print("  %-40s %.3f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# find header row
hi = next(i for i, r in enumerate(rows) if "Source" in r and any("Instructions Executed" in c for c in r))
h = rows[hi]
def col(name):
    for i, c in enumerate(h):
        if c == name: return i
    return None
ci_src, ci_inst = col("Source"), col("Instructions Executed")
ci_samp = col("Warp Stall Sampling (All Samples)") or col("Warp Stall Sampling (All Cycles)")
ci_line = col("#")
lines = []
tot_i = tot_s = 0
for r in rows[hi + 1:]:
    try:
        ins = float(r[ci_inst] or 0); smp = float(r[ci_samp] or 0) if ci_samp is not None else 0
    except Exception:
        continue
    tot_i += ins; tot_s += smp
    lines.append((smp, ins, r[ci_line] if ci_line is not None else "", r[ci_src].strip()[:110]))
print("-- top source lines by stall samples (samples%%, inst%%) -- total inst %.3g samples %.3g" % (tot_i, tot_s))
for smp, ins, ln, s in sorted(lines, reverse=True)[:top]:
    print("%5.1f%% %5.1f%%  %5s  %s" % (100 * smp / max(tot_s, 1), 100 * ins / max(tot_i, 1), ln, s))
