#!/usr/bin/env python
"""Per-source-line instruction / stall-sample table of one kernel of an .ncu-rep (CUDA-C correlation, all files).
usage: python tools/ncu_srclines.py prof.ncu-rep [launch_skip] [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
skip = sys.argv[2] if len(sys.argv) > 2 else "0"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
fname, h, L = "?", None, []
ti = ts = 0.0
for r in rows:
    if len(r) == 2 and r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        h = r
        ii, si = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
        continue
    if h is None or len(r) != len(h) or not r[0] or r[2] not in ("", "-"):
        continue  # SASS rows carry an address; source rows a line number
    try:
        ins, sm = float(r[ii] or 0), float(r[si] or 0)
    except ValueError:
        continue
    ti += ins
    ts += sm
    L.append((ins, sm, fname, r[0], r[1].strip()[:110]))
print("total warp-inst %.4g  samples %.4g" % (ti, ts))
print(" inst%  smpl%  file:line  source")
for ins, sm, f, l, s in sorted(L, reverse=True)[:top]:
    print("%5.1f  %5.1f  %s:%s  %s" % (100 * ins / max(ti, 1), 100 * sm / max(ts, 1), f, l, s))
