#!/usr/bin/env python
"""Executed warp-instructions per SASS opcode of one kernel in an .ncu-rep (source page, SASS view).
usage: python tools/ncu_opcodes.py prof.ncu-rep [top_n]"""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; ix = {n: i for i, n in enumerate(h)}
agg = collections.Counter(); tot = 0
for r in rows[hi + 1:]:
    if len(r) != len(h): continue
    ins = float(r[ix["Instructions Executed"]] or 0)
    sass = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+(\.[A-Z0-9_]+)*)", sass)
    op = m.group(2) if m else sass[:20]
    base = op.split(".")[0]
    key = base if base not in ("IMAD", "VIMNMX", "ISETP", "LOP3", "IADD3", "SHFL", "LDS", "STS", "SEL", "PRMT") else ".".join(op.split(".")[:2])
    agg[key] += ins; tot += ins
print("total warp-inst %.4g" % tot)
for k, v in agg.most_common(top):
    print("%6.2f%%  %10.3g  %s" % (100 * v / tot, v, k))
