#!/usr/bin/env python
"""Join an .ncu-rep SASS profile with nvdisasm line info of the in-tree library -> per-source-line table.
usage: python tools/ncu_lines.py prof.ncu-rep [kernel_substr] [top_n]"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else "fused_kernel"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(REPO, "yacrd_b200", "libyacrd_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
lines_of = []
for cb in cubin:
    out = subprocess.run(["nvdisasm", "-g", "-c", cb], capture_output=True, text=True).stdout
    on = False; cur = None; acc = []
    for l in out.split("\n"):
        if l.strip().startswith(".text."):
            on = kern in l
            if on: acc = []
        if not on: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", l):
            acc.append((cur, l.split("*/", 1)[1].strip()))
    if acc: lines_of = acc
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ix = {n: i for i, n in enumerate(h)}
data = rows[hi + 1:]
data = [r for r in data if len(r) == len(h)]
if len(data) != len(lines_of):
    print("warning: %d profiled SASS rows vs %d disassembled" % (len(data), len(lines_of)))
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
tot = [0.0, 0.0, 0.0]
for (loc, sass), r in zip(lines_of, data):
    ins = float(r[ix["Instructions Executed"]] or 0)
    smp = float(r[ix["Warp Stall Sampling (All Samples)"]] or 0)
    exc = float(r[ix["L1 Wavefronts Shared Excessive"]] or 0)
    a = agg[loc]; a[0] += ins; a[1] += smp; a[2] += exc
    tot[0] += ins; tot[1] += smp; tot[2] += exc
srcs = {}
def srcline(loc):
    if loc is None: return "?"
    f, n = loc
    p = os.path.join(REPO, "yacrd_b200", "csrc", f)
    if p not in srcs:
        try: srcs[p] = open(p).read().split("\n")
        except Exception: srcs[p] = []
    L = srcs[p]
    return L[n - 1].strip()[:100] if 0 < n <= len(L) else f
print("total warp-inst %.4g  samples %.4g  smem-excess-wavefronts %.4g" % tuple(tot))
print(" inst%  smpl%  bankx%  line  source")
key_ix = 1 if os.environ.get("BY_SAMPLES") else 0
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][key_ix])[:top]:
    print("%5.1f  %5.1f  %5.1f  %5s  %s" % (100 * a[0] / tot[0], 100 * a[1] / max(tot[1], 1), 100 * a[2] / max(tot[2], 1),
                                          loc[1] if loc else "?", srcline(loc)))
