#!/bin/bash
# per-warp timeline of the sorting kernel (YB_TRACE_CTA build) on shard 0 of an N-way split
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
export YB_LIB_PATH=$PWD/yacrd_b200/libyacrd_b200_trace.so
for n in ${NS:-8 1}; do
timeout 300 python bench.py --workload ${WL:-c3} --shard-of $n --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-configs --e2e-steps 1 --chunk-intervals 0 2>&1 | grep TRACE > $O/trace_of$n.txt
python - <<PY
import re,collections
L=[l.split() for l in open("$O/trace_of$n.txt")]
# group launches by t0 proximity
recs=[(int(l[1][4:]),int(l[3]),int(l[5]),int(l[7]),int(l[9]),int(l[11]),int(l[13])) for l in L]
recs.sort(key=lambda r:r[3])
groups=[]; 
for r in recs:
    if not groups or r[3]-groups[-1][0][3]>30000: groups.append([])
    groups[-1].append(r)
print("shard 1/$n: launches", len(groups))
for g in groups[-3:]:
    t0=min(r[3] for r in g)
    print(" val=%d  start skew %.1f us | first batch done %.1f..%.1f us | warp end %.1f..%.1f us | batches/warp %d..%d" % (g[0][0], (max(r[3] for r in g)-t0)/1e3, min(r[3]-t0+r[4] for r in g)/1e3, max(r[3]-t0+r[4] for r in g)/1e3, min(r[3]-t0+r[5] for r in g)/1e3, max(r[3]-t0+r[5] for r in g)/1e3, min(r[6] for r in g), max(r[6] for r in g)))
    for cta in (0,73,147):
        e=[ (r[3]-t0+r[5])/1e3 for r in g if r[1]==cta]
        if e: print("    cta %d warp ends: %s" % (cta, " ".join("%.0f"%x for x in sorted(e))))
PY
done
