#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-full}
timeout 900 python bench.py --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; tail -3 $O/${T}_bench.err
python - <<PY
import json
d=json.load(open("$O/${T}_bench.json"))
print("step %.4f frac %.3f e2e %.1fM parity %s up %.3f cpu %s" % (d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]/1e6, d["parity"], d["ms_per_step_with_upload_kernels"], d["cpu_baseline"]))
for c in d["configs"]: print(c["name"], "ms %.4f up %.4f frac %.3f" % (c["ms_per_step"], c["ms_per_step_with_upload_kernels"], c["roofline"]["frac"]), c["parity"]["match"], c.get("max_intervals_per_read"))
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/${T}_ref.json 2> $O/${T}_ref.err; echo "ref rc=$?"; head -c 600 $O/${T}_ref.json; echo
for wl in c5 c2; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 14 --csv --log-file $O/${T}_launches_$wl.csv python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-configs --e2e-steps 1 > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("$O/${T}_launches_$wl.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
t={}
for r in rows[hi+1:]:
    n=r[4].split('(')[0].split('::')[-1]
    t.setdefault(n,[]).append(float(r[-1])/1e3)
print("$wl", {k: round(sorted(x)[len(x)//2],1) for k,x in t.items()})
PY
done
