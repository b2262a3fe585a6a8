#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-r2m}
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${T}_tests.txt 2>&1; echo "tests rc=$?" >> $O/${T}_tests.txt
tail -4 $O/${T}_tests.txt
for wl in c5 c3; do
timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-configs --e2e-steps 2 > $O/${T}_bench_$wl.json 2> $O/${T}_bench_$wl.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 16 --csv --log-file $O/${T}_launches_$wl.csv python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-configs --e2e-steps 1 > /dev/null 2>&1
python - <<PY
import csv, json
try:
    d=json.load(open("$O/${T}_bench_$wl.json")); print("$wl", "step %.4f frac %.3f parity %s up %.3f e2e %.1fM" % (d["ms_per_step"], d["roofline"]["frac"], d["parity"]["match"], d["ms_per_step_with_upload_kernels"], d["e2e"]["value"]/1e6), end=" ")
except Exception as e: print("$wl bench failed", e, end=" ")
rows=list(csv.reader(open("$O/${T}_launches_$wl.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
t={}
for r in rows[hi+1:]:
    n=r[4].split('(')[0].split('::')[-1]
    t.setdefault(n,[]).append(float(r[-1])/1e3)
print({k: round(sorted(x)[len(x)//2],1) for k,x in t.items()})
PY
done
