#!/bin/bash
# what one rank of an N-GPU run computes, on one GPU: step time and launch list of shard 0 of an N-way split
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r2t}
O=gpurun_out
mkdir -p $O
for wl in ${WLS:-c3 c5}; do
for n in ${NS:-8 4}; do
timeout 300 python bench.py --workload $wl --shard-of $n --steps 30 --warmup 5 --no-cpu-baseline --no-configs --e2e-steps 1 > $O/${T}_${wl}_of$n.json 2> $O/${T}_${wl}_of$n.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file $O/${T}_launches_${wl}_of$n.csv python bench.py --workload $wl --shard-of $n --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-configs --e2e-steps 1 > /dev/null 2>&1
python - <<PY
import csv, json
try:
    d=json.load(open("$O/${T}_${wl}_of$n.json")); print("$wl 1/$n", "step %.4f parity %s first %.4f" % (d["ms_per_step"], d["parity"]["match"], d["details"]["first_step_ms"]), d["details"]["l2"][:20], end=" ")
except Exception as e: print("$wl bench failed", e, end=" ")
rows=list(csv.reader(open("$O/${T}_launches_${wl}_of$n.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
t={}
for r in rows[hi+1:]:
    k=r[4].split('(')[0].split('::')[-1]
    t.setdefault(k,[]).append(float(r[-1])/1e3)
print({k: round(sorted(x)[len(x)//2],1) for k,x in t.items()})
PY
done
done
