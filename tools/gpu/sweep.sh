#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-sweep}
run() {
  v=$1
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 20 --csv --log-file $O/${T}_launches_$v.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-configs --e2e-steps 1 > /dev/null 2>&1
  python - <<PY
import csv
try:
    rows=list(csv.reader(open("$O/${T}_launches_$v.csv")))
    hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
    t={}
    for r in rows[hi+1:]:
        n=r[4].split('(')[0].split('::')[-1]
        t.setdefault(n,[]).append(float(r[-1])/1e3)
    print("$v", {k: round(sorted(x)[len(x)//2],1) for k,x in t.items() if 'row_stats' not in k and 'scatter' not in k and 'array' not in k})
except Exception as e: print("$v launch list failed", e)
PY
}
run base
for v in $VARIANTS; do
  export YB_LIB_PATH=$PWD/yacrd_b200/libyacrd_b200_$v.so
  run $v
done
