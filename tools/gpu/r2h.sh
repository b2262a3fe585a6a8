#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-r2h}
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${T}_tests.txt 2>&1; echo "tests rc=$?" >> $O/${T}_tests.txt
tail -4 $O/${T}_tests.txt
run() {
  v=$1
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-configs --e2e-steps 2 > $O/${T}_bench_$v.json 2> $O/${T}_bench_$v.err
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 16 --csv --log-file $O/${T}_launches_$v.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-configs --e2e-steps 1 > /dev/null 2>&1
  python - <<PY
import json, csv
try:
    d=json.load(open("$O/${T}_bench_$v.json")); print("$v", "step %.4f frac %.3f parity %s up %.3f e2e %.1fM" % (d["ms_per_step"], d["roofline"]["frac"], d["parity"]["match"], d["ms_per_step_with_upload_kernels"], d["e2e"]["value"]/1e6), end=" ")
except Exception as e: print("$v bench failed", e, end=" ")
try:
    rows=list(csv.reader(open("$O/${T}_launches_$v.csv")))
    hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
    t={}
    for r in rows[hi+1:]:
        n=r[4].split('(')[0].split('::')[-1]
        t.setdefault(n,[]).append(float(r[-1])/1e3)
    print({k: round(sorted(x)[len(x)//2],1) for k,x in t.items()})
except Exception as e: print("launch list failed", e)
PY
}
run base
for v in $VARIANTS; do
  export YB_LIB_PATH=$PWD/yacrd_b200/libyacrd_b200_$v.so
  run $v
done
unset YB_LIB_PATH
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sort_kernel -s 3 -c 1 -o $O/${T}_sort python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-configs --e2e-steps 1 > $O/${T}_ncu_order.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:validate_kernel -s 0 -c 1 -o $O/${T}_validate python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-configs --e2e-steps 1 > $O/${T}_ncu_validate.log 2>&1
tail -2 $O/${T}_ncu_order.log 2>/dev/null
