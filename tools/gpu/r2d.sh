#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-r2d}
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-configs --e2e-steps 2 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; tail -3 $O/${T}_bench.err
python - <<PY
import json
try:
    d=json.load(open("$O/${T}_bench.json")); print("base", d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["parity"]["match"], d["ms_per_step_with_upload_kernels"])
except Exception as e: print("bench parse failed", e)
PY
for v in $VARIANTS; do
  YB_LIB_PATH=$PWD/yacrd_b200/libyacrd_b200_$v.so timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-configs --e2e-steps 2 > $O/${T}_bench_$v.json 2> $O/${T}_bench_$v.err
  python - <<PY
import json
try:
    d=json.load(open("$O/${T}_bench_$v.json")); print("$v", d["ms_per_step"], d["roofline"]["frac"], d["parity"]["match"])
except Exception as e: print("$v failed", e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file $O/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-configs --e2e-steps 1 > $O/${T}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sort_kernel<\(bool\)1>|sort_kernel<1>|sort_kernel<true>" -s 3 -c 1 -o $O/${T}_sort python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-configs --e2e-steps 1 > $O/${T}_ncu_sort.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:order_kernel -s 4 -c 1 -o $O/${T}_order python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-configs --e2e-steps 1 > $O/${T}_ncu_order.log 2>&1
ls -la $O | grep ${T}
