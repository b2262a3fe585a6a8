#!/bin/bash
# round 2, first GPU pass: parity, bench of the product library and of the A/B variants, launch list, one full capture
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r2a_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2a_tests.txt 2>&1; echo "tests rc=$?" >> $O/r2a_tests.txt
tail -5 $O/r2a_tests.txt
timeout 300 python bench.py --steps 50 --warmup 5 > $O/r2a_bench.json 2> $O/r2a_bench.err; echo "bench rc=$?"
for v in fma3 fma2 imad fma3i fma4i; do
  YB_LIB_PATH=$PWD/yacrd_b200/libyacrd_b200_$v.so timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 > $O/r2a_bench_$v.json 2> $O/r2a_bench_$v.err
  python - <<PY
import json
try:
    d=json.load(open("$O/r2a_bench_$v.json")); print("$v", d["ms_per_step"], d["roofline"]["frac"])
except Exception as e: print("$v failed", e)
PY
done
python -c "
import json; d=json.load(open('$O/r2a_bench.json')); print('base', d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r2a_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1 > $O/r2a_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sort_kernel -s 3 -c 1 -o $O/r2a_sort python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1 > $O/r2a_ncu_sort.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:order_kernel -s 3 -c 1 -o $O/r2a_order python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --e2e-steps 1 > $O/r2a_ncu_order.log 2>&1
ls -la $O | tail -20
