#!/bin/bash
# step time (graph, CUDA events) of the base build and of the variants in $VARIANTS, full workload and shard 0 of 8
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r2u}
O=gpurun_out
mkdir -p $O
if [ -n "$TESTS" ]; then timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
run() {
  v=$1
  for wl in ${WLS:-c3}; do
  for n in ${NS:-1 8}; do
    timeout 300 python bench.py --workload $wl --shard-of $n --steps 40 --warmup 5 --no-cpu-baseline --no-configs --e2e-steps 1 > $O/${T}_${v}_${wl}_of$n.json 2> $O/${T}_${v}_${wl}_of$n.err
    python - <<PY
import json
try:
    d=json.load(open("$O/${T}_${v}_${wl}_of$n.json")); print("$v $wl 1/$n step %.4f parity %s first %.4f" % (d["ms_per_step"], d["parity"]["match"], d["details"]["first_step_ms"]))
except Exception as e: print("$v $wl 1/$n failed", e)
PY
  done
  done
}
run base
for v in $VARIANTS; do
  export YB_LIB_PATH=$PWD/yacrd_b200/libyacrd_b200_$v.so
  run $v
done
