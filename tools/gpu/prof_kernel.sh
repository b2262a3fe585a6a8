#!/bin/bash
# one full ncu capture of kernel $KERNEL (regex) of bench workload $WL, launch index $SKIP
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=${TAG:-prof}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KERNEL} -s ${SKIP:-2} -c 1 -o gpurun_out/${T}_${KERNEL} python bench.py --workload ${WL:-c3} --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-configs --e2e-steps 1 > gpurun_out/${T}_ncu_${KERNEL}.log 2>&1
tail -2 gpurun_out/${T}_ncu_${KERNEL}.log
