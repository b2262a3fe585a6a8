#!/bin/bash
# one full capture of the packed sort kernel (launches matching sort_kernel alternate wide, packed: odd indices are packed)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=${TAG:-prof}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sort_kernel -s ${SKIP:-5} -c 1 -o gpurun_out/${T}_sort python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-configs --e2e-steps 1 > gpurun_out/${T}_ncu_sort.log 2>&1
tail -3 gpurun_out/${T}_ncu_sort.log
