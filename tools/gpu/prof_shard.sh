#!/bin/bash
# full ncu captures of $KERNELS on shard 0 of an N-way split (what one rank of N computes)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=${TAG:-profshard}
for k in ${KERNELS:-sort_kernel order_kernel}; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP:-2} -c 1 -o gpurun_out/${T}_${k}_of${N:-8} python bench.py --workload ${WL:-c3} --shard-of ${N:-8} --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-configs --e2e-steps 1 > gpurun_out/${T}_ncu_$k.log 2>&1
tail -1 gpurun_out/${T}_ncu_$k.log
done
