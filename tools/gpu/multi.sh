#!/bin/bash
# multi-GPU pass (gpurun --gpus N): the peer all-gather on one GPU per rank, then bench.py at N ranks (parity gate inside)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-multi}
N=${N:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader > $O/${T}_gpus.txt; cat $O/${T}_gpus.txt
timeout 600 python -m pytest tests/test_gpu_peer.py -m gpu -x -q -rs > $O/${T}_peer_tests.txt 2>&1; echo "peer tests rc=$?"; tail -5 $O/${T}_peer_tests.txt
for n in $NS; do
  for wl in $WLS; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --workload $wl --steps 30 --warmup 5 --no-configs > $O/${T}_bench_${wl}_n$n.json 2> $O/${T}_bench_${wl}_n$n.err
    echo "bench $wl n=$n rc=$?"; tail -2 $O/${T}_bench_${wl}_n$n.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/${T}_bench_${wl}_n$n.json").read().strip().split("\n")[-1])
    print("$wl n=$n step %.4f value %.3g e2e %.1fM parity %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"]/1e6, {k:v for k,v in d["parity"].items() if k in ("match","gathered_match")}), d["details"]["intervals_per_rank"])
except Exception as e: print("parse failed", e)
PY
  done
done
