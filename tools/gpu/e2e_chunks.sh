#!/bin/bash
# streamed-path tests, then the end-to-end arm for several chunk sizes
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export TAG=${TAG:-r2r}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_stream.py -m gpu -x -q 2>&1 | tail -3
for wl in ${WLS:-c3 c5}; do
for ch in ${CHUNKS:-0 2000000 4000000 8000000 16000000 32000000}; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 6 --chunk-intervals $ch > $O/${TAG}_e2e_${wl}_$ch.json 2> $O/${TAG}_e2e_${wl}_$ch.err
  python - <<PY
import json
try:
    d=json.load(open("$O/${TAG}_e2e_${wl}_$ch.json")); print("$wl chunk $ch: e2e %.1f M reads/s  h2d %d d2h %d" % (d["e2e"]["value"]/1e6, d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"]))
except Exception as e: print("$wl chunk $ch failed", e)
PY
done
done
