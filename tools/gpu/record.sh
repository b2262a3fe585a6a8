#!/bin/bash
# the round's last record: tests, the default bench line, launch lists, full captures of sort_kernel and bigscan_kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${TAG:-r2final2}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/${T}_tests.txt 2>&1; echo "tests rc=$?" >> $O/${T}_tests.txt; tail -3 $O/${T}_tests.txt
timeout 900 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --workload c5 --steps 50 --warmup 5 --no-cpu-baseline --no-configs > $O/${T}_bench_c5.json 2> $O/${T}_bench_c5.err
for wl in c3 c5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/${T}_launches_$wl.csv python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-configs --e2e-steps 1 > /dev/null 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/${T}_launches_c3_shard1of8.csv python bench.py --workload c3 --shard-of 8 --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-configs --e2e-steps 1 > /dev/null 2>&1
TAG=$T KERNEL=sort_kernel SKIP=2 WL=c3 bash tools/gpu/prof_kernel.sh
TAG=$T KERNEL=bigscan_kernel SKIP=2 WL=c5 bash tools/gpu/prof_kernel.sh
python tools/e2e_breakdown.py > $O/${T}_e2e_breakdown.txt 2>&1; tail -6 $O/${T}_e2e_breakdown.txt
python - <<PY
import json
for f in ("bench","bench_c5"):
    try:
        d=json.loads(open("$O/${T}_%s.json"%f).read().strip().split("\n")[-1])
        print(f, "value %.4g %s ms %.4f oneshot %.4f" % (d["value"], d["unit"], d["ms_per_step"], d["ms_per_step_with_upload_kernels"]), "e2e %.4g" % d["e2e"]["value"], "frac", d.get("roofline",{}).get("frac"), "parity", d.get("parity",{}).get("match"), [ (c["name"], round(c["ms_per_step"],4)) for c in d.get("configs",[])])
    except Exception as e: print(f, "failed", e)
PY
