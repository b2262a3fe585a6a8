#!/usr/bin/env python
"""bench.py — reads classified / s of the yacrd detect path (pile-up -> bad regions -> class) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload c3|c2|c5]

A "step" is one pass of the hot path over one resident batch: the detect kernels over the rank's CSR shard
(+ one all-gather of the 2-bit class bitmap when N > 1). Workload = BASELINE.json configs[2]/[3]: synthetic
2 M reads x mean 50 overlaps, ONT lengths, -c 4 -n 0.4 (seed 20261017), hash-sharded across
the N ranks (strong scaling: the 2 M-read job is fixed, per BASELINE.json configs[3]).

One JSON line on stdout (rank 0). `value` = whole-job reads/s with the CSR resident in HBM; `e2e` = the
same through the public API with pinned HOST buffers (H2D + kernels + D2H inside the timed region);
`roofline` = algorithmic bytes / step time against the measured HBM peak; `cpu_baseline` = the oracle
(C port of the reference's stack.rs) on this box's host cores, rank 0, N = 1 only.

--impl reference times that CPU port alone (the Rust reference cannot be built in this image: no cargo).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    # name: (global reads, mean intervals, profile, coverage, not_coverage, description)
    "c3": (2_000_000, 50.0, 0, 4, 0.4, "synthetic 2M reads x mean 50 overlaps, ONT lengths, -c 4 -n 0.4"),
    "c2": (100_000, 30.0, 0, 0, 0.8, "synthetic 100k reads x mean 30 overlaps, ONT lengths, -c 0 -n 0.8"),
    "c5": (500_000, 0.0, 1, 3, 0.4, "synthetic 500k reads PacBio Sequel lengths, skewed (max 5k ovl/read), -c 3 -n 0.4"),
}
L2_BYTES = 126 * 1024 * 1024
SEED = 20261017


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed regions run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        self.util = []

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
            }
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    self.util.append(nv.nvmlDeviceGetUtilizationRates(h).gpu)
                except Exception:
                    pass
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.01)
        except Exception as e:  # NVML missing: report nothing rather than invent
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        s = sorted(self.samples)
        # "under load": the upper half of the samples (idle samples before/after the loops sit at the bottom)
        load = s[len(s) // 2:] if s else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def run_reference(args, rank, world):
    """The reference's CPU algorithm (oracle/: C restatement of stack.rs + editor/mod.rs) on the host cores."""
    if rank != 0:
        return
    import yacrd_b200 as yb
    from oracle import yacrd_oracle as o
    n_glob, mean, profile, c, nn, desc = WORKLOADS[args.workload]
    sample_reads = min(n_glob, args.ref_sample)
    csr = yb.synth_csr(sample_reads, mean, profile=profile, seed=SEED)
    runner = o.PaddedRunner(csr.rowptr, csr.iv, csr.length)
    threads = o.max_threads()
    for _ in range(args.warmup):
        runner.run(c, nn, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        runner.run(c, nn, threads)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    v = sample_reads / dt
    sample = "first %d of %d reads of the workload (%d intervals) per step, %d threads" % (
        sample_reads, n_glob, csr.n_iv, threads)
    print(json.dumps({
        "impl": "reference", "metric": "reads classified/sec", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": desc, "seed": SEED, "note": "C port of yacrd 1.0.0 stack.rs:61-139 + editor/mod.rs:85-100 "
                   "(the Rust reference cannot be built here: no cargo/rustc); CSR in host memory -> classes + bad regions"},
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=0, help="0: min(steps, 10)")
    ap.add_argument("--ref-sample", type=int, default=500_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nccl-allgather", action="store_true", help="N > 1: NCCL all-gather instead of the peer-memory epilogue")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of one CUDA graph per step")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import yacrd_b200 as yb
    from yacrd_b200 import dist as ybd

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: yacrd_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world)

    n_glob, mean, profile, c, nn, desc = WORKLOADS[args.workload]
    t_gen = time.perf_counter()
    csr = yb.synth_csr(n_glob, mean, profile=profile, seed=SEED, shard=rank, n_shards=world)
    t_gen = time.perf_counter() - t_gen

    # ---- device-resident arm -------------------------------------------------------------------
    fm = yb.FullMemory(device=local_rank)
    fm.bind_csr(csr)
    fm.upload()
    fm.synchronize()
    _, _, counts = ybd.shard_layout(n_glob, world)
    slot = ybd.bitmap_bytes(int(counts.max()))
    # N > 1: the all-gather of the bitmap is fused into the kernels' epilogue over NVLink peer memory (CUDA IPC
    # buffers, yacrd_b200/dist.py:PeerGather); --nccl-allgather times the plain NCCL collective instead
    pg = None
    if use_dist and not args.nccl_allgather:
        try:
            pg = ybd.PeerGather(fm, slot)
        except Exception as e:
            sys.stderr.write("peer-memory all-gather unavailable (%r): using NCCL\n" % (e,))
        ok = torch.tensor([1 if pg is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if pg is not None and int(ok.item()) == 0:
            pg.close()
            pg = None
    gathered = pg.tensor() if pg is not None else torch.zeros(world, slot, dtype=torch.uint8, device=dev)
    if pg is None:
        fm.bind_device_bitmap(gathered[rank].data_ptr(), slot)
    # a non-default stream: the C ABI takes NULL as "the context's own stream", and CUDA events only see
    # the stream they are recorded on, so kernels, all-gather and events all go to this one
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    flush = None
    in_bytes = csr.nbytes
    if in_bytes < 2 * L2_BYTES:  # shard smaller than ~2x L2: flush L2 between timed iterations
        flush = torch.empty(2 * L2_BYTES, dtype=torch.uint8, device=dev)

    def step():
        fm.compute_device(c, nn, stream.cuda_stream)
        if use_dist and pg is None:
            ybd.allgather_bitmaps(gathered[rank], gathered)

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    launches_per_step = None
    graph = None
    if not args.no_graph:
        # the step is a handful of short kernels (+ one small collective): replay it as one CUDA graph so that the
        # timed loop is not bound by host launch latency (matters at N = 8, where a rank's share takes ~70 us)
        l0 = fm.stats()["kernel_launches"]
        step()
        barrier()
        launches_per_step = fm.stats()["kernel_launches"] - l0
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                fm.compute_device(c, nn, stream.cuda_stream)  # the library's launches only; the collective stays eager
            torch.cuda.set_stream(stream)
            torch.cuda.synchronize()
            graph.replay()
            torch.cuda.synchronize()
        except Exception as e:  # capture not possible (e.g. collective not capturable): time eager launches
            sys.stderr.write("CUDA graph capture failed, timing eager launches: %r\n" % (e,))
            graph = None
            torch.cuda.set_stream(stream)

    def run_step():
        if graph is not None:
            graph.replay()
            if use_dist and pg is None:
                ybd.allgather_bitmaps(gathered[rank], gathered)
        else:
            step()
    sampler = ClockSampler(torch.cuda._get_nvml_device_index(local_rank) if hasattr(torch.cuda, "_get_nvml_device_index") else local_rank)
    sampler.start()
    launches0 = fm.stats()["kernel_launches"]
    K = args.steps
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(K):
            run_step()
        e1.record(stream)
        barrier()
        ms_total = e0.elapsed_time(e1)
    else:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        barrier()
        for a, b in evs:
            flush.zero_()
            a.record(stream)
            run_step()
            b.record(stream)
        barrier()
        ms_total = sum(a.elapsed_time(b) for a, b in evs)
    launches = fm.stats()["kernel_launches"] - launches0
    if graph is not None:
        launches = launches_per_step * K  # replayed launches are not seen by the library's counter
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if use_dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / K
    value = n_glob / (ms_step * 1e-3)

    fm.download()
    st = fm.stats()
    n_gaps_local = st["n_gaps"]
    class_counts = [st["n_not_bad"], st["n_chimeric"], st["n_not_covered"]]
    alg_bytes = 8 * csr.n_iv + 13 * csr.n_reads + 8 * n_gaps_local  # this rank's launch
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (ms_step * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("%s_n%d" % (args.workload, world))
        except Exception:
            traffic = None

    # ---- end-to-end arm: public API, pinned host CSR in, host results out, every step ------------
    fm2 = yb.FullMemory(device=local_rank)
    ke = args.e2e_steps or min(K, 10)
    pg2 = ybd.PeerGather(fm2, slot) if pg is not None else None
    gathered2 = pg2.tensor() if pg2 is not None else torch.zeros(world, slot, dtype=torch.uint8, device=dev)
    host_gathered = torch.empty(world, slot, dtype=torch.uint8).pin_memory()

    def e2e_step():
        fm2.reset()
        fm2.bind_csr(csr)                              # host buffers (pinned)
        if pg2 is None:
            fm2.bind_device_bitmap(gathered2[rank].data_ptr(), slot)
        bp = yb.FromOverlap(fm2, c, nn)
        bp.compute_all_bad_part()                      # H2D + kernels + D2H of classes / bad-region CSR
        if use_dist:
            if pg2 is None:
                ybd.allgather_bitmaps(gathered2[rank], gathered2)
            host_gathered.copy_(gathered2, non_blocking=False)
        return int(bp.classes()[:16].sum()) + len(bp.gap_csr()[1])

    for _ in range(2):
        e2e_step()
    barrier()
    s0 = fm2.stats()
    t0 = time.perf_counter()
    for _ in range(ke):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    s1 = fm2.stats()
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if use_dist:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n_glob / (float(te.item()) / ke)
    h2d = (s1["h2d_bytes"] - s0["h2d_bytes"]) // ke
    d2h = (s1["d2h_bytes"] - s0["d2h_bytes"]) // ke + (world * slot if use_dist else 0)
    clocks = sampler.result()

    # ---- CPU baseline (rank 0, N = 1 only) -------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import yacrd_oracle as o
        sample_reads = min(n_glob, args.ref_sample)
        sub = yb.synth_csr(sample_reads, mean, profile=profile, seed=SEED)
        runner = o.PaddedRunner(sub.rowptr, sub.iv, sub.length)
        threads = o.max_threads()
        runner.run(c, nn, threads)
        reps, t0 = 0, time.perf_counter()
        while reps < 5 and time.perf_counter() - t0 < 15.0:
            runner.run(c, nn, threads)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        t1 = time.perf_counter()
        runner.run(c, nn, 1)
        dt1 = time.perf_counter() - t1
        cpu = {"value": sample_reads / dt, "unit": "reads/s", "cores": threads, "kind": "port",
               "sample": "first %d of %d reads (%d intervals), mean of %d passes, all %d host threads; "
                         "1 thread (the reference's default -t): %.0f reads/s" % (sample_reads, n_glob, sub.n_iv, reps,
                                                                                  threads, sample_reads / dt1)}

    if rank == 0:
        print(json.dumps({
            "metric": "reads classified/sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": K,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": desc, "seed": SEED, "reads_global": n_glob, "reads_rank0": csr.n_reads,
                       "intervals_rank0": csr.n_iv, "sharding": ("mix64(read index) % n_gpus, no data-path collective; the 2-bit class bitmap is all-gathered every step "
                                    + ("by peer stores from the kernels' epilogue + flag barrier (NVLink, CUDA IPC)" if pg is not None
                                       else "with one NCCL all-gather")) if use_dist else "single GPU",
                       "l2": "inputs larger than L2 (%.0f MB per rank)" % (in_bytes / 1e6) if flush is None
                       else "L2 flushed (256 MB memset) before every timed step; steps timed individually",
                       "classes_rank0": dict(zip(["NotBad", "Chimeric", "NotCovered"], class_counts)),
                       "gaps_rank0": n_gaps_local, "gen_seconds": round(t_gen, 2),
                       "launch": ("one CUDA graph per step" + (" (kernels) + eager NCCL all-gather" if use_dist and pg is None else "")) if graph is not None else "eager launches",
                       "per_upload": "once per uploaded CSR, outside the device-resident step and inside e2e: row statistics, "
                                     "interval validation (0 <= b < e <= len) and the size-class worklist (16 B per read), "
                                     "0.27 ms of kernels at 2 M reads (profiles/r1_v8_launches.csv)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "rank 0's shard; duration = whole step (CUDA events on the launching stream)"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": ke, "note": "reset + bind pinned host CSR + yb_compute_all_bad_part (H2D, kernels, D2H) per step"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }), flush=True)
    if use_dist:
        torch.cuda.synchronize()
        dist.barrier()
    for g_ in (pg, pg2):
        if g_ is not None:
            g_.close()
    fm.close()
    fm2.close()
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
