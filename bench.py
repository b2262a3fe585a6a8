#!/usr/bin/env python
"""bench.py — reads classified / s of the yacrd detect path (pile-up -> bad regions -> class) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload c3|c2|c5]

A "step" is one pass of the hot path over one resident batch: the detect kernels over the rank's CSR shard (with
N > 1 the all-gather of the 2-bit class bitmap is fused into the kernels' epilogue). Workload = BASELINE.json
configs[2]/[3]: synthetic 2 M reads x mean 50 overlaps, ONT lengths, -c 4 -n 0.4 (seed 20261017),
hash-sharded across the N ranks (strong scaling: the 2 M-read job is fixed, per BASELINE.json configs[3]).

One JSON line on stdout (rank 0). `value` = whole-job reads/s with the CSR resident in HBM; `e2e` = the same through
the public API with pinned HOST buffers (H2D + kernels + D2H inside the timed region); `roofline` = algorithmic
bytes / step time against the measured HBM peak; `parity` = the device results of THIS run compared with the CPU
oracle before the timed loop (every rank, and the gathered bitmap of the whole job when N > 1); `configs` = the same
measurement (shorter) for BASELINE.json's other synthetic configs; `cpu_baseline` = the oracle (C port of the
reference's stack.rs) on this box's host cores, rank 0, N = 1 only.

--impl reference times that CPU port alone on the full workload (the Rust reference cannot be built in this image: no
cargo). Its process never loads the product library (the generator comes from workload/libyacrd_synth.so).
"""
from __future__ import annotations

import argparse
import hashlib
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


def workload_config(name, world):
    """The `config` object: a description of the workload only, identical in the native and the reference arm."""
    n_glob, mean, profile, c, nn, desc = WORKLOADS[name]
    return {"workload": desc, "name": name, "seed": SEED, "reads_global": n_glob, "coverage": c, "not_coverage": nn,
            "sharding": "mix64(read index) % n_gpus" if world > 1 else "single GPU"}


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed regions run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

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
    """The reference's CPU algorithm (oracle/: C restatement of stack.rs + editor/mod.rs) on the host cores, on the
    whole workload, all host threads. Rank 0 only; this process never loads the product library."""
    if rank != 0:
        return
    import workload
    from oracle import yacrd_oracle as o
    n_glob, mean, profile, c, nn, desc = WORKLOADS[args.workload]
    sample_reads = n_glob if args.ref_sample <= 0 else min(n_glob, args.ref_sample)
    csr = workload.synth_csr(sample_reads, mean, profile=profile, seed=SEED)
    runner = o.PaddedRunner(csr.rowptr, csr.iv, csr.length)
    threads = args.ref_threads if args.ref_threads > 0 else o.max_threads()
    for _ in range(args.warmup):
        runner.run(c, nn, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        runner.run(c, nn, threads)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    v = sample_reads / dt
    sample = ("the whole workload" if sample_reads == n_glob else "first %d of %d reads" % (sample_reads, n_glob)) + \
             " (%d reads, %d intervals) per step, %d threads" % (sample_reads, csr.n_iv, threads)
    print(json.dumps({
        "impl": "reference", "metric": "reads classified/sec", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(args.workload, world),
        "details": {"note": "C port of yacrd 1.0.0 stack.rs:61-139 + editor/mod.rs:85-100 (the Rust reference cannot be built "
                            "here: no cargo/rustc); CSR in host memory -> classes + bad regions; generator: workload/libyacrd_synth.so"},
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


class Arm:
    """One workload on this rank: CSR shard, device-resident context, parity check, timed steps."""

    def __init__(self, name, rank, world, local_rank, dev, stream, args):
        import torch
        import yacrd_b200 as yb
        from yacrd_b200 import dist as ybd
        self.torch, self.yb, self.ybd = torch, yb, ybd
        self.name, self.rank, self.world, self.dev, self.stream, self.args = name, rank, world, dev, stream, args
        self.n_glob, mean, profile, self.c, self.nn, self.desc = WORKLOADS[name]
        t0 = time.perf_counter()
        # --shard-of N (development aid, one GPU): time shard 0 of an N-way split, i.e. what one rank of an N-GPU run computes
        n_shards = args.shard_of if (world == 1 and args.shard_of > 1) else world
        self.csr = yb.synth_csr(self.n_glob, mean, profile=profile, seed=SEED, shard=rank, n_shards=n_shards)
        self.t_gen = time.perf_counter() - t0
        self.fm = yb.FullMemory(device=local_rank)
        self.fm.bind_csr(self.csr)
        self.fm.upload()
        self.fm.synchronize()
        _, _, counts = ybd.shard_layout(self.n_glob, world)
        self.slot = ybd.bitmap_bytes(int(counts.max()))
        self.use_dist = world > 1
        # N > 1: the all-gather of the bitmap is fused into the kernels' epilogue over NVLink peer memory (CUDA IPC
        # buffers, yacrd_b200/dist.py:PeerGather); --nccl-allgather times the plain NCCL collective instead
        self.pg = None
        if self.use_dist and not args.nccl_allgather:
            import torch.distributed as dist
            try:
                self.pg = ybd.PeerGather(self.fm, self.slot)
            except Exception as e:
                sys.stderr.write("peer-memory all-gather unavailable (%r): using NCCL\n" % (e,))
            ok = torch.tensor([1 if self.pg is not None else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if self.pg is not None and int(ok.item()) == 0:
                self.pg.close()
                self.pg = None
        self.gathered = None
        if self.use_dist and self.pg is None:
            self.gathered = torch.zeros(world, self.slot, dtype=torch.uint8, device=dev)
            self.fm.bind_device_bitmap(self.gathered[rank].data_ptr(), self.slot)
        self.flush = None
        if self.csr.nbytes < 1.5 * L2_BYTES:  # shard not clearly larger than L2 (126 MB): flush L2 between timed iterations
            self.flush = torch.empty(2 * L2_BYTES, dtype=torch.uint8, device=dev)
        self.graph = None
        self.launches_per_step = None

    def step(self):
        self.fm.compute_device(self.c, self.nn, self.stream.cuda_stream)
        if self.use_dist and self.pg is None:
            self.ybd.allgather_bitmaps(self.gathered[self.rank], self.gathered)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.use_dist:
            import torch.distributed as dist
            dist.barrier()
            self.torch.cuda.synchronize()

    def gathered_now(self):
        """[world, slot] device tensor holding every rank's bitmap of the last step (after the consumer-side wait)."""
        if self.pg is not None:
            self.pg.wait(self.stream.cuda_stream)
            self.torch.cuda.synchronize()
            return self.pg.current()
        return self.gathered

    def parity(self):
        """Device results of one step vs the CPU oracle on the same shard (and, N > 1, the gathered bitmap of the whole
        job vs the oracle's classes of every shard). Outside every timed region."""
        from oracle import yacrd_oracle as o
        torch = self.torch
        self.step()
        self.barrier()
        gathered = self.gathered_now().cpu().numpy() if self.use_dist else None
        self.fm.download()
        cls, gp, gaps = self.fm.classes(), self.fm.gap_ptr(), self.fm.gaps()
        w_cls, w_gp, w_gaps = o.run_csr(self.csr.rowptr, self.csr.iv, self.csr.length, self.c, self.nn)
        ok = bool(np.array_equal(cls, w_cls) and np.array_equal(gp.astype(np.uint64), w_gp) and np.array_equal(gaps, w_gaps))
        h = hashlib.sha256()
        h.update(np.ascontiguousarray(cls).tobytes())
        h.update(np.ascontiguousarray(gaps).tobytes())
        hw = hashlib.sha256()
        hw.update(np.ascontiguousarray(w_cls).tobytes())
        hw.update(np.ascontiguousarray(w_gaps).tobytes())
        out = {"checked": True, "match": ok, "sha256": h.hexdigest(), "oracle_sha256": hw.hexdigest(),
               "what": "(class, bad regions) of rank 0's shard in read order vs oracle/yacrd_oracle.c on the same CSR"}
        if self.use_dist:
            import torch.distributed as dist
            # the oracle's bitmap of every shard, exchanged with a plain collective (checking only), against what the
            # kernels' epilogues left in this rank's gather buffer
            mine = np.zeros(self.slot, dtype=np.uint8)
            bm = self.ybd.pack_bitmap(w_cls)
            mine[:len(bm)] = bm
            want = torch.zeros(self.world, self.slot, dtype=torch.uint8, device=self.dev)
            dist.all_gather_into_tensor(want.view(-1), torch.from_numpy(mine).to(self.dev))
            want = want.cpu().numpy()
            _, _, counts = self.ybd.shard_layout(self.n_glob, self.world)
            g_ok = all(np.array_equal(self.ybd.unpack_bitmap(gathered[s], int(counts[s])), self.ybd.unpack_bitmap(want[s], int(counts[s])))
                       for s in range(self.world))
            flags = torch.tensor([1 if ok else 0, 1 if g_ok else 0], device=self.dev)
            dist.all_reduce(flags, op=dist.ReduceOp.MIN)
            out["match"] = bool(int(flags[0].item()))
            out["gathered_match"] = bool(int(flags[1].item()))
            out["what"] += "; all ranks agree; gathered_match: every rank's [world, slot] gather buffer decodes to the oracle's classes of the whole job"
        return out

    def capture(self):
        torch = self.torch
        l0 = self.fm.stats()["kernel_launches"]
        self.step()
        self.barrier()
        self.launches_per_step = self.fm.stats()["kernel_launches"] - l0
        if self.args.no_graph:
            return
        # the step is two or three short kernels: replay it as one CUDA graph so that the timed loop is not bound by host
        # launch latency (matters at N = 8, where a rank's share takes ~60 us)
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.stream):
                self.fm.compute_device(self.c, self.nn, self.stream.cuda_stream)  # the library's launches only
            torch.cuda.set_stream(self.stream)
            torch.cuda.synchronize()
            g.replay()
            torch.cuda.synchronize()
            self.graph = g
        except Exception as e:  # capture not possible: time eager launches
            sys.stderr.write("CUDA graph capture failed, timing eager launches: %r\n" % (e,))
            self.graph = None
            torch.cuda.set_stream(self.stream)

    def run_step(self):
        if self.graph is not None:
            self.graph.replay()
            if self.use_dist and self.pg is None:
                self.ybd.allgather_bitmaps(self.gathered[self.rank], self.gathered)
        else:
            self.step()

    def timed(self, K):
        """K steps, CUDA events on the launching stream, max over ranks. With peers the consumer-side wait for the last
        step's bitmaps is inside the timed region, so all K all-gathers have landed when the clock stops."""
        torch = self.torch
        stream = self.stream
        if self.flush is None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.barrier()
            e0.record(stream)
            for _ in range(K):
                self.run_step()
            if self.pg is not None:
                self.pg.wait(stream.cuda_stream)
            e1.record(stream)
            self.barrier()
            ms_total = e0.elapsed_time(e1)
        else:
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
            self.barrier()
            for a, b in evs:
                self.flush.zero_()
                a.record(stream)
                self.run_step()
                if self.pg is not None:
                    self.pg.wait(stream.cuda_stream)
                b.record(stream)
            self.barrier()
            ms_total = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms_total], dtype=torch.float64, device=self.dev)
        if self.use_dist:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / K

    def measure(self, K, W):
        """-> dict(value, ms_per_step, roofline, parity, ...) for this workload."""
        for _ in range(max(3, W)):
            self.step()
        self.barrier()
        par = self.parity()
        self.capture()
        ms_step = self.timed(K)
        self.fm.download()
        st = self.fm.stats()
        alg_bytes = 8 * self.csr.n_iv + 13 * self.csr.n_reads + 8 * st["n_gaps"]  # this rank's launch
        peak, peak_src = measured_peak()
        achieved = alg_bytes / (ms_step * 1e-3) / 1e9
        # one shot on a freshly uploaded CSR, device side: the per-upload kernels and the first step, which also tests
        # every interval (best of 3; every rank makes the calls, the step all-gathers)
        shots = [self.fm.time_one_shot(self.c, self.nn) for _ in range(3)]
        up_ms, first_ms = min(shots, key=lambda t: t[0] + t[1])
        self.barrier()
        iv_all = self.csr.n_iv
        if self.use_dist:
            import torch.distributed as dist
            t = self.torch.tensor([float(self.csr.n_iv)], dtype=self.torch.float64, device=self.dev)
            lst = [self.torch.zeros_like(t) for _ in range(self.world)]
            dist.all_gather(lst, t)
            iv_all = [int(x.item()) for x in lst]
        return {
            "name": self.name, "workload": self.desc, "n_gpus": self.world, "value": self.n_glob / (ms_step * 1e-3), "unit": "reads/s",
            "ms_per_step": ms_step, "steps": K,
            "ms_per_step_with_upload_kernels": up_ms + first_ms, "upload_kernels_ms": up_ms, "first_step_ms": first_ms,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_with_upload_kernels": alg_bytes / ((up_ms + first_ms) * 1e-3) / 1e9 / peak,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "rank 0's shard; duration = whole step (CUDA events on the launching stream)"},
            "parity": par,
            "reads_rank0": self.csr.n_reads, "intervals_rank0": self.csr.n_iv, "intervals_per_rank": iv_all,
            "classes_rank0": {"NotBad": st["n_not_bad"], "Chimeric": st["n_chimeric"], "NotCovered": st["n_not_covered"]},
            "gaps_rank0": st["n_gaps"], "max_intervals_per_read": st["max_intervals_per_read"],
            "l2": "inputs larger than L2 (%.0f MB per rank)" % (self.csr.nbytes / 1e6) if self.flush is None
            else "L2 flushed (256 MB memset) before every timed step; steps timed individually",
            "launch": "one CUDA graph per step" if self.graph is not None else "eager launches",
            "gpu_launches_per_step": int(self.launches_per_step), "gen_seconds": round(self.t_gen, 2),
        }

    def close(self):
        if self.pg is not None:
            self.pg.close()
        self.fm.close()
        self.csr.free()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=0, help="0: min(steps, 10)")
    ap.add_argument("--ref-sample", type=int, default=0, help="reference arm / cpu baseline: reads per step (0 = the whole workload)")
    ap.add_argument("--ref-threads", type=int, default=0, help="reference arm: host threads (0 = all; 1 = the reference's default -t)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short measurements of the other BASELINE configs")
    ap.add_argument("--nccl-allgather", action="store_true", help="N > 1: NCCL all-gather instead of the peer-memory epilogue")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of one CUDA graph per step")
    ap.add_argument("--shard-of", type=int, default=0, help="development aid (one GPU): run shard 0 of an N-way split; `value` is then meaningless")
    ap.add_argument("--chunk-intervals", type=int, default=16_000_000,
                    help="e2e arm: streamed batches of about this many intervals (128 MB of intervals); 0 = one shot")
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

    # a non-default stream: the C ABI takes NULL as "the context's own stream", and CUDA events only see
    # the stream they are recorded on, so kernels and events all go to this one
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sampler = ClockSampler(torch.cuda._get_nvml_device_index(local_rank) if hasattr(torch.cuda, "_get_nvml_device_index") else local_rank)
    sampler.start()

    K = args.steps
    arm = Arm(args.workload, rank, world, local_rank, dev, stream, args)
    main_res = arm.measure(K, args.warmup)
    n_glob, mean, profile, c, nn, desc = WORKLOADS[args.workload]
    csr = arm.csr

    # ---- end-to-end arm: public API, pinned host CSR in, host results out, every step ------------
    fm2 = yb.FullMemory(device=local_rank)
    ke = args.e2e_steps or min(K, 10)
    streamed = args.chunk_intervals > 0
    fm2.set_chunk_intervals(args.chunk_intervals)
    # streamed batches keep the class bitmap on the host: with N > 1 it is all-gathered by NCCL after the last chunk
    pg2 = ybd.PeerGather(fm2, arm.slot) if arm.pg is not None and not streamed else None
    gathered2 = None if pg2 is not None or not use_dist else torch.zeros(world, arm.slot, dtype=torch.uint8, device=dev)
    host_gathered = torch.empty(world, arm.slot, dtype=torch.uint8).pin_memory()

    def e2e_step():
        fm2.reset()
        fm2.bind_csr(csr)                              # host buffers (pinned)
        if use_dist and pg2 is None and not streamed:
            fm2.bind_device_bitmap(gathered2[rank].data_ptr(), arm.slot)
        bp = yb.FromOverlap(fm2, c, nn)
        bp.compute_all_bad_part()                      # H2D + kernels + D2H of classes / bad-region CSR
        if use_dist:
            if streamed:
                bm = torch.from_numpy(bp.class_bitmap())
                gathered2[rank, :bm.numel()].copy_(bm, non_blocking=True)
            if pg2 is None:
                ybd.allgather_bitmaps(gathered2[rank], gathered2)
                host_gathered.copy_(gathered2, non_blocking=False)
            else:
                pg2.wait()
                fm2.synchronize()
                host_gathered.copy_(pg2.current(), non_blocking=False)
        return int(bp.classes()[:16].sum()) + len(bp.gap_csr()[1])

    for _ in range(2):
        e2e_step()
    arm.barrier()
    s0 = fm2.stats()
    t0 = time.perf_counter()
    for _ in range(ke):
        e2e_step()
    arm.barrier()
    e2e_s = time.perf_counter() - t0
    s1 = fm2.stats()
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if use_dist:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n_glob / (float(te.item()) / ke)
    h2d = (s1["h2d_bytes"] - s0["h2d_bytes"]) // ke
    d2h = (s1["d2h_bytes"] - s0["d2h_bytes"]) // ke + (world * arm.slot if use_dist else 0)
    if use_dist and streamed:
        h2d += arm.slot
    clocks = sampler.result()
    if pg2 is not None:
        pg2.close()
    fm2.close()

    # ---- CPU baseline (rank 0, N = 1 only): the oracle on the host cores, a bounded sample of the same workload ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import yacrd_oracle as o
        sample_reads = min(n_glob, 500_000)
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
        sub.free()
    launches = main_res["gpu_launches_per_step"] * K
    traffic = None
    tp = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("%s_n%d" % (args.workload, world))
        except Exception:
            traffic = None
    arm.close()

    # ---- the other synthetic configs of BASELINE.json, shorter: C2 (one GPU by definition) and C5 (at this N) ----
    configs = [{k: main_res[k] for k in ("name", "workload", "n_gpus", "value", "unit", "ms_per_step", "ms_per_step_with_upload_kernels",
                                         "roofline", "parity", "intervals_per_rank", "max_intervals_per_read")}]
    if not args.no_configs:
        for name in ("c2", "c5"):
            if name == args.workload or (name == "c2" and world > 1):
                continue
            a2 = Arm(name, rank, world, local_rank, dev, stream, args)
            r2 = a2.measure(min(K, 20), 3)
            configs.append({k: r2[k] for k in ("name", "workload", "n_gpus", "value", "unit", "ms_per_step", "ms_per_step_with_upload_kernels",
                                               "roofline", "parity", "intervals_per_rank", "max_intervals_per_read", "l2")})
            a2.close()

    if rank == 0:
        roof = dict(main_res["roofline"])
        roof["traffic"] = traffic
        print(json.dumps({
            "metric": "reads classified/sec", "value": main_res["value"], "unit": "reads/s", "n_gpus": world, "steps": K,
            "warmup": max(3, args.warmup), "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config(args.workload, world),
            "ms_per_step_with_upload_kernels": main_res["ms_per_step_with_upload_kernels"],
            "details": {k: main_res[k] for k in ("reads_rank0", "intervals_rank0", "intervals_per_rank", "classes_rank0", "gaps_rank0",
                                                 "l2", "launch", "gen_seconds", "upload_kernels_ms", "first_step_ms")} | {
                "allgather": ("the 2-bit class bitmap is all-gathered every step "
                              + ("by peer stores from the ordering kernel's epilogue + per-rank flags (NVLink, CUDA IPC), no separate collective"
                                 if arm.pg is not None else "with one NCCL all-gather")) if use_dist else "single GPU",
                "per_upload": "ms_per_step_with_upload_kernels = one shot on a fresh CSR, device side: the per-upload kernels (row "
                              "statistics, size-class worklist: upload_kernels_ms) + the first detect step, which also tests every "
                              "interval 0 <= b < e <= len (first_step_ms); ms_per_step is a later step on the resident CSR"},
            "roofline": roof,
            "parity": main_res["parity"],
            "configs": configs,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": ke, "chunk_intervals": args.chunk_intervals,
                    "note": "reset + bind pinned host CSR + yb_compute_all_bad_part (H2D, kernels, D2H) per step"
                            + ("; streamed: row chunks on two lanes, transfers overlapped with the kernels" if streamed else "; one shot")},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }), flush=True)
    if use_dist:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
