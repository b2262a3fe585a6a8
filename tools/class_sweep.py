#!/usr/bin/env python
"""Per-size-class throughput of the detect step (run on the B200 box): the synthetic workload restricted to rows whose
interval count k lies in one range, replicated to ~24 M intervals. Prints ns per row and ps per interval per range.
usage: python tools/class_sweep.py [-c 4] [lo-hi ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yacrd_b200 as yb  # noqa: E402
from yacrd_b200.api import PinnedCsr  # noqa: E402


def subset(base, lo, hi, target_iv):
    k = np.diff(base.rowptr.astype(np.int64))
    rows = np.nonzero((k >= lo) & (k <= hi))[0]
    if len(rows) == 0:
        return None
    reps = max(1, int(target_iv // max(1, k[rows].sum())))
    rows = np.tile(rows, reps)
    kk = k[rows]
    rowptr = np.zeros(len(rows) + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(kk)
    src = np.repeat(base.rowptr[rows].astype(np.int64) - rowptr[:-1], kk) + np.arange(rowptr[-1])
    out = PinnedCsr(len(rows), int(rowptr[-1]))
    out.rowptr[:] = rowptr.astype(np.uint32)
    out.length[:] = base.length[rows]
    out.iv[:] = base.iv[src]
    return out


def main():
    args = sys.argv[1:]
    c = 4
    if args and args[0] == "-c":
        c = int(args[1])
        args = args[2:]
    ranges = [tuple(int(x) for x in a.split("-")) for a in args] or [
        (1, 3), (4, 11), (12, 19), (20, 27), (28, 35), (36, 43), (44, 51), (52, 59), (60, 67), (68, 75), (76, 91), (92, 107),
        (108, 123), (124, 160), (161, 250), (251, 500)]
    base = yb.synth_csr(400000, 50)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    print("c = %d" % c)
    print("%9s %9s %11s %9s %10s %12s" % ("k range", "rows", "intervals", "us/step", "ns/row", "ps/interval"))
    for lo, hi in ranges:
        csr = subset(base, lo, hi, float(os.environ.get("SWEEP_INTERVALS", "24e6")))
        if csr is None:
            continue
        fm = yb.FullMemory(device=0)
        fm.bind_csr(csr)
        fm.upload()
        for _ in range(3):
            fm.compute_device(c, 0.4, stream.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        K = 10
        for _ in range(K):
            fm.compute_device(c, 0.4, stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / K
        print("%4d-%-4d %9d %11d %9.1f %10.2f %12.1f" % (lo, hi, csr.n_reads, csr.n_iv, us, us * 1e3 / csr.n_reads, us * 1e6 / csr.n_iv))
        fm.close()


if __name__ == "__main__":
    main()
