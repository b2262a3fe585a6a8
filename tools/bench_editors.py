#!/usr/bin/env python
"""Throughput of the post-detection editors (run on the B200 box): a synthetic fastq whose reads are the reads of a
synthetic CSR (names = read indices, lengths = the CSR's), edited with the device results of that CSR.
usage: python tools/bench_editors.py [n_reads]"""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yacrd_b200 as yb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
csr = yb.synth_csr(n, 50)
fm = yb.FullMemory(device=0)
fm.bind_csr(csr)
bp = yb.FromOverlap(fm, 4, 0.4)
bp.compute_all_bad_part()
rng = np.random.default_rng(1)
tmp = tempfile.mkdtemp()
src = os.path.join(tmp, "reads.fastq")
t0 = time.perf_counter()
with open(src, "wb") as fh:
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for r in range(n):
        ln = int(csr.length[r])
        seq = acgt[rng.integers(0, 4, ln)].tobytes()
        fh.write(b"@%d synthetic length=%d\n" % (r, ln) + seq + b"\n+\n" + b"?" * ln + b"\n")
size = os.path.getsize(src)
print("fastq: %d reads, %.1f MB (written in %.1f s)" % (n, size / 1e6, time.perf_counter() - t0))
for name, fn in (("filter", yb.filter), ("extract", yb.extract), ("split", yb.split), ("scrubb", yb.scrubbing)):
    out = os.path.join(tmp, "out.%s.fastq" % name)
    t = time.perf_counter()
    fn(src, out, bp, 0.4)
    dt = time.perf_counter() - t
    print("%-8s %.3f s  %7.1f MB/s in  -> %.1f MB out" % (name, dt, size / 1e6 / dt, os.path.getsize(out) / 1e6))
    os.remove(out)
os.remove(src)
