#!/usr/bin/env python
"""Ingestion throughput: synthetic PAF text -> interned reads + CSR, host only.
usage: python tools/bench_ingest.py [n_records] [n_reads] [threads ...]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import yacrd_b200 as yb
from yacrd_b200 import _native as N

n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else max(1000, n_rec // 25)
threads = [int(x) for x in sys.argv[3:]] or [1, 2, 4, 8, 0]
import workload
L = N.lib()
W = workload.lib()
need = W.yb_synth_paf(20261017, n_reads, n_rec, None, 0)
buf = np.empty(need, dtype=np.uint8)
t0 = time.perf_counter()
nb = W.yb_synth_paf(20261017, n_reads, n_rec, buf.ctypes.data, need)
print("synthetic PAF: %d records, %d reads, %.1f MB (generated in %.2f s), %d cores" % (n_rec, n_reads, nb / 1e6, time.perf_counter() - t0, os.cpu_count()))
ref = None
for th in threads:
    fm = yb.FullMemory(host_only=True, ingest_threads=th)
    t0 = time.perf_counter()
    fm._ck(L.yb_init_buffer(fm._h, C.cast(buf.ctypes.data, C.c_char_p), nb, ord("p")))
    dt = time.perf_counter() - t0
    sig = (fm.n_reads(), fm.read_at(0), fm.read_at(fm.n_reads() - 1), tuple(fm.overlap(fm.read_at(17))[:3]))
    ref = ref or sig
    assert sig == ref
    print("threads %2s: %6.2f s  %7.1f MB/s  %6.2f M records/s  (%d reads)" % (th or "all", dt, nb / dt / 1e6, n_rec / dt / 1e6, fm.n_reads()))
    fm.close()
