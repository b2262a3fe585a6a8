#!/usr/bin/env python
"""Whole-pipeline wall clock of the driver on a synthetic PAF file: read + parse -> CSR -> H2D -> kernels -> D2H -> report.
usage: python tools/bench_cli.py [n_records] [n_reads] [-c C] [-n N]"""
import ctypes as C, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from yacrd_b200 import _native as N

n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else n_rec // 25
import workload
L = N.lib()
W = workload.lib()
need = W.yb_synth_paf(20261017, n_reads, n_rec, None, 0)
buf = np.empty(need, dtype=np.uint8)
nb = W.yb_synth_paf(20261017, n_reads, n_rec, buf.ctypes.data, need)
path = "/tmp/yb_synth_%d.paf" % n_rec
buf[:nb].tofile(path)
del buf
print("PAF: %d records, %d reads, %.1f MB -> %s" % (n_rec, n_reads, nb / 1e6, path))
cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "yacrd_b200", "yacrd-b200")
for threads in ("1", "0"):
    t0 = time.perf_counter()
    r = subprocess.run([cli, "-i", path, "-o", "/tmp/yb_out.yacrd", "-c", "4", "-n", "0.4", "-t", threads, "--timing"], capture_output=True, text=True)
    dt = time.perf_counter() - t0
    print("-t %s: wall %.2f s (%.2f M records/s)  rc=%d  %s" % (threads, dt, n_rec / dt / 1e6, r.returncode, r.stderr.strip()[-400:]))
os.remove(path)
