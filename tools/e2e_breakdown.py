#!/usr/bin/env python
"""Wall-clock breakdown of the end-to-end call (bind -> upload -> kernels -> download) on one GPU."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yacrd_b200 as yb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
csr = yb.synth_csr(n, 50)
fm = yb.FullMemory(device=0)
for it in range(4):
    t0 = time.perf_counter(); fm.reset(); fm.bind_csr(csr)
    t1 = time.perf_counter(); fm.upload(); fm.synchronize()
    t2 = time.perf_counter(); fm.compute_device(4, 0.4); fm.synchronize()
    t3 = time.perf_counter(); fm.download()
    t4 = time.perf_counter()
    print("iter %d: reset+bind %.2f ms | freeze+H2D %.2f ms (%.1f GB/s if all copy) | kernels %.2f ms | D2H %.2f ms | total %.2f ms"
          % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, csr.nbytes / (t2 - t1) / 1e9, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t4 - t0) * 1e3))
