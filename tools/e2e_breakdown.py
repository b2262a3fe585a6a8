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
# the same as one call, one shot and streamed (yb_set_chunk_intervals): reads/s end to end
for chunk in (0, 4_000_000, 8_000_000, 16_000_000, 32_000_000):
    fm.set_chunk_intervals(chunk)
    best = 1e9
    for it in range(4):
        t0 = time.perf_counter(); fm.reset(); fm.bind_csr(csr); fm.compute_all(4, 0.4)
        best = min(best, time.perf_counter() - t0)
    print("compute_all, chunk_intervals %9d: %.2f ms = %.1f M reads/s, %.1f GB/s of input over PCIe"
          % (chunk, best * 1e3, n / best / 1e6, csr.nbytes / best / 1e9))
