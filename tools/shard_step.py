#!/usr/bin/env python
"""One rank's detect step of an n-way sharded config 3 on a single GPU (for ncu launch lists of the N > 1 shapes).
usage: python tools/shard_step.py [n_shards] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yacrd_b200 as yb  # noqa: E402

n_shards = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
csr = yb.synth_csr(2_000_000, 50, shard=0, n_shards=n_shards)
fm = yb.FullMemory(device=0)
fm.bind_csr(csr)
fm.upload()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
for _ in range(3):
    fm.compute_device(4, 0.4, stream.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(steps):
    fm.compute_device(4, 0.4, stream.cuda_stream)
e1.record(stream)
torch.cuda.synchronize()
print("shard 0 of %d: %d reads, %d intervals, %.1f us/step (eager launches)" % (n_shards, csr.n_reads, csr.n_iv, e0.elapsed_time(e1) * 1e3 / steps))
