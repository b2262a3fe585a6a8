#!/usr/bin/env python
"""Small detect runs for compute-sanitizer (memcheck / racecheck / synccheck): every tier, both key widths.
usage: compute-sanitizer --tool racecheck python tools/sanitize_run.py"""
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yacrd_b200 as yb  # noqa: E402
from oracle import yacrd_oracle as o  # noqa: E402


def random_csr(rng, n_reads, ks, lens):
    rows, ll = [], []
    for _ in range(n_reads):
        length = rng.choice(lens)
        ivs = []
        for _ in range(rng.choice(ks)):
            b = rng.randrange(0, length)
            ivs.append((b, rng.randrange(b + 1, length + 1)))
        rows.append(ivs)
        ll.append(length)
    rowptr = np.zeros(n_reads + 1, dtype=np.uint32)
    rowptr[1:] = np.cumsum([len(r) for r in rows])
    iv = np.array([p for r in rows for p in r], dtype=np.uint32).reshape(-1, 2)
    return rowptr, iv, np.array(ll, dtype=np.uint32)


rng = random.Random(9)
rowptr, iv, length = random_csr(rng, 600, [0, 1, 5, 16, 17, 40, 59, 60, 64, 100, 123, 130, 200, 300, 513, 700, 1500],
                                [3, 50, 4000, 65534, 70000])
for rl in ("0", "128"):
    os.environ["YB_RL_MAX_SLOTS"] = rl
    for c in (0, 4):
        fm = yb.FullMemory(device=0)
        fm.add_csr(rowptr, iv, length)
        bp = yb.FromOverlap(fm, c, 0.4)
        bp.compute_all_bad_part()
        cls, gp, gaps = o.run_csr(rowptr, iv, length, c, 0.4)
        g_gp, g_gaps = bp.gap_csr()
        assert np.array_equal(bp.classes(), cls) and np.array_equal(g_gaps, gaps)
        fm.close()
print("sanitize_run ok")
