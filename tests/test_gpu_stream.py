"""Streamed batches (yb_set_chunk_intervals, the reference's -d/--ondisk-buffer-size): chunks through two lanes give
bit for bit what the one-shot call gives, whatever the chunk size. Needs a GPU."""
import random

import numpy as np
import pytest

import yacrd_b200 as yb
from oracle import yacrd_oracle as o
from tests.test_gpu_parity import _malformed_csr, _random_csr
from yacrd_b200 import dist as ybd

pytestmark = pytest.mark.gpu


def _streamed(rowptr, iv, length, c, n, chunk):
    fm = yb.FullMemory()
    fm.add_csr(rowptr, iv, length)
    fm.set_chunk_intervals(chunk)
    bp = yb.FromOverlap(fm, c, n)
    bp.compute_all_bad_part()
    gp, gaps = bp.gap_csr()
    out = (bp.classes().copy(), gp.copy(), gaps.copy(), bp.class_bitmap().copy(), fm.stats(),
           [bp.report_line(i) for i in (0, len(length) // 2, len(length) - 1)])
    fm.close()
    return out


@pytest.mark.parametrize("chunk", [1, 5000, 40000, 10**6, 10**9])
def test_streamed_equals_one_shot_and_oracle(chunk):
    rng = random.Random(4242)
    rowptr, iv, length, bad_rows = _malformed_csr(rng, 9000, [0, 1, 2, 5, 16, 17, 40, 64, 65, 100, 200, 513, 700, 3000],
                                                  [1, 50, 3000, 65534, 65535, 250000], 0.02)
    cls, gp, gaps, bm, st, lines = _streamed(rowptr, iv, length, 3, 0.4, chunk)
    w_cls, w_gp, w_gaps = o.run_csr(rowptr, iv, length, 3, 0.4)
    assert np.array_equal(gp.astype(np.uint64), w_gp) and np.array_equal(gaps, w_gaps) and np.array_equal(cls, w_cls)
    assert np.array_equal(ybd.unpack_bitmap(bm, len(cls)), w_cls)
    one = _streamed(rowptr, iv, length, 3, 0.4, 0)
    assert lines == one[5]
    for k in ("n_reads", "n_intervals", "n_gaps", "n_not_bad", "n_chimeric", "n_not_covered", "max_intervals_per_read",
              "n_malformed_intervals", "n_literal_reads"):
        assert st[k] == one[4][k], k
    assert st["n_not_bad"] + st["n_chimeric"] + st["n_not_covered"] == len(cls)
    assert 0 < st["n_literal_reads"] <= bad_rows


def test_streamed_context_is_reusable_and_bound_csr_streams_too():
    rng = random.Random(7)
    fm = yb.FullMemory()
    fm.set_chunk_intervals(20000)
    for n_reads in (5000, 1500, 7000):
        rowptr, iv, length = _random_csr(rng, n_reads, [0, 3, 30, 90, 300], [100, 9000, 70000])
        csr = yb.PinnedCsr(n_reads, len(iv))
        csr.rowptr[:] = rowptr
        csr.iv[:] = iv
        csr.length[:] = length
        fm.reset()
        fm.bind_csr(csr)
        fm.compute_all(2, 0.8)
        w_cls, w_gp, w_gaps = o.run_csr(rowptr, iv, length, 2, 0.8)
        assert np.array_equal(fm.classes(), w_cls)
        assert np.array_equal(fm.gap_ptr().astype(np.uint64), w_gp) and np.array_equal(fm.gaps(), w_gaps)
        csr.free()
    fm.close()
