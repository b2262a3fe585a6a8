"""Pins the CPU oracle (oracle/) to the reference's own vectors. CPU only."""
import os
import random

import numpy as np
import pytest

from oracle import yacrd_oracle as o
from tests import kats
from tests.conftest import GOLDEN, read_sorted_lines


@pytest.mark.parametrize("name,ivs,length,cov,expect", kats.STACK_KATS, ids=[k[0] for k in kats.STACK_KATS])
def test_stack_kats(name, ivs, length, cov, expect):
    assert o.compute_bad_part(ivs, length, cov) == expect
    assert o.c_compute_bad_part(ivs, length, cov) == expect


@pytest.mark.parametrize("bads,length,n,expect", kats.TYPE_KATS)
def test_type_of_read_kats(bads, length, n, expect):
    assert o.type_of_read(length, bads, n) == expect
    assert o.c_type_of_read(length, bads, n) == expect


@pytest.mark.parametrize("rid,length,bads,n,line", kats.REPORT_KATS, ids=[k[0] for k in kats.REPORT_KATS])
def test_report_lines(rid, length, bads, n, line):
    assert o.report_line(rid, length, bads, n) == line
    assert o.c_report_line(rid, length, bads, n) == line


@pytest.mark.parametrize("name,ivs,length,cov,gaps,cls", kats.QUIRK_KATS, ids=[k[0] for k in kats.QUIRK_KATS])
def test_quirks(name, ivs, length, cov, gaps, cls):
    for f in (o.compute_bad_part, o.c_compute_bad_part):
        assert f(ivs, length, cov) == gaps
    assert o.type_of_read(length, gaps, 0.8) == cls
    assert o.c_type_of_read(length, gaps, 0.8) == cls


def test_unknown_read_is_not_bad():
    # stack.rs:164-169 returns (vec![], 0) for an unknown id => 0/0 = NaN => NotBad (editor/mod.rs:88)
    assert o.type_of_read(0, [], 0.8) == o.NOT_BAD
    assert o.c_type_of_read(0, [], 0.8) == o.NOT_BAD


def test_golden_paf_matches_reference_truth():
    """tests/run.rs:96-117: reads.paf, default flags (-c 0 -n 0.8) == truth.yacrd as a line set."""
    reads = o.ingest_paf(os.path.join(GOLDEN, "c1_overlaps.paf"))
    assert len(reads) == 230 and sum(len(v[0]) for v in reads.values()) == 2572
    truth = read_sorted_lines(os.path.join(GOLDEN, "c1_truth.sorted.yacrd"))
    assert sorted(o.detect_lines(reads, 0, 0.8)) == truth
    # same through the C oracle's batch driver + line formatter
    ids = list(reads)
    rowptr = np.zeros(len(ids) + 1, dtype=np.uint64)
    rowptr[1:] = np.cumsum([len(reads[i][0]) for i in ids])
    iv = np.array([p for i in ids for p in reads[i][0]], dtype=np.uint32)
    length = np.array([reads[i][1] for i in ids], dtype=np.uint32)
    for threads in (1, 3):
        cls, gap_ptr, gaps = o.run_csr(rowptr, iv, length, 0, 0.8, threads=threads)
        lines = []
        for r, rid in enumerate(ids):
            g = [tuple(map(int, x)) for x in gaps[int(gap_ptr[r]):int(gap_ptr[r + 1])]]
            line = o.c_report_line(rid, int(length[r]), g, 0.8)
            assert line.split("\t")[0] == o.TYPE_NAMES[cls[r]]
            lines.append(line)
        assert sorted(lines) == truth


def test_golden_m4_matches_reference_truth():
    reads = o.ingest_m4(os.path.join(GOLDEN, "c1_overlaps.m4"))
    truth = read_sorted_lines(os.path.join(GOLDEN, "c1_truth.sorted.yacrd"))
    assert sorted(o.detect_lines(reads, 0, 0.8)) == truth


@pytest.mark.parametrize("c,n", [(4, 0.4), (3, 0.4), (1, 0.8)])
def test_oracle_derived_presets(c, n):
    reads = o.ingest_paf(os.path.join(GOLDEN, "c1_overlaps.paf"))
    want = read_sorted_lines(os.path.join(GOLDEN, "c1_oracle_c%d_n%s.sorted.yacrd" % (c, n)))
    assert sorted(o.detect_lines(reads, c, n)) == want


def test_reads2ovl_kat():
    """reads2ovl/mod.rs:170-237: two-record PAF and M4 give the same three reads."""
    import tempfile
    paf = ("1\t12000\t20\t4500\t-\t2\t10000\t5500\t10000\t4500\t4500\t255\n"
           "1\t12000\t5500\t10000\t-\t3\t10000\t0\t4500\t4500\t4500\t255\n")
    m4 = ("1 2 0.1 2 0 20 4500 12000 0 5500 10000 10000\n"
          "1 3 0.1 2 0 5500 10000 12000 0 0 4500 10000\n")
    for text, suffix, fn in ((paf, ".paf", o.ingest_paf), (m4, ".m4", o.ingest_m4)):
        with tempfile.NamedTemporaryFile("w", suffix=suffix, delete=False) as fh:
            fh.write(text)
        try:
            reads = fn(fh.name)
        finally:
            os.unlink(fh.name)
        assert set(reads) == {"1", "2", "3"}
        assert reads["1"][0] == [(20, 4500), (5500, 10000)]
        assert reads["2"][0] == [(5500, 10000)]
        assert reads["3"][0] == [(0, 4500)]
        assert reads["1"][1] == 12000 and reads["2"][1] == 10000


def test_python_and_c_restatements_agree_fuzz():
    rng = random.Random(20261017)
    for _ in range(3000):
        length = rng.choice([1, 2, 5, 20, 100, 1000, 70000])
        k = rng.choice([0, 1, 2, 3, 5, 9, 17, 40])
        ivs = []
        for _ in range(k):
            b = rng.randrange(0, length)
            e = rng.randrange(b + 1, length + 1)
            ivs.append((b, e))
        c = rng.choice([0, 0, 1, 2, 3, 4, 7, 50])
        a = o.compute_bad_part(ivs, length, c)
        assert a == o.c_compute_bad_part(ivs, length, c)
        for n in (0.0, 0.4, 0.8, 1.0):
            assert o.type_of_read(length, a, n) == o.c_type_of_read(length, a, n)
