"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the reference's golden vectors.
Integer/index work: the bar is bit-exact. Needs a GPU: run with -m gpu on the B200 box."""
import os
import random

import numpy as np
import pytest

import yacrd_b200 as yb
from oracle import yacrd_oracle as o
from tests import kats
from tests.conftest import GOLDEN, read_sorted_lines
from yacrd_b200 import dist as ybd

pytestmark = pytest.mark.gpu


def run_gpu_csr(rowptr, iv, length, c, n):
    fm = yb.FullMemory()
    fm.add_csr(rowptr, iv, length)
    bp = yb.FromOverlap(fm, c, n)
    bp.compute_all_bad_part()
    gp, gaps = bp.gap_csr()
    out = (bp.classes().copy(), gp.copy(), gaps.copy(), bp.class_bitmap().copy())
    fm.close()
    return out


def assert_same_as_oracle(rowptr, iv, length, c, n, threads=0):
    cls, gp, gaps, bm = run_gpu_csr(rowptr, iv, length, c, n)
    w_cls, w_gp, w_gaps = o.run_csr(rowptr, iv, length, c, n, threads=threads)
    assert np.array_equal(gp.astype(np.uint64), w_gp), "gap_ptr differs (first at read %d)" % int(
        np.nonzero(gp.astype(np.uint64) != w_gp)[0][0] - 1)
    assert np.array_equal(gaps, w_gaps)
    assert np.array_equal(cls, w_cls)
    assert np.array_equal(ybd.unpack_bitmap(bm, len(cls)), w_cls)
    return cls, gp, gaps


def _one_read(ivs, length, c, n=0.8):
    fm = yb.FullMemory()
    if ivs:
        for iv in ivs:
            fm.add_overlap_and_length("r", iv, length)
    else:
        fm.add_length("r", length)
    bp = yb.FromOverlap(fm, c, n)
    bp.compute_all_bad_part()
    gaps, ln = bp.get_bad_part("r")
    t = bp.type_of_read("r")
    line = bp.report_line(0)
    fm.close()
    return gaps, ln, t, line


@pytest.mark.parametrize("name,ivs,length,cov,expect", kats.STACK_KATS, ids=[k[0] for k in kats.STACK_KATS])
def test_stack_kats(name, ivs, length, cov, expect):
    gaps, ln, _, _ = _one_read(ivs, length, cov)
    assert gaps == expect and ln == length


def test_type_of_read_kats():
    # editor/mod.rs:114-128 pins types for the stack.rs:312-333 reads A..F at n = 0.8
    for (name, ivs, length, cov, expect), (bads, _, n, want) in zip(kats.STACK_KATS[:6], kats.TYPE_KATS):
        assert expect == bads
        gaps, _, t, line = _one_read(ivs, length, cov, n)
        assert int(t) == want
        assert line == o.report_line("r", length, gaps, n)


@pytest.mark.parametrize("name,ivs,length,cov,gaps,cls", kats.QUIRK_KATS, ids=[k[0] for k in kats.QUIRK_KATS])
def test_quirks(name, ivs, length, cov, gaps, cls):
    got, ln, t, _ = _one_read(ivs, length, cov)
    assert got == gaps and int(t) == cls


def test_unknown_read_is_empty_and_not_bad():
    fm = yb.FullMemory()
    fm.add_overlap_and_length("a", (10, 90), 100)
    bp = yb.FromOverlap(fm, 0)
    bp.compute_all_bad_part()
    assert bp.get_bad_part("nope") == ([], 0)  # stack.rs:164-169
    assert bp.type_of_read("nope") == yb.ReadType.NotBad
    assert fm.length("nope") == 0 and fm.overlap("nope") == []
    fm.close()


def test_store_semantics_first_seen_length_and_duplicates():
    fm = yb.FullMemory()
    fm.add_overlap_and_length("a", (10, 90), 100)
    fm.add_overlap_and_length("a", (10, 90), 5000)  # later length ignored (fullmemory.rs:82-90)
    fm.add_overlap_and_length("b", (0, 50), 60)
    assert fm.length("a") == 100 and fm.overlap("a") == [(10, 90), (10, 90)]
    assert fm.read_ids() == ["a", "b"] and fm.get_reads() == {"a", "b"}
    fm.add_length("b", 70)  # add_length sets unconditionally (fullmemory.rs:78-80)
    assert fm.length("b") == 70
    bp = yb.FromOverlap(fm, 1)
    bp.compute_all_bad_part()
    assert bp.get_bad_part("a") == ([(0, 10), (90, 100)], 100)  # duplicates both count toward depth
    assert bp.get_bad_part("b") == ([(0, 70)], 70)
    fm.close()


@pytest.mark.parametrize("fname,truth,c,n", [
    ("c1_overlaps.paf", "c1_truth.sorted.yacrd", 0, 0.8),          # tests/run.rs:96-117
    ("c1_overlaps.m4", "c1_truth.sorted.yacrd", 0, 0.8),
    ("c1_overlaps.paf", "c1_oracle_c4_n0.4.sorted.yacrd", 4, 0.4),  # oracle-derived presets
    ("c1_overlaps.paf", "c1_oracle_c3_n0.4.sorted.yacrd", 3, 0.4),
    ("c1_overlaps.paf", "c1_oracle_c1_n0.8.sorted.yacrd", 1, 0.8),
])
def test_golden_report(tmp_path, fname, truth, c, n):
    fm = yb.FullMemory()
    fm.init(os.path.join(GOLDEN, fname))
    assert fm.n_reads() == 230
    bp = yb.FromOverlap(fm, c, n)
    bp.compute_all_bad_part()
    out = str(tmp_path / "out.yacrd")
    bp.write_report(out)
    want = read_sorted_lines(os.path.join(GOLDEN, truth))
    assert read_sorted_lines(out) == want  # byte-identical after sort (order is not part of the contract)
    assert sorted(bp.report_lines()) == want
    # FromReport round trip (stack.rs:176-257): same lines, same classes, no pile-up
    rp = yb.FromReport(out, not_coverage=n)
    rp.compute_all_bad_part()
    assert sorted(rp.report_lines()) == want
    assert np.array_equal(rp.classes(), bp.classes())
    rid = fm.read_at(7)
    assert rp.get_bad_part(rid) == bp.get_bad_part(rid)
    rp.ctx.close()
    fm.close()


def test_reads2ovl_kat():
    """reads2ovl/mod.rs:170-237 through the device library's ingestion."""
    paf = ("1\t12000\t20\t4500\t-\t2\t10000\t5500\t10000\t4500\t4500\t255\n"
           "1\t12000\t5500\t10000\t-\t3\t10000\t0\t4500\t4500\t4500\t255\n")
    m4 = ("1 2 0.1 2 0 20 4500 12000 0 5500 10000 10000\n"
          "1 3 0.1 2 0 5500 10000 12000 0 0 4500 10000\n")
    for text, fmt in ((paf, "paf"), (m4, "m4")):
        fm = yb.FullMemory()
        fm.init_buffer(text, fmt)
        assert fm.get_reads() == {"1", "2", "3"}
        assert fm.overlap("1") == [(20, 4500), (5500, 10000)]
        assert fm.overlap("2") == [(5500, 10000)] and fm.overlap("3") == [(0, 4500)]
        assert fm.length("1") == 12000 and fm.length("2") == 10000
        fm.close()


def test_bad_record_is_reading_error():
    fm = yb.FullMemory()
    with pytest.raises(yb.YacrdError) as e:
        fm.init_buffer("1\t12000\tx\t4500\t-\t2\t10000\t5500\t10000\n", "paf")
    assert e.value.kind == "ReadingError"
    fm.close()
    with pytest.raises(yb.YacrdError) as e:
        yb.FullMemory().init("/nonexistent/x.paf")
    assert e.value.kind == "CantReadFile"
    with pytest.raises(yb.YacrdError) as e:
        yb.FullMemory().init("reads.fastq")
    assert e.value.kind == "CantRunOperationOnFile"
    with pytest.raises(yb.YacrdError) as e:
        yb.FullMemory().init("reads.txt")
    assert e.value.kind == "UnableToDetectFileFormat"


def test_corrupt_report_is_an_error():
    # stack.rs:393-409
    with pytest.raises(yb.YacrdError) as e:
        yb.FromReport(text="NotBad\tSRR8494940.65223\t2706\t1131,0,1131;16,2690,2706\n"
                           "NotCovered\tSRR8494940.141626\t30116\t326,0,326;27159,2957\n")
    assert e.value.kind == "CorruptYacrdReport"


def test_failed_report_load_leaves_an_empty_usable_context():
    fm = yb.FullMemory()
    bad = b"NotBad\tgood\t100\t\nNotCovered\tbroken\t30116\t326,0,326;27159,2957\n"
    assert fm._L.yb_init_report_buffer(fm._h, bad, len(bad)) == -7
    assert fm.n_reads() == 0 and fm.length("good") == 0
    huge = b"NotBad\tr\t4294967296\t\n"  # the classifier holds lengths in 32 bits
    assert fm._L.yb_init_report_buffer(fm._h, huge, len(huge)) == -11
    assert fm.n_reads() == 0
    fm.add_overlap_and_length("good", (10, 90), 100)  # the same context takes a detect batch afterwards
    bp = yb.FromOverlap(fm, 0)
    bp.compute_all_bad_part()
    assert bp.get_bad_part("good") == ([(0, 10), (90, 100)], 100)
    fm.close()


def test_malformed_intervals_follow_the_reference_kats():
    """begin >= end or end > length: the reference has no such test and its heap sweep gives a definite answer
    (stack.rs:61-139); those reads take literal_kernel. One read each, against the oracle's literal sweep."""
    for ivs, ln in (([(50, 50)], 100), ([(60, 40)], 100), ([(10, 120)], 100), ([(0, 30), (20, 10), (25, 200)], 100),
                    ([(5, 5), (5, 5), (90, 100)], 100)):
        for c in (0, 1):
            gaps, length, _, line = _one_read(ivs, ln, c)
            want = o.compute_bad_part(ivs, ln, c)
            assert gaps == want and length == ln, (ivs, ln, c, gaps, want)


def _random_csr(rng, n_reads, k_choices, len_choices):
    rows, lens = [], []
    for _ in range(n_reads):
        length = rng.choice(len_choices)
        k = rng.choice(k_choices)
        ivs = []
        for _ in range(k):
            if rng.random() < 0.3:
                b = rng.randrange(0, min(length, 4))
            else:
                b = rng.randrange(0, length)
            e = length if rng.random() < 0.25 else rng.randrange(b + 1, length + 1)
            ivs.append((b, e))
        rows.append(ivs)
        lens.append(length)
    rowptr = np.zeros(n_reads + 1, dtype=np.uint32)
    rowptr[1:] = np.cumsum([len(r) for r in rows])
    iv = np.array([p for r in rows for p in r], dtype=np.uint32).reshape(-1, 2)
    return rowptr, iv, np.array(lens, dtype=np.uint32)


@pytest.mark.parametrize("c", [0, 1, 2, 4, 9, 100])
def test_fuzz_small_reads_every_register_tier(c):
    rng = random.Random(1000 + c)
    rowptr, iv, length = _random_csr(rng, 6000, [0, 1, 2, 3, 4, 7, 8, 9, 15, 16, 17, 31, 32, 33, 50, 63, 64, 65, 100,
                                                 127, 128, 129, 160, 161, 200, 255, 256, 257, 400, 512],
                                     [1, 2, 3, 8, 20, 64, 1000, 65533, 65534, 65535, 65536, 250000, 2**31 - 1])
    for n in (0.4, 0.8):
        assert_same_as_oracle(rowptr, iv, length, c, n)


@pytest.mark.parametrize("c", [0, 1, 3, 4, 7, 8, 15, 16, 17, 30, 63, 64, 127, 5000, 2**32 - 1])
def test_every_k_up_to_530_and_every_class_boundary(c):
    """Rows of every k in 0..530 (every size class of the register tier: 16, 32, 48, 64, 80, 96, 128, 160, 256, 512 key
    slots, both ends of each, and the hand-over to the CTA tier), short and 16-bit-limit lengths, behind random rows."""
    rng = random.Random(4242 + c % 1000)
    rowptr, iv, length = _random_csr(rng, 141 * 40, list(range(141)), [1, 5, 40, 3000, 65534, 65535])
    rows = [[(rng.randrange(0, 900), 1000 - rng.randrange(0, 90)) for _ in range(k)] for k in range(531)]
    rp2 = np.zeros(532, dtype=np.uint32)
    rp2[1:] = np.cumsum([len(r) for r in rows])
    iv2 = np.array([p for r in rows for p in r], dtype=np.uint32).reshape(-1, 2)
    rowptr = np.concatenate([rowptr, rowptr[-1] + rp2[1:]]).astype(np.uint32)
    iv = np.concatenate([iv, iv2])
    length = np.concatenate([length, np.full(531, 1000)]).astype(np.uint32)
    assert_same_as_oracle(rowptr, iv, length, c, 0.4)


def _malformed_csr(rng, n_reads, k_choices, len_choices, p_bad_row):
    rowptr, iv, length = _random_csr(rng, n_reads, k_choices, len_choices)
    iv = iv.copy()
    bad_rows = 0
    for r in range(n_reads):
        k = int(rowptr[r + 1] - rowptr[r])
        if k == 0 or rng.random() >= p_bad_row:
            continue
        bad_rows += 1
        for _ in range(rng.choice([1, 1, 2, k])):
            i = int(rowptr[r]) + rng.randrange(k)
            kind = rng.randrange(4)
            ln = int(length[r])
            if kind == 0:
                iv[i] = (iv[i][0], iv[i][0])                       # empty
            elif kind == 1:
                iv[i] = (iv[i][1], iv[i][0])                       # reversed
            elif kind == 2:
                iv[i] = (iv[i][0], ln + rng.randrange(1, 70000))   # beyond the read's end
            else:
                iv[i] = (rng.randrange(0, 2**32), rng.randrange(0, 2**32))
    return rowptr, iv, length, bad_rows


@pytest.mark.parametrize("c", [0, 2, 4])
def test_fuzz_malformed_rows_take_the_literal_sweep(c):
    """3 % of the rows (every size class, the CTA tier included) get empty / reversed / out-of-range / random intervals:
    the whole batch still equals the oracle's literal heap sweep, and the stats say how many rows took literal_kernel."""
    rng = random.Random(99 + c)
    rowptr, iv, length, bad_rows = _malformed_csr(rng, 8000, [0, 1, 2, 5, 16, 17, 40, 64, 65, 100, 200, 513, 700],
                                                  [1, 50, 3000, 65534, 65535, 250000], 0.03)
    fm = yb.FullMemory()
    fm.add_csr(rowptr, iv, length)
    bp = yb.FromOverlap(fm, c, 0.4)
    bp.compute_all_bad_part()
    gp, gaps = bp.gap_csr()
    w_cls, w_gp, w_gaps = o.run_csr(rowptr, iv, length, c, 0.4)
    assert np.array_equal(gp.astype(np.uint64), w_gp) and np.array_equal(gaps, w_gaps) and np.array_equal(bp.classes(), w_cls)
    st = fm.stats()
    assert 0 < st["n_literal_reads"] <= bad_rows and st["n_malformed_intervals"] >= st["n_literal_reads"]
    fm.close()


def test_all_rows_malformed():
    rng = random.Random(5150)
    rowptr, iv, length, _ = _malformed_csr(rng, 3000, [1, 3, 20, 70, 600], [100, 70000], 1.0)
    assert_same_as_oracle(rowptr, iv, length, 1, 0.4)


def test_fuzz_tiny_positions_many_ties():
    rng = random.Random(5)
    rowptr, iv, length = _random_csr(rng, 20000, [1, 2, 3, 5, 8, 13, 21, 40], [1, 2, 3, 4, 5, 6, 9])
    for c in (0, 1, 3):
        assert_same_as_oracle(rowptr, iv, length, c, 0.5)


def test_fuzz_cta_tier_and_mixed_sizes():
    rng = random.Random(77)
    rowptr, iv, length = _random_csr(rng, 300, [0, 3, 257, 300, 511, 512, 513, 1000, 2047, 2048, 2049, 5000, 8192],
                                     [50, 3000, 250000])
    for c in (0, 3, 40):
        assert_same_as_oracle(rowptr, iv, length, c, 0.4)


def test_huge_reads_global_scratch_tier():
    rng = random.Random(3)
    rowptr, iv, length = _random_csr(rng, 6, [2, 8193, 20000, 40000], [100000, 2**31 - 1])
    for c in (0, 7):
        assert_same_as_oracle(rowptr, iv, length, c, 0.4)


def test_empty_context_and_reads_without_intervals():
    cls, gp, gaps, bm = run_gpu_csr(np.zeros(1, np.uint32), np.zeros((0, 2), np.uint32), np.zeros(0, np.uint32), 0, 0.8)
    assert len(cls) == 0 and list(gp) == [0] and len(gaps) == 0
    rowptr = np.array([0, 0, 0, 1, 1], dtype=np.uint32)
    iv = np.array([[5, 9]], dtype=np.uint32)
    length = np.array([100, 0, 10, 7], dtype=np.uint32)
    assert_same_as_oracle(rowptr, iv, length, 0, 0.8)


def test_config2_100k_reads_c0():
    """BASELINE config 2: synthetic 100k reads x mean 30 overlaps, ONT lengths, -c 0 (-n 0.8)."""
    csr = yb.synth_csr(100000, 30)
    cls, gp, gaps = assert_same_as_oracle(csr.rowptr, csr.iv, csr.length, 0, 0.8)
    assert np.bincount(cls, minlength=3).min() > 0


def test_config5_like_skewed_pacbio():
    """BASELINE config 5 at 1/10 size on one GPU: PacBio lengths, Pareto k capped at 5000, -c 3 -n 0.4."""
    csr = yb.synth_csr(50000, 0, profile=yb.SYNTH_PACBIO_SKEW)
    assert_same_as_oracle(csr.rowptr, csr.iv, csr.length, 3, 0.4)


def test_config5_full_size_500k_reads():
    """BASELINE config 5 at full size (500 k reads, PacBio Sequel lengths, Pareto k capped at 5000, -c 3 -n 0.4): every
    tier runs (register tier, position scan, CTA sort). Full comparison with the oracle, one shot and streamed."""
    csr = yb.synth_csr(500_000, 0, profile=yb.SYNTH_PACBIO_SKEW)
    w_cls, w_gp, w_gaps = o.run_csr(csr.rowptr, csr.iv, csr.length, 3, 0.4)
    fm = yb.FullMemory()
    for chunk in (0, 3_000_000):
        fm.reset()
        fm.set_chunk_intervals(chunk)
        fm.bind_csr(csr)
        bp = yb.FromOverlap(fm, 3, 0.4)
        bp.compute_all_bad_part()
        gp, gaps = bp.gap_csr()
        assert np.array_equal(gp.astype(np.uint64), w_gp) and np.array_equal(gaps, w_gaps) and np.array_equal(bp.classes(), w_cls)
        assert np.array_equal(ybd.unpack_bitmap(bp.class_bitmap(), len(w_cls)), w_cls)
        st = fm.stats()
        assert st["max_intervals_per_read"] > 2048 and st["n_gaps"] == len(w_gaps)
    fm.close()
    csr.free()


def test_config3_full_size_2m_reads():
    """BASELINE config 3 at full size (2 M reads x mean 50, -c 4 -n 0.4): full comparison with the oracle
    (multi-threaded C) plus size-independent properties."""
    csr = yb.synth_csr(2_000_000, 50)
    fm = yb.FullMemory()
    fm.bind_csr(csr)
    bp = yb.FromOverlap(fm, 4, 0.4)
    bp.compute_all_bad_part()
    cls, (gp, gaps) = bp.classes().copy(), bp.gap_csr()
    gp, gaps = gp.copy(), gaps.copy()
    st = fm.stats()
    # properties: monotone offsets, gaps sorted and disjoint inside a read, inside [0, len]
    assert gp[0] == 0 and (np.diff(gp.astype(np.int64)) >= 0).all() and gp[-1] == len(gaps) == st["n_gaps"]
    assert (gaps[:, 0] <= gaps[:, 1]).all()
    cnt = np.diff(gp.astype(np.int64))
    lens = np.repeat(csr.length, cnt)
    assert (gaps[:, 1] <= lens).all()
    inner = np.ones(len(gaps), dtype=bool)
    inner[gp[:-1][cnt > 0]] = False
    assert (gaps[1:, 0][inner[1:]] >= gaps[:-1, 1][inner[1:]]).all()
    assert st["n_not_bad"] + st["n_chimeric"] + st["n_not_covered"] == 2_000_000
    assert list(np.bincount(cls, minlength=3)) == [st["n_not_bad"], st["n_chimeric"], st["n_not_covered"]]
    # idempotence: a second pass over the same resident CSR gives the same bits
    bp.compute_all_bad_part()
    assert np.array_equal(cls, bp.classes()) and np.array_equal(gaps, bp.gap_csr()[1])
    # full comparison
    w_cls, w_gp, w_gaps = o.run_csr(csr.rowptr, csr.iv, csr.length, 4, 0.4)
    assert np.array_equal(gp.astype(np.uint64), w_gp) and np.array_equal(gaps, w_gaps) and np.array_equal(cls, w_cls)
    fm.close()


def test_sharded_equals_unsharded():
    """Config 4's layout on one GPU: per-shard kernels + bitmap slots reassemble to the unsharded answer."""
    n, G, c, nn = 40000, 4, 4, 0.4
    full = yb.synth_csr(n, 50)
    want, _, _ = o.run_csr(full.rowptr, full.iv, full.length, c, nn)
    _, _, counts = ybd.shard_layout(n, G)
    slot = ybd.bitmap_bytes(int(counts.max()))
    gathered = np.zeros((G, slot), dtype=np.uint8)
    for g in range(G):
        part = yb.synth_csr(n, 50, shard=g, n_shards=G)
        fm = yb.FullMemory()
        fm.bind_csr(part)
        bp = yb.FromOverlap(fm, c, nn)
        bp.compute_all_bad_part()
        bm = bp.class_bitmap()
        gathered[g, : len(bm)] = bm
        fm.close()
    assert np.array_equal(ybd.global_classes(gathered, n, G), want)


def test_lazy_device_context():
    """YB_FLAG_LAZY_DEVICE: yb_create does not touch CUDA; the first call that needs the device opens it (the driver
    parses its input while yb_device_warmup starts CUDA on another thread)."""
    from yacrd_b200 import _native as N
    assert N.lib().yb_device_warmup(-1) == N.OK
    fm = yb.FullMemory(lazy_device=True)
    fm.init(os.path.join(GOLDEN, "c1_overlaps.paf"))
    assert fm.n_reads() == 230
    bp = yb.FromOverlap(fm, 0, 0.8)
    bp.compute_all_bad_part()
    assert sorted(bp.report_lines()) == read_sorted_lines(os.path.join(GOLDEN, "c1_truth.sorted.yacrd"))
    fm.close()
    yb.FullMemory(lazy_device=True).close()  # never opened its device: nothing to release
    rp = yb.FromReport(os.path.join(GOLDEN, "c1_truth.sorted.yacrd"))
    rp.compute_all_bad_part()
    rp.ctx.close()


def test_reset_reuses_the_context_for_the_next_batch():
    fm = yb.FullMemory()
    for seed in (1, 2):
        csr = yb.synth_csr(3000, 20, seed=seed)
        fm.reset()
        fm.bind_csr(csr)
        bp = yb.FromOverlap(fm, 2, 0.4)
        bp.compute_all_bad_part()
        w_cls, w_gp, w_gaps = o.run_csr(csr.rowptr, csr.iv, csr.length, 2, 0.4)
        assert np.array_equal(bp.classes(), w_cls) and np.array_equal(bp.gap_csr()[1], w_gaps)
    fm.close()


def test_from_report_kats():
    """stack.rs:262-310 (report parsed back), :393-409 (corrupt), :411-431 (perfect read)."""
    text = ("NotBad\tSRR8494940.65223\t2706\t1131,0,1131;16,2690,2706\n"
            "NotCovered\tSRR8494940.141626\t30116\t326,0,326;27159,2957,30116\n"
            "Chimeric\tSRR8494940.91655\t15691\t151,0,151;4056,7213,11269;58,15633,15691\n")
    rp = yb.FromReport(text=text)
    rp.compute_all_bad_part()
    assert rp.get_reads() == {"SRR8494940.65223", "SRR8494940.141626", "SRR8494940.91655"}
    assert rp.get_bad_part("SRR8494940.65223") == ([(0, 1131), (2690, 2706)], 2706)
    assert rp.get_bad_part("SRR8494940.141626") == ([(0, 326), (2957, 30116)], 30116)
    assert rp.get_bad_part("SRR8494940.91655") == ([(0, 151), (7213, 11269), (15633, 15691)], 15691)
    assert [l + "\n" for l in rp.report_lines()] == text.splitlines(keepends=True)  # types recomputed on device
    rp.ctx.close()
    with pytest.raises(yb.YacrdError):
        yb.FromReport(text=text[:-13] + "58,156\n")
    rp = yb.FromReport(text="NotBad\tperfect\t2706\t\n")
    rp.compute_all_bad_part()
    assert rp.get_reads() == {"perfect"} and rp.get_bad_part("perfect") == ([], 2706)
    assert rp.report_lines() == ["NotBad\tperfect\t2706\t"]
    rp.ctx.close()
