"""Ingestion: the multi-threaded PAF / m4 ingester gives exactly what the sequential loop and
the reference's rules give (reads2ovl/mod.rs:83-145, io.rs:24-50, fullmemory.rs:82-90): first-seen read order,
first-seen length, arrival order of the intervals inside a read, same errors. Host logic only: runs without a GPU
through YB_FLAG_HOST_ONLY contexts (which cannot compute)."""
import os
import random

import pytest

import yacrd_b200 as yb
from oracle import yacrd_oracle as o
from tests.conftest import GOLDEN


def _random_paf(rng, n_records, n_ids, m4=False, crlf=False, blank_every=0):
    ids = ["read%d/%d" % (i, rng.randrange(1000)) if i % 3 else "r%d" % i for i in range(n_ids)]
    lens = {}
    lines = []
    for i in range(n_records):
        a, b = rng.choice(ids), rng.choice(ids)
        la = lens.setdefault(a, rng.randrange(100, 60000)) if rng.random() < 0.9 else rng.randrange(100, 60000)
        lb = lens.setdefault(b, rng.randrange(100, 60000)) if rng.random() < 0.9 else rng.randrange(100, 60000)
        ba = rng.randrange(0, la - 1)
        ea = rng.randrange(ba + 1, la + 1)
        bb = rng.randrange(0, lb - 1)
        eb = rng.randrange(bb + 1, lb + 1)
        if m4:
            lines.append("%s %s -%d.5 %d 0 %d %d %d %d %d %d %d" % (a, b, i % 90, 80 + i % 19, ba, ea, la, i & 1, bb, eb, lb))
        else:
            lines.append("%s\t%d\t%d\t%d\t%s\t%s\t%d\t%d\t%d\t%d\t%d\t255\ttp:A:S" % (a, la, ba, ea, "+-"[i & 1], b, lb, bb, eb, ea - ba, ea - ba))
        if blank_every and i % blank_every == 0:
            lines.append("")
    nl = "\r\n" if crlf else "\n"
    return nl.join(lines) + (nl if rng.random() < 0.5 else "")


def _snapshot(fm):
    ids = fm.read_ids()
    return ids, [fm.length(i) for i in ids], [fm.overlap(i) for i in ids]


@pytest.mark.parametrize("fmt,crlf,threads", [("paf", False, 2), ("paf", True, 5), ("m4", False, 8), ("paf", False, 13)])
def test_parallel_ingest_equals_sequential_and_reference_rules(tmp_path, fmt, crlf, threads):
    rng = random.Random(hash((fmt, crlf, threads)) & 0xFFFF)
    text = _random_paf(rng, 6000, 700, m4=(fmt == "m4"), crlf=crlf, blank_every=97)
    seq = yb.FullMemory(host_only=True, ingest_threads=1)
    seq.init_buffer(text, fmt)
    par = yb.FullMemory(host_only=True, ingest_threads=threads)
    par.init_buffer(text, fmt)
    a, b = _snapshot(seq), _snapshot(par)
    assert a == b
    path = tmp_path / ("x." + fmt)
    path.write_bytes(text.encode())
    want = (o.ingest_m4 if fmt == "m4" else o.ingest_paf)(str(path))
    assert a[0] == list(want.keys())
    assert a[1] == [v[1] for v in want.values()]
    assert a[2] == [v[0] for v in want.values()]
    seq.close()
    par.close()


def test_parallel_ingest_reference_unit_vectors():
    """reads2ovl/mod.rs:170-237 (PAF_FILE / M4_FILE), through the parallel path."""
    paf = ("1\t12000\t20\t4500\t-\t2\t10000\t5500\t10000\t4500\t4500\t255\n"
           "1\t12000\t5500\t10000\t-\t3\t10000\t0\t4500\t4500\t4500\t255\n")
    m4 = ("1 2 0.1 2 0 20 4500 12000 0 5500 10000 10000\n"
          "1 3 0.1 2 0 5500 10000 12000 0 0 4500 10000\n")
    for text, fmt in ((paf, "paf"), (m4, "m4")):
        fm = yb.FullMemory(host_only=True, ingest_threads=3)
        fm.init_buffer(text, fmt)
        assert fm.get_reads() == {"1", "2", "3"}
        assert fm.overlap("1") == [(20, 4500), (5500, 10000)]
        assert fm.overlap("2") == [(5500, 10000)]
        assert fm.overlap("3") == [(0, 4500)]
        assert fm.length("1") == 12000 and fm.length("2") == 10000 and fm.length("nope") == 0
        fm.close()


def test_parallel_ingest_reports_the_same_bad_record():
    rng = random.Random(3)
    lines = _random_paf(rng, 3000, 100).split("\n")
    lines[1777] = lines[1777].replace("\t", " ", 3)  # too few tab-separated columns that parse
    text = "\n".join(lines)
    msgs = []
    for threads in (1, 6):
        fm = yb.FullMemory(host_only=True, ingest_threads=threads)
        with pytest.raises(yb.YacrdError) as e:
            fm.init_buffer(text, "paf")
        assert e.value.kind == "ReadingError"
        msgs.append(str(e.value))
        fm.close()
    assert msgs[0] == msgs[1] and "record 1778" in msgs[0]


def test_adding_after_a_parallel_ingest_keeps_everything():
    rng = random.Random(9)
    text = _random_paf(rng, 2000, 150)
    par = yb.FullMemory(host_only=True, ingest_threads=4)
    par.init_buffer(text, "paf")
    seq = yb.FullMemory(host_only=True, ingest_threads=1)
    seq.init_buffer(text, "paf")
    for fm in (par, seq):
        fm.add_overlap_and_length("r0", (1, 2), 77)       # known read: length stays
        fm.add_overlap_and_length("brand_new", (5, 9), 10)
    assert _snapshot(par) == _snapshot(seq)
    assert par.overlap("brand_new") == [(5, 9)] and par.length("brand_new") == 10
    par.close()
    seq.close()


def test_host_only_context_cannot_compute():
    fm = yb.FullMemory(host_only=True)
    fm.add_overlap_and_length("a", (1, 5), 10)
    with pytest.raises(yb.YacrdError) as e:
        yb.FromOverlap(fm, 0, 0.8).compute_all_bad_part()
    assert e.value.kind == "Cuda"
    fm.close()


@pytest.mark.parametrize("codec", ["gz", "bz2", "xz"])
def test_compressed_overlap_files_are_sniffed_by_magic_number(tmp_path, codec):
    """util.rs:57-87 (niffler): the compression of the input is detected from its first bytes, not from its name."""
    import bz2
    import gzip
    import lzma
    raw = open(os.path.join(GOLDEN, "c1_overlaps.paf"), "rb").read()
    packed = {"gz": gzip.compress, "bz2": bz2.compress, "xz": lzma.compress}[codec](raw)
    path = str(tmp_path / "overlaps.paf")  # no suffix that tells
    open(path, "wb").write(packed)
    a = yb.FullMemory(host_only=True)
    a.init(path)
    b = yb.FullMemory(host_only=True)
    b.init(os.path.join(GOLDEN, "c1_overlaps.paf"))
    assert a.read_ids() == b.read_ids()
    for rid in b.read_ids()[:40]:
        assert a.overlap(rid) == b.overlap(rid) and a.length(rid) == b.length(rid)
    a.close()
    b.close()
    open(path, "wb").write(packed[: len(packed) // 2])  # a truncated stream is an error, not a short file
    c = yb.FullMemory(host_only=True)
    with pytest.raises(yb.YacrdError):
        c.init(path)
    c.close()
