"""The post-detection editors (reference src/editor/{scrubbing,filter,extract,split}.rs, main.rs:87-117).

CPU part: the Python restatement (oracle/editors_oracle.py) against the reference's own goldens — the four
tests/truth.*.fastq of tests/run.rs:163-300, committed as digests (tests/golden/c1_editors.json) — and against the
reference's unit-test vectors (scrubbing.rs:240-396, split.rs:228-322, filter.rs / extract.rs tests).
GPU part (-m gpu): yb_edit / the Python mirror / the CLI subcommands over device results, against the same goldens,
the same KATs and, on fasta and overlap files, the restatement."""
import gzip
import hashlib
import json
import os
import subprocess

import pytest

from oracle import editors_oracle as eo
from oracle import yacrd_oracle as o
from tests.conftest import GOLDEN, REPO

DIGESTS = json.load(open(os.path.join(GOLDEN, "c1_editors.json")))
OPS = {"scrubb": eo.SCRUBB, "filter": eo.FILTER, "extract": eo.EXTRACT, "split": eo.SPLIT}

FASTA_LONG = b">1\nACTGGGGGGACTGGGGGGACTG\n>2\nACTG\n>3\nACTG\n"
FASTQ_LONG = b"@1\nACTGGGGGGACTGGGGGGACTG\n+\n??????????????????????\n@2\nACTG\n+\n????\n@3\nACTG\n+\n????\n"
FASTA_SHORT = b">1\nACTG\n>2\nACTG\n>3\nACTG\n"
FASTQ_SHORT = b"@1\nACTG\n+\n????\n@2\nACTG\n+\n????\n@3\nACTG\n+\n????\n"
PAF = (b"1\t12000\t20\t4500\t-\t2\t10000\t5500\t10000\t4500\t4500\t255\n"
       b"1\t12000\t5500\t10000\t-\t3\t10000\t0\t4500\t4500\t4500\t255\n")
M4 = b"1 2 0.1 2 0 100 450 1000 0 550 900 1000\n1 3 0.1 2 0 550 900 1000 0 100 450 1000\n"
# (name, editor, file type, read "1": overlaps and length (coverage 0, not_coverage 0.8), input, expected output)
EDITOR_KATS = [
    ("scrubb_fasta_keep_begin_end", "scrubb", "fasta", [(0, 4), (9, 13), (18, 22)], 22, FASTA_LONG,  # scrubbing.rs:248-282
     b">1_0_4\nACTG\n>1_9_13\nACTG\n>1_18_22\nACTG\n>2\nACTG\n>3\nACTG\n"),
    ("scrubb_fasta_keep_middle", "scrubb", "fasta", [(4, 18)], 22, FASTA_LONG,  # scrubbing.rs:284-307
     b">1_4_18\nGGGGGACTGGGGGG\n>2\nACTG\n>3\nACTG\n"),
    ("scrubb_fastq_keep_begin_end", "scrubb", "fastq", [(0, 4), (9, 13), (18, 22)], 22, FASTQ_LONG,  # scrubbing.rs:309-362
     b"@1_0_4\nACTG\n+\n????\n@1_9_13\nACTG\n+\n????\n@1_18_22\nACTG\n+\n????\n@2\nACTG\n+\n????\n@3\nACTG\n+\n????\n"),
    ("scrubb_fastq_keep_middle", "scrubb", "fastq", [(4, 18)], 22, FASTQ_LONG,  # scrubbing.rs:364-395
     b"@1_4_18\nGGGGGACTGGGGGG\n+\n??????????????\n@2\nACTG\n+\n????\n@3\nACTG\n+\n????\n"),
    ("split_fasta", "split", "fasta", [(9, 13), (18, 22)], 22, FASTA_LONG,  # split.rs:236-270
     b">1_0_13\nACTGGGGGGACTG\n>1_18_22\nACTG\n>2\nACTG\n>3\nACTG\n"),
    ("split_fastq", "split", "fastq", [(9, 13), (18, 22)], 22, FASTQ_LONG,  # split.rs:272-321
     b"@1_0_13\nACTGGGGGGACTG\n+\n?????????????\n@1_18_22\nACTG\n+\n????\n@2\nACTG\n+\n????\n@3\nACTG\n+\n????\n"),
    ("filter_fasta", "filter", "fasta", [(10, 490), (510, 1000)], 1000, FASTA_SHORT, b">2\nACTG\n>3\nACTG\n"),  # filter.rs:236-266
    ("filter_fastq", "filter", "fastq", [(10, 490), (510, 1000)], 1000, FASTQ_SHORT, b"@2\nACTG\n+\n????\n@3\nACTG\n+\n????\n"),
    ("filter_paf", "filter", "paf", [(10, 490), (510, 1000)], 1000, PAF, b""),  # filter.rs:304-324
    ("filter_m4", "filter", "m4", [(10, 490), (510, 1000)], 1000, M4, b""),  # filter.rs:326-346
    ("extract_fasta", "extract", "fasta", [(10, 490), (510, 1000)], 1000, FASTA_SHORT, b">1\nACTG\n"),  # extract.rs tests
    ("extract_fastq", "extract", "fastq", [(10, 490), (510, 1000)], 1000, FASTQ_SHORT, b"@1\nACTG\n+\n????\n"),
    ("extract_paf", "extract", "paf", [(10, 490), (510, 1000)], 1000, PAF, PAF),
    ("extract_m4", "extract", "m4", [(10, 490), (510, 1000)], 1000, M4, M4),
]


def oracle_edit(op, ftype, data, lookup, n):
    if ftype == "fastq":
        return eo.fastq(OPS[op], data, lookup, n)
    if ftype == "fasta":
        return eo.fasta(OPS[op], data, lookup, n)
    return eo.overlaps(OPS[op], data, lookup, n, b"\t" if ftype == "paf" else b" ", 5 if ftype == "paf" else 1)


def reads_fastq():
    return gzip.open(os.path.join(GOLDEN, "c1_reads.fastq.gz"), "rb").read()


# ---- CPU: the restatement against the reference's goldens and unit-test vectors -----------------------------------
def test_fixture_is_the_reference_input():
    raw = reads_fastq()
    assert hashlib.sha256(raw).hexdigest() == DIGESTS["input"]["sha256"] and raw.count(b"\n") // 4 == 461


@pytest.mark.parametrize("op", sorted(OPS))
def test_oracle_reproduces_reference_goldens(op):
    look = eo.report_lookup(open(os.path.join(GOLDEN, "c1_truth.sorted.yacrd")).read())
    got = eo.fastq(OPS[op], reads_fastq(), look, 0.8)
    assert len(got) == DIGESTS[op]["bytes"] and got.count(b"\n") // 4 == DIGESTS[op]["records"]
    assert hashlib.sha256(got).hexdigest() == DIGESTS[op]["sha256"]
    if op == "extract":
        assert got == open(os.path.join(GOLDEN, "c1_truth.extract.fastq"), "rb").read()


@pytest.mark.parametrize("kat", EDITOR_KATS, ids=[k[0] for k in EDITOR_KATS])
def test_oracle_reference_unit_vectors(kat):
    _, op, ftype, ovls, length, data, want = kat
    bads = o.compute_bad_part(ovls, length, 0)
    look = lambda rid: (bads, length) if rid == "1" else ([], 0)  # noqa: E731
    assert oracle_edit(op, ftype, data, look, 0.8) == want


def test_oracle_fasta_wraps_at_80_and_keeps_description_on_whole_records():
    seq = b"ACGT" * 50
    data = b">r1 some description\n" + seq[:70] + b"\n" + seq[70:] + b"\n"
    out = eo.fasta(eo.FILTER, data, lambda rid: ([], 0), 0.8)
    assert out == b">r1 some description\n" + seq[:80] + b"\n" + seq[80:160] + b"\n" + seq[160:] + b"\n"


# ---- GPU: the product's editors ------------------------------------------------------------------------------------
def _detect_golden(yb):
    fm = yb.FullMemory()
    fm.init(os.path.join(GOLDEN, "c1_overlaps.paf"))
    bp = yb.FromOverlap(fm, 0, 0.8)
    bp.compute_all_bad_part()
    return fm, bp


@pytest.mark.gpu
@pytest.mark.parametrize("op", sorted(OPS))
def test_editors_reproduce_reference_goldens(op, tmp_path):
    import yacrd_b200 as yb
    src = str(tmp_path / "reads.fastq")
    open(src, "wb").write(reads_fastq())
    fm, bp = _detect_golden(yb)
    out = str(tmp_path / ("reads.%s.fastq" % op))
    {"scrubb": yb.scrubbing, "filter": yb.filter, "extract": yb.extract, "split": yb.split}[op](src, out, bp, 0.8)
    got = open(out, "rb").read()
    assert len(got) == DIGESTS[op]["bytes"] and hashlib.sha256(got).hexdigest() == DIGESTS[op]["sha256"]
    fm.close()
    # the same from an existing report (main.rs:43-45: FromReport), as `yacrd -i report.yacrd ... <editor>` does
    rp = yb.FromReport(os.path.join(GOLDEN, "c1_truth.sorted.yacrd"), not_coverage=0.8)
    rp.compute_all_bad_part()
    out2 = str(tmp_path / ("again.%s.fastq" % op))
    {"scrubb": yb.scrubbing, "filter": yb.filter, "extract": yb.extract, "split": yb.split}[op](src, out2, rp, 0.8)
    assert open(out2, "rb").read() == got
    rp.ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kat", EDITOR_KATS, ids=[k[0] for k in EDITOR_KATS])
def test_editors_reference_unit_vectors(kat, tmp_path):
    import yacrd_b200 as yb
    _, op, ftype, ovls, length, data, want = kat
    fm = yb.FullMemory()
    fm.add_length("1", length)
    for iv in ovls:
        fm.add_overlap("1", iv)
    bp = yb.FromOverlap(fm, 0, 0.8)
    bp.compute_all_bad_part()
    src, out = str(tmp_path / ("in." + ftype)), str(tmp_path / ("out." + ftype))
    open(src, "wb").write(data)
    {"scrubb": yb.scrubbing, "filter": yb.filter, "extract": yb.extract, "split": yb.split}[op](src, out, bp, 0.8)
    assert open(out, "rb").read() == want
    fm.close()


@pytest.mark.gpu
def test_editors_on_fasta_and_overlaps_match_the_restatement(tmp_path):
    """The golden reads as fasta (long lines: re-wrapped at 80) and the golden PAF / M4 themselves through every editor
    that accepts them, at -c 4 -n 0.4 as well, against oracle/editors_oracle.py."""
    import yacrd_b200 as yb
    fq = reads_fastq().split(b"\n")
    fasta = b"".join(b">" + fq[i][1:] + b"\n" + fq[i + 1] + b"\n" for i in range(0, len(fq) - 1, 4))
    inputs = {"fasta": fasta, "paf": open(os.path.join(GOLDEN, "c1_overlaps.paf"), "rb").read(),
              "m4": open(os.path.join(GOLDEN, "c1_overlaps.m4"), "rb").read()}
    for c, n in ((0, 0.8), (4, 0.4)):
        fm = yb.FullMemory()
        fm.init(os.path.join(GOLDEN, "c1_overlaps.paf"))
        bp = yb.FromOverlap(fm, c, n)
        bp.compute_all_bad_part()
        look = eo.report_lookup("\n".join(bp.report_lines()))
        for ftype, data in inputs.items():
            src = str(tmp_path / ("in." + ftype))
            open(src, "wb").write(data)
            for op in (("scrubb", "filter", "extract", "split") if ftype == "fasta" else ("filter", "extract")):
                out = str(tmp_path / ("out.%s.%s" % (op, ftype)))
                {"scrubb": yb.scrubbing, "filter": yb.filter, "extract": yb.extract, "split": yb.split}[op](src, out, bp, n)
                assert open(out, "rb").read() == oracle_edit(op, ftype, data, look, n), (c, n, ftype, op)
        fm.close()


@pytest.mark.gpu
def test_editor_errors(tmp_path):
    import yacrd_b200 as yb
    fm, bp = _detect_golden(yb)
    out = str(tmp_path / "o")
    for fn, path, kind in ((yb.scrubbing, os.path.join(GOLDEN, "c1_overlaps.paf"), "CantRunOperationOnFile"),  # scrubbing.rs:48-57
                           (yb.split, os.path.join(GOLDEN, "c1_overlaps.m4"), "CantRunOperationOnFile"),
                           (yb.filter, os.path.join(GOLDEN, "c1_truth.sorted.yacrd"), "CantRunOperationOnFile"),
                           (yb.filter, str(tmp_path / "reads.txt"), "UnableToDetectFileFormat"),
                           (yb.extract, str(tmp_path / "missing.fastq"), "CantReadFile")):
        with pytest.raises(yb.YacrdError) as e:
            fn(path, out, bp, 0.8)
        assert e.value.kind == kind, (path, e.value.kind)
    bad = str(tmp_path / "bad.fastq")
    open(bad, "wb").write(b"@r1\nACGT\n+\n")
    with pytest.raises(yb.YacrdError) as e:
        yb.filter(bad, out, bp, 0.8)
    assert e.value.kind == "ReadingError"
    with pytest.raises(ValueError):
        yb.filter(bad, out, bp, 0.4)  # not the not_coverage the classes were computed with
    fm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("codec", ["gz", "bz2", "xz"])
def test_compressed_input_keeps_its_compression(tmp_path, codec):
    """util.rs:57-87: compression is sniffed from the magic number and the output keeps the input's codec (level 1)."""
    import bz2
    import lzma
    import yacrd_b200 as yb
    pack = {"gz": gzip.compress, "bz2": bz2.compress, "xz": lzma.compress}[codec]
    unpack = {"gz": gzip.decompress, "bz2": bz2.decompress, "xz": lzma.decompress}[codec]
    magic = {"gz": b"\x1f\x8b", "bz2": b"BZh", "xz": b"\xfd7zXZ\x00"}[codec]
    paf = str(tmp_path / ("overlaps.paf." + codec))
    open(paf, "wb").write(pack(open(os.path.join(GOLDEN, "c1_overlaps.paf"), "rb").read()))
    reads = str(tmp_path / ("reads.fastq." + codec))
    open(reads, "wb").write(pack(reads_fastq()))
    fm = yb.FullMemory()
    fm.init(paf)
    bp = yb.FromOverlap(fm, 0, 0.8)
    bp.compute_all_bad_part()
    want = sorted(l for l in open(os.path.join(GOLDEN, "c1_truth.sorted.yacrd")).read().split("\n") if l)
    assert sorted(bp.report_lines()) == want
    for op, fn in (("scrubb", yb.scrubbing), ("filter", yb.filter), ("extract", yb.extract), ("split", yb.split)):
        out = str(tmp_path / ("reads.%s.fastq.%s" % (op, codec)))
        fn(reads, out, bp, 0.8)
        raw = open(out, "rb").read()
        assert raw.startswith(magic)
        assert hashlib.sha256(unpack(raw)).hexdigest() == DIGESTS[op]["sha256"]
    fm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("op", sorted(OPS))
def test_cli_subcommands_replay_run_rs(op, tmp_path):
    """tests/run.rs:163-300: `yacrd -i reads.paf -o result.yacrd <editor> -i reads.fastq -o reads.<editor>.fastq`."""
    exe = os.path.join(REPO, "yacrd_b200", "yacrd-b200")
    src = str(tmp_path / "reads.fastq")
    open(src, "wb").write(reads_fastq())
    rep, out = str(tmp_path / "result.yacrd"), str(tmp_path / ("reads.%s.fastq" % op))
    r = subprocess.run([exe, "-i", os.path.join(GOLDEN, "c1_overlaps.paf"), "-o", rep, op, "-i", src, "-o", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert sorted(l for l in open(rep).read().split("\n") if l) == \
        sorted(l for l in open(os.path.join(GOLDEN, "c1_truth.sorted.yacrd")).read().split("\n") if l)
    assert hashlib.sha256(open(out, "rb").read()).hexdigest() == DIGESTS[op]["sha256"]
