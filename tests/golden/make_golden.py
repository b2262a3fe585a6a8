#!/usr/bin/env python
"""Regenerates the golden fixtures in this directory from the reference's own test data.

Run in the authoring container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Inputs  (reference, read-only): tests/reads.paf, tests/truth.yacrd  (tests/run.rs:96-117 pins
        `yacrd -i tests/reads.paf -o out` == truth.yacrd as an unordered line set).
Outputs (committed):
  c1_overlaps.paf          the 1286 overlap records, 12 mandatory PAF columns only (the reference
                           parses the first 9, io.rs:24-34; the SAM-like tags are dropped)
  c1_overlaps.m4           the same overlaps re-expressed in BLASR m4 column order (io.rs:37-50), to
                           exercise init_m4 (reads2ovl/mod.rs:115-145) on the same truth
  c1_truth.sorted.yacrd    truth.yacrd, LC_ALL=C sorted (the reference's order is hash-map order and
                           is not part of its contract, tests/run.rs:33-62)
  c1_oracle_c{C}_n{N}.sorted.yacrd
                           ORACLE-DERIVED (not reference goldens): the pinned oracle's report for the
                           README presets the reference never pins (-c 4 -n 0.4, -c 3 -n 0.4,
                           -c 1 -n 0.8). Regression vectors for the oracle itself.
  c1_reads.fastq.gz        tests/reads.fastq (461 reads), gzip -9 with a zero timestamp: the editors' input
  c1_editors.json          sha256 / byte count / record count of tests/truth.{filter,extract,split,scrubb}.fastq
                           (tests/run.rs:163-300 pins the editors' output byte for byte, in input order)
  c1_truth.extract.fastq   the smallest of the four, in full (a readable diff when a digest does not match)
"""
import gzip
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
REF = "/root/reference/tests"

from oracle import yacrd_oracle as o  # noqa: E402


def main():
    recs = [l.rstrip("\n").split("\t") for l in open(os.path.join(REF, "reads.paf")) if l.strip()]
    with open(os.path.join(HERE, "c1_overlaps.paf"), "w") as fh:
        for f in recs:
            fh.write("\t".join(f[:12]) + "\n")
    with open(os.path.join(HERE, "c1_overlaps.m4"), "w") as fh:
        for f in recs:
            ida, la, ba, ea, st, idb, lb, bb, eb = f[:9]
            # read_a read_b error shared_min strand_a beg_a end_a len_a strand_b beg_b end_b len_b
            fh.write(" ".join([ida, idb, "0.1", f[9], "0", ba, ea, la, "1" if st == "-" else "0",
                               bb, eb, lb]) + "\n")
    truth = sorted(open(os.path.join(REF, "truth.yacrd")).read().splitlines())
    with open(os.path.join(HERE, "c1_truth.sorted.yacrd"), "w") as fh:
        fh.write("\n".join(truth) + "\n")
    reads = o.ingest_paf(os.path.join(HERE, "c1_overlaps.paf"))
    assert sorted(o.detect_lines(reads, 0, 0.8)) == truth, "oracle does not reproduce truth.yacrd"
    assert sorted(o.detect_lines(o.ingest_m4(os.path.join(HERE, "c1_overlaps.m4")), 0, 0.8)) == truth
    for c, n in ((4, 0.4), (3, 0.4), (1, 0.8)):
        lines = sorted(o.detect_lines(reads, c, n))
        with open(os.path.join(HERE, "c1_oracle_c%d_n%s.sorted.yacrd" % (c, n)), "w") as fh:
            fh.write("\n".join(lines) + "\n")
        cnt = {t: sum(l.startswith(t + "\t") for l in lines) for t in o.TYPE_NAMES}
        print("c=%d n=%s -> %s" % (c, n, cnt))
    from oracle import editors_oracle as eo
    raw = open(os.path.join(REF, "reads.fastq"), "rb").read()
    with open(os.path.join(HERE, "c1_reads.fastq.gz"), "wb") as fh:
        with gzip.GzipFile(filename="", mode="wb", compresslevel=9, fileobj=fh, mtime=0) as gz:
            gz.write(raw)
    look = eo.report_lookup(open(os.path.join(REF, "truth.yacrd")).read())
    digests = {}
    for op, name in ((eo.FILTER, "filter"), (eo.EXTRACT, "extract"), (eo.SPLIT, "split"), (eo.SCRUBB, "scrubb")):
        want = open(os.path.join(REF, "truth.%s.fastq" % name), "rb").read()
        assert eo.fastq(op, raw, look, 0.8) == want, "editors oracle does not reproduce truth.%s.fastq" % name
        digests[name] = {"sha256": hashlib.sha256(want).hexdigest(), "bytes": len(want), "records": want.count(b"\n") // 4}
    digests["input"] = {"sha256": hashlib.sha256(raw).hexdigest(), "bytes": len(raw), "records": raw.count(b"\n") // 4}
    with open(os.path.join(HERE, "c1_editors.json"), "w") as fh:
        json.dump(digests, fh, indent=1, sort_keys=True)
        fh.write("\n")
    shutil.copyfile(os.path.join(REF, "truth.extract.fastq"), os.path.join(HERE, "c1_truth.extract.fastq"))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
