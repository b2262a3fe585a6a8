"""CPU restatement of yacrd's post-detection editors — TEST INFRASTRUCTURE ONLY (see yacrd_oracle.py's header:
only tests/, __graft_entry__.smoke() and bench.py's reference legs may import anything under oracle/).

Follows, line by line, the reference's src/editor/scrubbing.rs:73-236, filter.rs:66-228, extract.rs:66-232,
split.rs:73-226 over plain Python bytes. Record syntax is what the reference obtains from noodles-fasta 0.45 /
noodles-fastq 0.16 (Cargo.lock; the crates are not vendored): fastq = 4 lines, '+' line written bare, name /
description separated by the first space; fasta sequences re-wrapped at 80 columns.
Pinned by the reference's goldens: tests/truth.{filter,extract,split,scrubb}.fastq are reproduced byte for byte from
tests/reads.fastq + tests/truth.yacrd (tests/test_editors.py, digests in tests/golden/c1_editors.json), and by the
fasta / fastq KATs of scrubbing.rs:240-396, split.rs, filter.rs, extract.rs (tests/kats.py)."""
from . import yacrd_oracle as o

SCRUBB, FILTER, EXTRACT, SPLIT = 0, 1, 2, 3


def _plan(op, bads, length, not_covered):
    """-> (drop, whole, [(b, e), ...])"""
    t = o.type_of_read(length, bads, not_covered)
    if op == FILTER:  # filter.rs:92
        return t != o.NOT_BAD, True, []
    if op == EXTRACT:  # extract.rs:91
        return t == o.NOT_BAD, True, []
    if t == o.NOT_COVERED:  # scrubbing.rs:93, split.rs:95
        return True, False, []
    if op == SCRUBB:  # scrubbing.rs:95-121
        if not bads:
            return False, True, []
        poss = [0]
        for b, e in bads:
            poss += [b, e]
        if poss[-1] != length:
            poss.append(length)
        it = poss[2:] if poss[0] == 0 and poss[1] == 0 else poss
        return False, False, [(it[i], it[i + 1]) for i in range(0, len(it) - 1, 2)]
    if t == o.NOT_BAD:  # split.rs:97
        return False, True, []
    poss = [0]
    for b, e in bads:  # split.rs:104-112
        if b == 0 or e == length:
            continue
        poss += [b, e]
    poss.append(length)
    return False, False, [(poss[i], poss[i + 1]) for i in range(0, len(poss) - 1, 2)]


def fastq(op, data, get_bad_part, not_covered):
    """data: bytes of a fastq file; get_bad_part(id: str) -> (bads, length) (unknown id => ([], 0), stack.rs:164-169)."""
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    out = []
    i = 0
    while i < len(lines):
        if not lines[i]:
            i += 1
            continue
        d, seq, plus, qual = (l.rstrip(b"\r") for l in lines[i:i + 4])
        assert d[:1] == b"@" and plus[:1] == b"+"
        i += 4
        name, _, desc = d[1:].partition(b" ")
        rid = name.decode().split()[0]
        bads, length = get_bad_part(rid)
        drop, whole, cuts = _plan(op, bads, length, not_covered)
        if drop:
            continue
        tail = (b" " + desc) if desc else b""
        if whole:
            out.append(b"@" + name + tail + b"\n" + seq + b"\n+\n" + qual + b"\n")
            continue
        for b, e in cuts:
            if b > len(seq) or e > len(seq):
                break
            out.append(b"@" + name + b"_%d_%d" % (b, e) + tail + b"\n" + seq[b:e] + b"\n+\n" + qual[b:e] + b"\n")
    return b"".join(out)


def _wrap(seq):
    return b"".join(seq[i:i + 80] + b"\n" for i in range(0, len(seq), 80))


def fasta(op, data, get_bad_part, not_covered):
    out = []
    for chunk in data.split(b">")[1:]:
        head, _, body = chunk.partition(b"\n")
        head = head.rstrip(b"\r")
        seq = b"".join(l.rstrip(b"\r") for l in body.split(b"\n"))
        k = 0
        while k < len(head) and head[k:k + 1] not in (b" ", b"\t"):
            k += 1
        name, desc = head[:k], head[k + 1:]
        bads, length = get_bad_part(name.decode())
        drop, whole, cuts = _plan(op, bads, length, not_covered)
        if drop:
            continue
        if whole:
            out.append(b">" + name + ((b" " + desc) if desc else b"") + b"\n" + _wrap(seq))
            continue
        for b, e in cuts:
            if b > len(seq) or e > len(seq):
                break
            out.append(b">" + name + b"_%d_%d\n" % (b, e) + _wrap(seq[b:e]))  # the description is dropped (scrubbing.rs:139-150)
    return b"".join(out)


def overlaps(op, data, get_bad_part, not_covered, sep, col_b):
    """filter.rs:139-228 / extract.rs:139-232 on paf (sep '\\t', col_b 5) or m4 (sep ' ', col_b 1)."""
    out = []
    for line in data.split(b"\n"):
        line = line.rstrip(b"\r")
        if not line:
            continue
        f = line.split(sep)
        ok = [o.type_of_read(ln, bads, not_covered) == o.NOT_BAD
              for bads, ln in (get_bad_part(f[0].decode()), get_bad_part(f[col_b].decode()))]
        if (op == FILTER and all(ok)) or (op == EXTRACT and not all(ok)):
            out.append(line + b"\n")
    return b"".join(out)


def report_lookup(report_text):
    """get_bad_part over a .yacrd report (stack.rs:182-215 line syntax)."""
    table = {}
    for line in report_text.splitlines():
        if not line:
            continue
        _, rid, length, bads = line.split("\t")
        table[rid] = ([(int(x.split(",")[1]), int(x.split(",")[2])) for x in bads.split(";") if x], int(length))
    return lambda rid: table.get(rid, ([], 0))
