"""CPU ORACLE (test infrastructure, NOT the product).

Two independent restatements of the reference's detect hot path (natir/yacrd 1.0.0, Rust):

* pure-Python, literal (``heapq`` + ``list.sort``): ``compute_bad_part`` <- src/stack.rs:61-139,
  ``type_of_read`` <- src/editor/mod.rs:85-100, ``report_line`` <- src/editor/mod.rs:61-83,102-107,
  ``ingest_paf`` / ``ingest_m4`` <- src/reads2ovl/mod.rs:83-145 + src/io.rs:24-50 +
  src/reads2ovl/fullmemory.rs:82-90 (first-seen length wins, no dedup). Small cases only.
* ctypes binding of ``oracle/yacrd_oracle.c`` (same algorithm in C, pthread batch driver) for sizes
  up to BASELINE.json's configs and for the CPU baseline.

PARITY PIN: the Rust reference cannot be built here (no cargo/rustc) => no oracle/_ref; both
restatements are pinned by the reference's own vectors (tests/test_oracle.py, tests/golden/).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.
"""
from __future__ import annotations

import ctypes
import heapq
import os
import subprocess

import numpy as np

NOT_BAD, CHIMERIC, NOT_COVERED = 0, 1, 2
TYPE_NAMES = ("NotBad", "Chimeric", "NotCovered")  # editor/mod.rs:51-58
U32 = 0xFFFFFFFF

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libyacrd_oracle.so")


# --------------------------------------------------------------------------------------------
# pure-Python literal restatement
# --------------------------------------------------------------------------------------------
def compute_bad_part(ovls, length, coverage):
    """src/stack.rs:61-139, line by line. ``ovls``: iterable of (begin, end)."""
    gaps = []
    stack = []  # BinaryHeap<Reverse<u32>>
    ovls = sorted((int(b), int(e)) for b, e in ovls)  # stack.rs:66
    first_covered = 0
    last_covered = 0
    for b, e in ovls:  # stack.rs:71
        while stack:  # stack.rs:72
            head = stack[0]
            if head > b:  # stack.rs:73-75
                break
            if len(stack) > coverage:  # stack.rs:77-79
                last_covered = head
            heapq.heappop(stack)  # stack.rs:80
        if len(stack) <= coverage:  # stack.rs:83
            if last_covered != 0:
                gaps.append((last_covered, b))  # stack.rs:85
            else:
                first_covered = b  # stack.rs:87
        heapq.heappush(stack, e)  # stack.rs:90
    while len(stack) > coverage:  # stack.rs:93
        last_covered = stack[0]
        if last_covered >= length:  # stack.rs:101-103
            break
        heapq.heappop(stack)
    if first_covered != 0:  # stack.rs:107-109
        gaps.insert(0, (0, first_covered))
    if last_covered != length:  # stack.rs:111-113
        gaps.append((last_covered, length & U32))
    if not gaps:  # stack.rs:115-117
        return gaps
    clean = []  # stack.rs:119-138
    begin, end = gaps[0]
    for g1, g2 in zip(gaps, gaps[1:]):
        if g1[0] == g2[0]:
            begin = g1[0]
            end = max(g1[1], g2[1])
        else:
            clean.append((begin, end))
            begin, end = g2
    clean.append((begin, end))
    return clean


def type_of_read(length, badregions, not_covered):
    """src/editor/mod.rs:85-100 (u32 wrapping sum, f64 ratio, NotCovered first, strict >)."""
    bad_region_len = 0
    for b, e in badregions:
        bad_region_len = (bad_region_len + ((e - b) & U32)) & U32
    if length == 0:
        ratio = float("nan") if bad_region_len == 0 else float("inf")
    else:
        ratio = float(bad_region_len) / float(length)
    if ratio > not_covered:
        return NOT_COVERED
    for b, e in badregions:
        if b != 0 and e != (length & U32):
            return CHIMERIC
    return NOT_BAD


def bad_region_format(bads):
    """src/editor/mod.rs:102-107."""
    return ";".join("%d,%d,%d" % ((e - b) & U32, b, e) for b, e in bads)


def report_line(read, length, badregions, not_covered):
    """src/editor/mod.rs:61-83 (without the trailing newline)."""
    t = type_of_read(length, badregions, not_covered)
    return "%s\t%s\t%d\t%s" % (TYPE_NAMES[t], read, length, bad_region_format(badregions))


def _ingest(path, sep, cols):
    """src/reads2ovl/mod.rs:83-145: positional parse; two intervals per record;
    fullmemory.rs:82-90: first-seen length wins, duplicates kept. Returns an insertion-ordered
    dict id -> ([(b, e), ...], length)."""
    ia, la, ba, ea, ib, lb, bb, eb = cols
    reads = {}
    with open(path, "rt") as fh:
        for line in fh:
            line = line.rstrip("\n").rstrip("\r")
            if not line:
                continue
            f = line.split(sep)
            for rid, ln, b, e in ((f[ia], f[la], f[ba], f[ea]), (f[ib], f[lb], f[bb], f[eb])):
                ent = reads.get(rid)
                if ent is None:
                    reads[rid] = ([(int(b), int(e))], int(ln))
                else:
                    ent[0].append((int(b), int(e)))
    return reads


def ingest_paf(path):
    """io.rs:24-34 column order: read_a len_a beg_a end_a strand read_b len_b beg_b end_b."""
    return _ingest(path, "\t", (0, 1, 2, 3, 5, 6, 7, 8))


def ingest_m4(path):
    """io.rs:37-50: read_a read_b err shared strand_a beg_a end_a len_a strand_b beg_b end_b len_b."""
    return _ingest(path, " ", (0, 7, 5, 6, 1, 11, 9, 10))


def detect_lines(reads, coverage, not_covered):
    """main.rs:78-84 restated over an ingest dict; returns the report lines (unordered contract)."""
    out = []
    for rid, (ovls, length) in reads.items():
        out.append(report_line(rid, length, compute_bad_part(ovls, length, coverage), not_covered))
    return out


# --------------------------------------------------------------------------------------------
# C oracle binding
# --------------------------------------------------------------------------------------------
_lib = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        u32p = ctypes.POINTER(ctypes.c_uint32)
        u64p = ctypes.POINTER(ctypes.c_uint64)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        L.yo_compute_bad_part.restype = ctypes.c_uint32
        L.yo_compute_bad_part.argtypes = [u32p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64, u32p]
        L.yo_type_of_read.restype = ctypes.c_int
        L.yo_type_of_read.argtypes = [ctypes.c_uint64, u32p, ctypes.c_uint32, ctypes.c_double]
        L.yo_format_line.restype = ctypes.c_long
        L.yo_format_line.argtypes = [ctypes.c_char_p, ctypes.c_uint64, u32p, ctypes.c_uint32,
                                     ctypes.c_double, ctypes.c_char_p, ctypes.c_size_t]
        L.yo_max_threads.restype = ctypes.c_int
        L.yo_run_csr.restype = ctypes.c_uint64
        L.yo_run_csr.argtypes = [u64p, u32p, u32p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_double,
                                 u8p, u64p, u32p, ctypes.c_int]
        L.yo_run_csr_padded.restype = ctypes.c_uint64
        L.yo_run_csr_padded.argtypes = [u64p, u32p, u32p, ctypes.c_uint32, ctypes.c_uint64,
                                        ctypes.c_double, u8p, u32p, u32p, ctypes.c_int]
        _lib = L
    return _lib


def _p(a, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


def c_compute_bad_part(ovls, length, coverage):
    iv = np.ascontiguousarray(np.asarray(ovls, dtype=np.uint32).reshape(-1, 2))
    k = iv.shape[0]
    gaps = np.zeros((k + 2, 2), dtype=np.uint32)
    n = lib().yo_compute_bad_part(_p(iv, ctypes.c_uint32), k, int(length), int(coverage),
                                  _p(gaps, ctypes.c_uint32))
    return [(int(b), int(e)) for b, e in gaps[:n]]


def c_type_of_read(length, bads, not_covered):
    g = np.ascontiguousarray(np.asarray(bads, dtype=np.uint32).reshape(-1, 2))
    return lib().yo_type_of_read(int(length), _p(g, ctypes.c_uint32), g.shape[0], float(not_covered))


def c_report_line(read, length, bads, not_covered):
    g = np.ascontiguousarray(np.asarray(bads, dtype=np.uint32).reshape(-1, 2))
    cap = 64 + len(read) + 36 * (g.shape[0] + 1)
    buf = ctypes.create_string_buffer(cap)
    n = lib().yo_format_line(read.encode(), int(length), _p(g, ctypes.c_uint32), g.shape[0],
                             float(not_covered), buf, cap)
    assert n > 0
    return buf.raw[: n - 1].decode()


def run_csr(rowptr, iv, length, coverage, not_covered, threads=0):
    """Batch oracle over a CSR. Returns (cls u8[n], gap_ptr u64[n+1], gaps u32[g,2])."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.uint64)
    iv = np.ascontiguousarray(iv, dtype=np.uint32).reshape(-1, 2)
    length = np.ascontiguousarray(length, dtype=np.uint32)
    n = length.shape[0]
    assert rowptr.shape[0] == n + 1 and int(rowptr[-1]) == iv.shape[0]
    cls = np.zeros(n, dtype=np.uint8)
    gap_ptr = np.zeros(n + 1, dtype=np.uint64)
    gaps = np.zeros((iv.shape[0] + 2 * n + 1, 2), dtype=np.uint32)
    tot = lib().yo_run_csr(_p(rowptr, ctypes.c_uint64), _p(iv, ctypes.c_uint32),
                           _p(length, ctypes.c_uint32), n, int(coverage), float(not_covered),
                           _p(cls, ctypes.c_uint8), _p(gap_ptr, ctypes.c_uint64),
                           _p(gaps, ctypes.c_uint32), int(threads))
    return cls, gap_ptr, gaps[:tot].copy()


class PaddedRunner:
    """Pre-allocated buffers for timing yo_run_csr_padded (the CPU-baseline region)."""

    def __init__(self, rowptr, iv, length):
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.uint64)
        self.iv = np.ascontiguousarray(iv, dtype=np.uint32).reshape(-1, 2)
        self.length = np.ascontiguousarray(length, dtype=np.uint32)
        self.n = self.length.shape[0]
        self.cls = np.zeros(self.n, dtype=np.uint8)
        self.cnt = np.zeros(self.n, dtype=np.uint32)
        self.padded = np.zeros((self.iv.shape[0] + 2 * self.n + 1, 2), dtype=np.uint32)

    def run(self, coverage, not_covered, threads=0):
        return lib().yo_run_csr_padded(_p(self.rowptr, ctypes.c_uint64), _p(self.iv, ctypes.c_uint32),
                                       _p(self.length, ctypes.c_uint32), self.n, int(coverage),
                                       float(not_covered), _p(self.cls, ctypes.c_uint8),
                                       _p(self.cnt, ctypes.c_uint32), _p(self.padded, ctypes.c_uint32),
                                       int(threads))


def max_threads():
    return lib().yo_max_threads()
