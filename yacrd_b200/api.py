"""Host-side mirror of the reference's interface for the detect path, over the C ABI.

Reference                                   here
---------                                   ----
reads2ovl::FullMemory (fullmemory.rs:29-99)  FullMemory      add_overlap / add_length / add_overlap_and_length /
trait Reads2Ovl (reads2ovl/mod.rs:43-163)                    init / overlap / length / get_reads  (+ add_csr, bind_csr)
stack::FromOverlap (stack.rs:45-174)         FromOverlap     compute_all_bad_part / get_bad_part / get_reads
stack::FromReport (stack.rs:176-257)         FromReport      same BadPart surface over an existing .yacrd
editor::ReadType (editor/mod.rs:43-59)       ReadType
editor::report (editor/mod.rs:61-83)         BadPart.write_report / report_line
util::get_file_type (util.rs:39-55)          get_file_type

Every call that computes goes through libyacrd_b200.so (sm_100a kernels); nothing here computes a pile-up.
"""
from __future__ import annotations

import ctypes as C
import os
import enum

import numpy as np

from . import _native as N


class ReadType(enum.IntEnum):  # editor/mod.rs:43-59
    NotBad = 0
    Chimeric = 1
    NotCovered = 2

    def as_str(self):
        return self.name


def version():
    return N.lib().yb_version().decode()


def get_file_type(filename):
    """util.rs:39-55. Returns 'm4', 'paf', 'yacrd', 'fastq', 'fasta', 'yovl' or None."""
    t = N.lib().yb_file_type(filename.encode())
    return {ord("m"): "m4", ord("p"): "paf", ord("y"): "yacrd", ord("q"): "fastq", ord("a"): "fasta",
            ord("o"): "yovl"}.get(t)


def _b(s):
    return s if isinstance(s, bytes) else str(s).encode()


def _view(ptr, n, dtype):
    """numpy view (no copy) of n items at a raw address; empty array for n == 0."""
    dtype = np.dtype(dtype)
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_uint8 * (n * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n)


class Context:
    """Owns one yb_ctx (one CUDA device + stream). Single-threaded, like the &mut self traits."""

    def __init__(self, device=-1, read_buffer_size=8192, flags=0, ingest_threads=0):
        self._L = N.lib()
        opts = N.YbOpts(device, read_buffer_size, flags, ingest_threads)
        self._h = self._L.yb_create(C.byref(opts))
        if not self._h:
            raise N.YacrdError(-12, self._L.yb_create_error().decode())
        self._keep = []  # buffers the context borrows

    def close(self):
        if getattr(self, "_h", None):
            self._L.yb_destroy(self._h)
            self._h = None

    __del__ = close

    def _ck(self, rc):
        if rc != N.OK:
            raise N.YacrdError(rc, self._L.yb_last_error(self._h).decode())
        return rc

    def reset(self):
        """Empty the store/results, keep the buffers (next get_overlaps batch, stack.rs:148-161)."""
        self._keep.clear()
        self._ck(self._L.yb_reset(self._h))

    # ---- staged API ----
    def upload(self):
        self._ck(self._L.yb_upload(self._h))

    def compute_device(self, coverage, not_coverage, stream=None):
        self._ck(self._L.yb_compute_device(self._h, int(coverage), float(not_coverage), stream))

    def download(self):
        self._ck(self._L.yb_download(self._h))

    def synchronize(self):
        self._ck(self._L.yb_synchronize(self._h))

    def compute_all(self, coverage, not_coverage):
        self._ck(self._L.yb_compute_all_bad_part(self._h, int(coverage), float(not_coverage)))

    def set_chunk_intervals(self, n_intervals):
        """Streamed batches (the reference's --ondisk-buffer-size, in intervals): compute_all sends a larger CSR through
        the device in chunks, transfers overlapped with the kernels. 0 = one shot."""
        self._ck(self._L.yb_set_chunk_intervals(self._h, int(n_intervals)))

    def time_one_shot(self, coverage, not_coverage):
        """-> (ms of the per-upload kernels, ms of the first, validating detect step) on the uploaded CSR."""
        a, b = C.c_float(), C.c_float()
        self._ck(self._L.yb_time_one_shot(self._h, int(coverage), float(not_coverage), C.byref(a), C.byref(b)))
        return a.value, b.value

    def time_upload_kernels(self):
        """Device milliseconds of the once-per-upload kernels (row statistics, validation, worklist), re-run on the resident CSR."""
        ms = C.c_float(0)
        self._ck(self._L.yb_time_upload_kernels(self._h, C.byref(ms)))
        return float(ms.value)

    def stats(self):
        st = N.YbStats()
        self._ck(self._L.yb_get_stats(self._h, C.byref(st)))
        return {k: int(getattr(st, k)) for k, _ in st._fields_}

    @property
    def stream(self):
        return self._L.yb_stream(self._h)

    def bind_device_bitmap(self, ptr, nbytes):
        self._ck(self._L.yb_bind_device_bitmap(self._h, ptr, nbytes))

    def device_bitmap(self):
        n = C.c_size_t()
        p = self._L.yb_device_class_bitmap(self._h, C.byref(n))
        return p, n.value

    # ---- results (host views, valid until the next mutating call) ----
    def classes(self):
        n = C.c_size_t()
        return _view(self._L.yb_classes(self._h, C.byref(n)), n.value, np.uint8)

    def class_bitmap(self):
        n = C.c_size_t()
        return _view(self._L.yb_class_bitmap(self._h, C.byref(n)), n.value, np.uint8)

    def gap_ptr(self):
        n = C.c_size_t()
        return _view(self._L.yb_gap_ptr(self._h, C.byref(n)), n.value, np.uint32)

    def gaps(self):
        n = C.c_size_t()
        p = self._L.yb_gaps(self._h, C.byref(n))
        return _view(p, 2 * n.value, np.uint32).reshape(-1, 2)


class PinnedCsr:
    """A host CSR in page-locked memory (yb_host_alloc): rowptr u32[n+1], iv u32[m,2], length u32[n]."""

    def __init__(self, n_reads, n_iv):
        L = N.lib()
        self._L = L
        self.n_reads, self.n_iv = int(n_reads), int(n_iv)
        self._p = [L.yb_host_alloc(4 * (self.n_reads + 1)), L.yb_host_alloc(8 * max(1, self.n_iv)),
                   L.yb_host_alloc(4 * max(1, self.n_reads))]
        if not all(self._p):
            raise MemoryError("yb_host_alloc failed")
        self.rowptr = _view(self._p[0], self.n_reads + 1, np.uint32)
        self.iv = _view(self._p[1], 2 * self.n_iv, np.uint32).reshape(-1, 2)
        self.length = _view(self._p[2], self.n_reads, np.uint32)
        self.global_idx = None

    @property
    def nbytes(self):
        return 4 * (self.n_reads + 1) + 8 * self.n_iv + 4 * self.n_reads

    def free(self):
        for p in self._p:
            if p:
                self._L.yb_host_free(p)
        self._p = [None] * 3
        self.rowptr = self.iv = self.length = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _workload():
    """The synthetic workload generator lives outside the product library (workload/libyacrd_synth.so)."""
    try:
        import workload
    except ImportError:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import workload
    return workload


def synth_shard_of(read, n_shards):
    return _workload().lib().yb_synth_shard_of(int(read), int(n_shards))


def synth_csr(n_reads, mean_intervals, profile=N.SYNTH_ONT, seed=20261017, shard=0, n_shards=1, threads=0):
    """Synthetic workload for one shard, as a PinnedCsr (global_idx filled in)."""
    W = _workload()
    L = W.lib()
    spec = W.SynthSpec(seed, n_reads, shard, n_shards, profile, float(mean_intervals))
    n_local = L.yb_synth_count(C.byref(spec))
    gidx = np.zeros(max(1, n_local), dtype=np.uint32)
    rowptr = np.zeros(n_local + 1, dtype=np.uint32)
    length = np.zeros(max(1, n_local), dtype=np.uint32)
    tot = L.yb_synth_plan(C.byref(spec), gidx.ctypes.data, rowptr.ctypes.data, length.ctypes.data)
    if tot > 0xFFFFFFF0:
        raise ValueError("shard has more than 2^32-16 intervals")
    csr = PinnedCsr(n_local, tot)
    csr.rowptr[:] = rowptr
    csr.length[:] = length[:n_local]
    csr.global_idx = gidx[:n_local]
    rc = L.yb_synth_fill(C.byref(spec), gidx.ctypes.data, csr.rowptr.ctypes.data, csr.length.ctypes.data,
                         n_local, csr.iv.ctypes.data if tot else None, threads)
    if rc != 0:
        raise RuntimeError("yb_synth_fill failed (%d)" % rc)
    return csr


class FullMemory(Context):
    """reads2ovl::FullMemory (fullmemory.rs:29-99) + trait Reads2Ovl (reads2ovl/mod.rs:43-163)."""

    def __init__(self, read_buffer_size=8192, device=-1, ingest_threads=0, host_only=False, lazy_device=False):
        super().__init__(device=device, read_buffer_size=read_buffer_size, ingest_threads=ingest_threads,
                         flags=(N.FLAG_HOST_ONLY if host_only else 0) | (N.FLAG_LAZY_DEVICE if lazy_device else 0))

    def init(self, filename):  # mod.rs:44-81
        self._ck(self._L.yb_init_file(self._h, _b(filename)))

    def init_buffer(self, text, fmt):
        text = _b(text)
        self._ck(self._L.yb_init_buffer(self._h, text, len(text), ord("p") if fmt == "paf" else ord("m")))

    def add_overlap(self, id, ovl):  # mod.rs:154
        i = _b(id)
        self._ck(self._L.yb_add_overlap(self._h, i, len(i), ovl[0], ovl[1]))

    def add_length(self, id, length):  # mod.rs:155
        i = _b(id)
        self._ck(self._L.yb_add_length(self._h, i, len(i), length))

    def add_overlap_and_length(self, id, ovl, length):  # mod.rs:157
        i = _b(id)
        self._ck(self._L.yb_add_overlap_and_length(self._h, i, len(i), ovl[0], ovl[1], length))

    def add_csr(self, rowptr, iv, length, ids=None):
        rowptr = np.ascontiguousarray(rowptr, dtype=np.uint32)
        iv = np.ascontiguousarray(iv, dtype=np.uint32).reshape(-1, 2)
        length = np.ascontiguousarray(length, dtype=np.uint32)
        n = length.shape[0]
        assert rowptr.shape[0] == n + 1
        if ids is None:
            self._ck(self._L.yb_add_csr(self._h, rowptr.ctypes.data, iv.ctypes.data, length.ctypes.data, n, None, None))
        else:
            bs = [_b(i) for i in ids]
            arr = (C.c_char_p * n)(*bs)
            lens = (C.c_size_t * n)(*[len(b) for b in bs])
            self._ck(self._L.yb_add_csr(self._h, rowptr.ctypes.data, iv.ctypes.data, length.ctypes.data, n,
                                        C.cast(arr, C.c_void_p), C.cast(lens, C.c_void_p)))

    def bind_csr(self, csr):
        """Zero-copy: the context borrows a PinnedCsr (or any object with rowptr/iv/length u32 arrays)."""
        self._keep.append(csr)
        self._ck(self._L.yb_bind_csr(self._h, csr.rowptr.ctypes.data, csr.iv.ctypes.data, csr.length.ctypes.data,
                                     len(csr.length)))

    def overlap(self, id):  # mod.rs:150
        i = _b(id)
        p = N._u32p()
        n = C.c_uint32()
        self._ck(self._L.yb_overlap(self._h, i, len(i), C.byref(p), C.byref(n)))
        if n.value == 0:
            return []
        a = _view(C.cast(p, C.c_void_p).value, 2 * n.value, np.uint32)
        return [(int(a[2 * j]), int(a[2 * j + 1])) for j in range(n.value)]

    def length(self, id):  # mod.rs:152
        i = _b(id)
        return int(self._L.yb_length(self._h, i, len(i)))

    def n_reads(self):
        return int(self._L.yb_n_reads(self._h))

    def read_at(self, idx):
        p = C.c_void_p()
        n = C.c_size_t()
        self._ck(self._L.yb_read_at(self._h, idx, C.byref(p), C.byref(n)))
        return C.string_at(p.value, n.value).decode()

    def read_ids(self):
        """All read ids in first-seen order."""
        return [self.read_at(i) for i in range(self.n_reads())]

    def get_reads(self):  # mod.rs:160
        return set(self.read_ids())


class _BadPart:
    """trait BadPart (stack.rs:35-41) over a Context."""

    ctx: FullMemory
    coverage = 0
    not_coverage = 0.8

    def get_bad_part(self, id):  # stack.rs:164-169: unknown id => ([], 0)
        i = _b(id)
        p = N._u32p()
        n = C.c_uint32()
        ln = C.c_uint64()
        cl = C.c_uint8()
        c = self.ctx
        c._ck(c._L.yb_get_bad_part(c._h, i, len(i), C.byref(p), C.byref(n), C.byref(ln), C.byref(cl)))
        a = _view(C.cast(p, C.c_void_p).value, 2 * n.value, np.uint32)
        return [(int(a[2 * j]), int(a[2 * j + 1])) for j in range(n.value)], int(ln.value)

    def type_of_read(self, id):
        """editor::type_of_read (editor/mod.rs:85-100) as computed by the kernels for `not_coverage`."""
        i = _b(id)
        cl = C.c_uint8()
        c = self.ctx
        c._ck(c._L.yb_get_bad_part(c._h, i, len(i), None, None, None, C.byref(cl)))
        return ReadType(cl.value)

    def get_reads(self):  # stack.rs:171-173
        return self.ctx.get_reads()

    def report_line(self, idx):
        c = self.ctx
        cap = 1 << 16
        while True:
            buf = C.create_string_buffer(cap)
            n = c._L.yb_format_report_line(c._h, idx, buf, cap)
            if n >= 0:
                return buf.raw[:n].decode()
            if cap > (1 << 28):
                c._ck(int(n))
            cap *= 16

    def report_lines(self):
        return [self.report_line(i) for i in range(self.ctx.n_reads())]

    def write_report(self, path):  # main.rs:62-84
        c = self.ctx
        c._ck(c._L.yb_write_report(c._h, _b(path)))

    def classes(self):
        return self.ctx.classes()

    def class_bitmap(self):
        return self.ctx.class_bitmap()

    def gap_csr(self):
        return self.ctx.gap_ptr(), self.ctx.gaps()


class FromOverlap(_BadPart):
    """stack::FromOverlap (stack.rs:45-174). `not_coverage` (-n) is taken here because the kernels fuse
    type_of_read (editor/mod.rs:85-100) into the pile-up."""

    def __init__(self, ovl: FullMemory, coverage: int, not_coverage: float = 0.8):
        self.ctx = ovl
        self.coverage = int(coverage)
        self.not_coverage = float(not_coverage)

    def compute_all_bad_part(self):  # stack.rs:143-162
        self.ctx.compute_all(self.coverage, self.not_coverage)


class FromReport(_BadPart):
    """stack::FromReport (stack.rs:176-257): bad regions come from an existing .yacrd; compute only classifies."""

    def __init__(self, input_path=None, text=None, not_coverage: float = 0.8, device=-1):
        self.ctx = FullMemory(device=device)
        self.not_coverage = float(not_coverage)
        c = self.ctx
        if text is not None:
            t = _b(text)
            c._ck(c._L.yb_init_report_buffer(c._h, t, len(t)))
        else:
            c._ck(c._L.yb_init_report(c._h, _b(input_path)))

    def compute_all_bad_part(self):  # stack.rs:245 (no pile-up); classification runs on the device
        self.ctx.compute_all(0, self.not_coverage)


# ---- editors (reference src/editor/{scrubbing,filter,extract,split}.rs): same names and argument order ----------
def _edit(op, input_path, output_path, badregions, not_covered, buffer_size):
    """The class of a read was computed on the device with the BadPart's own not_coverage; the reference passes the
    same value to its editors (main.rs:87-117), so a different one here is an error, not a silent re-classification."""
    if float(not_covered) != float(badregions.not_coverage):
        raise ValueError("editors use the not_coverage the bad parts were classified with (%r), got %r"
                         % (badregions.not_coverage, not_covered))
    c = badregions.ctx
    c._ck(c._L.yb_edit(c._h, op, _b(input_path), _b(output_path)))


def scrubbing(input_path, output_path, badregions, not_covered, buffer_size=8192):  # editor/scrubbing.rs:34-71
    _edit(N.EDIT_SCRUBB, input_path, output_path, badregions, not_covered, buffer_size)


def filter(input_path, output_path, badregions, not_covered, buffer_size=8192):  # editor/filter.rs:34-63
    _edit(N.EDIT_FILTER, input_path, output_path, badregions, not_covered, buffer_size)


def extract(input_path, output_path, badregions, not_covered, buffer_size=8192):  # editor/extract.rs:34-63
    _edit(N.EDIT_EXTRACT, input_path, output_path, badregions, not_covered, buffer_size)


def split(input_path, output_path, badregions, not_covered, buffer_size=8192):  # editor/split.rs:34-71
    _edit(N.EDIT_SPLIT, input_path, output_path, badregions, not_covered, buffer_size)
