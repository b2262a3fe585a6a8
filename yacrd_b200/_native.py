"""ctypes binding of the C ABI declared in include/yacrd_b200.h (libyacrd_b200.so, built in-tree by
``python -m yacrd_b200.build`` / ``__graft_entry__.build()``).

There is no Python or CPU implementation of the detect path behind this module: if the shared library
is missing, importing the symbols fails loudly, and without a CUDA device ``yb_create`` returns NULL.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# YB_LIB_PATH: a development knob (A/B runs of differently compiled kernels); the product is the in-tree library
LIB_PATH = os.environ.get("YB_LIB_PATH") or os.path.join(_HERE, "libyacrd_b200.so")

OK = 0
ERR_NAMES = {
    -1: "CantReadFile", -2: "CantWriteFile", -3: "UnableToDetectFileFormat", -4: "CantRunOperationOnFile",
    -5: "ReadingError", -6: "WritingError", -7: "CorruptYacrdReport", -8: "MalformedInterval",
    -9: "InvalidArgument", -10: "State", -11: "TooLarge", -12: "Cuda", -13: "NoMem",
}
NOT_BAD, CHIMERIC, NOT_COVERED = 0, 1, 2
EDIT_SCRUBB, EDIT_FILTER, EDIT_EXTRACT, EDIT_SPLIT = 0, 1, 2, 3  # yb_editor
FLAG_KEEP_HOST_INTERVALS, FLAG_HOST_ONLY, FLAG_LAZY_DEVICE = 1, 2, 4
SYNTH_ONT, SYNTH_PACBIO_SKEW = 0, 1


class YbOpts(C.Structure):
    _fields_ = [("device", C.c_int32), ("read_buffer_size", C.c_uint32), ("flags", C.c_uint32),
                ("ingest_threads", C.c_uint32)]


class YbStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "n_reads", "n_intervals", "n_gaps", "n_not_bad", "n_chimeric", "n_not_covered",
        "max_intervals_per_read", "n_reads_warp", "n_reads_cta", "n_reads_huge", "kernel_launches",
        "h2d_bytes", "d2h_bytes", "n_malformed_intervals", "n_literal_reads")]


class YbSynthSpec(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_reads", C.c_uint32), ("shard", C.c_uint32),
                ("n_shards", C.c_uint32), ("profile", C.c_uint32), ("mean_intervals", C.c_double)]


_u8p, _u32p, _u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
_vp, _cp, _sz = C.c_void_p, C.c_char_p, C.c_size_t

# name -> (restype, argtypes): every symbol include/yacrd_b200.h declares
SIGNATURES = {
    "yb_create": (_vp, [C.POINTER(YbOpts)]),
    "yb_create_error": (_cp, []),
    "yb_destroy": (None, [_vp]),
    "yb_reset": (C.c_int, [_vp]),
    "yb_last_error": (_cp, [_vp]),
    "yb_version": (_cp, []),
    "yb_type_name": (_cp, [C.c_int]),
    "yb_add_overlap_and_length": (C.c_int, [_vp, _cp, _sz, C.c_uint32, C.c_uint32, C.c_uint64]),
    "yb_add_overlap": (C.c_int, [_vp, _cp, _sz, C.c_uint32, C.c_uint32]),
    "yb_add_length": (C.c_int, [_vp, _cp, _sz, C.c_uint64]),
    "yb_add_csr": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint32, _vp, _vp]),
    "yb_bind_csr": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint32]),
    "yb_host_alloc": (_vp, [_sz]),
    "yb_host_free": (None, [_vp]),
    "yb_init_file": (C.c_int, [_vp, _cp]),
    "yb_init_buffer": (C.c_int, [_vp, _cp, _sz, C.c_int]),
    "yb_file_type": (C.c_int, [_cp]),
    "yb_length": (C.c_uint64, [_vp, _cp, _sz]),
    "yb_overlap": (C.c_int, [_vp, _cp, _sz, C.POINTER(_u32p), _u32p]),
    "yb_n_reads": (C.c_uint32, [_vp]),
    "yb_read_at": (C.c_int, [_vp, C.c_uint32, C.POINTER(_vp), C.POINTER(_sz)]),
    "yb_read_index": (C.c_int64, [_vp, _cp, _sz]),
    "yb_compute_all_bad_part": (C.c_int, [_vp, C.c_uint64, C.c_double]),
    "yb_get_bad_part": (C.c_int, [_vp, _cp, _sz, C.POINTER(_u32p), _u32p, _u64p, _u8p]),
    "yb_get_bad_part_at": (C.c_int, [_vp, C.c_uint32, C.POINTER(_u32p), _u32p, _u64p, _u8p]),
    "yb_write_report": (C.c_int, [_vp, _cp]),
    "yb_edit": (C.c_int, [_vp, C.c_int, _cp, _cp]),
    "yb_device_warmup": (C.c_int, [C.c_int]),
    "yb_format_report_line": (C.c_int64, [_vp, C.c_uint32, _vp, _sz]),
    "yb_classes": (_vp, [_vp, C.POINTER(_sz)]),
    "yb_class_bitmap": (_vp, [_vp, C.POINTER(_sz)]),
    "yb_gap_ptr": (_vp, [_vp, C.POINTER(_sz)]),
    "yb_gaps": (_vp, [_vp, C.POINTER(_sz)]),
    "yb_init_report": (C.c_int, [_vp, _cp]),
    "yb_init_report_buffer": (C.c_int, [_vp, _cp, _sz]),
    "yb_upload": (C.c_int, [_vp]),
    "yb_compute_device": (C.c_int, [_vp, C.c_uint64, C.c_double, _vp]),
    "yb_download": (C.c_int, [_vp]),
    "yb_synchronize": (C.c_int, [_vp]),
    "yb_device_class_bitmap": (_vp, [_vp, C.POINTER(_sz)]),
    "yb_bind_device_bitmap": (C.c_int, [_vp, _vp, _sz]),
    "yb_stream": (_vp, [_vp]),
    "yb_device_classes": (_vp, [_vp, C.POINTER(_sz)]),
    "yb_device_gap_ptr": (_vp, [_vp, C.POINTER(_sz)]),
    "yb_device_gaps": (_vp, [_vp, C.POINTER(_sz)]),
    "yb_get_stats": (C.c_int, [_vp, C.POINTER(YbStats)]),
    "yb_peer_alloc": (_vp, [_vp, _sz, _vp]),
    "yb_peer_open": (_vp, [_vp, _vp]),
    "yb_peer_close": (C.c_int, [_vp, _vp]),
    "yb_peer_free": (C.c_int, [_vp, _vp]),
    "yb_bind_peers": (C.c_int, [_vp, _vp, _vp, C.c_uint32, C.c_uint32, _sz]),
    "yb_peer_wait": (C.c_int, [_vp, _vp]),
    "yb_time_upload_kernels": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "yb_set_chunk_intervals": (C.c_int, [_vp, C.c_uint32]),
    "yb_time_one_shot": (C.c_int, [_vp, C.c_uint64, C.c_double, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
}

_lib = None


class NativeLibraryMissing(ImportError):
    pass


def lib():
    """The loaded C-ABI library. Raises NativeLibraryMissing if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                "%s is missing: build it with `python -m yacrd_b200.build` (nvcc, sm_100a). "
                "yacrd_b200 has no CPU fallback for the detect path." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class YacrdError(RuntimeError):
    def __init__(self, code, message):
        self.code = code
        self.kind = ERR_NAMES.get(code, "Unknown")
        super().__init__("%s (%d): %s" % (self.kind, code, message))
