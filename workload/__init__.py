"""Synthetic workloads of BASELINE.json's configs: workload/libyacrd_synth.so (workload/synth.cpp,
declarations in workload/synth.h), measurement and test infrastructure that is not part of the product library. Both
bench arms and the tests generate their inputs here; `synth_csr` fills plain numpy buffers (the reference arm's process
never maps libyacrd_b200.so), yacrd_b200.synth_csr fills a page-locked PinnedCsr through the same entry points."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libyacrd_synth.so")
SYNTH_ONT, SYNTH_PACBIO_SKEW = 0, 1
_lib = None


class SynthSpec(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_reads", C.c_uint32), ("shard", C.c_uint32),
                ("n_shards", C.c_uint32), ("profile", C.c_uint32), ("mean_intervals", C.c_double)]


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.yb_synth_count.restype = C.c_uint32
        L.yb_synth_count.argtypes = [C.POINTER(SynthSpec)]
        L.yb_synth_plan.restype = C.c_uint64
        L.yb_synth_plan.argtypes = [C.POINTER(SynthSpec), C.c_void_p, C.c_void_p, C.c_void_p]
        L.yb_synth_fill.restype = C.c_int
        L.yb_synth_fill.argtypes = [C.POINTER(SynthSpec), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
        L.yb_synth_shard_of.restype = C.c_uint32
        L.yb_synth_shard_of.argtypes = [C.c_uint32, C.c_uint32]
        L.yb_synth_paf.restype = C.c_uint64
        L.yb_synth_paf.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p, C.c_uint64]
        _lib = L
    return _lib


class HostCsr:
    """rowptr u32[n+1], iv u32[m,2], length u32[n] in plain numpy memory (+ global_idx of every local read)."""

    def __init__(self, rowptr, iv, length, global_idx):
        self.rowptr, self.iv, self.length, self.global_idx = rowptr, iv, length, global_idx
        self.n_reads, self.n_iv = len(length), len(iv)

    @property
    def nbytes(self):
        return 4 * (self.n_reads + 1) + 8 * self.n_iv + 4 * self.n_reads


def synth_csr(n_reads, mean_intervals, profile=SYNTH_ONT, seed=20261017, shard=0, n_shards=1, threads=0):
    L = lib()
    spec = SynthSpec(seed, n_reads, shard, n_shards, profile, float(mean_intervals))
    n_local = L.yb_synth_count(C.byref(spec))
    gidx = np.zeros(max(1, n_local), dtype=np.uint32)
    rowptr = np.zeros(n_local + 1, dtype=np.uint32)
    length = np.zeros(max(1, n_local), dtype=np.uint32)
    tot = L.yb_synth_plan(C.byref(spec), gidx.ctypes.data, rowptr.ctypes.data, length.ctypes.data)
    iv = np.zeros((max(1, tot), 2), dtype=np.uint32)
    if L.yb_synth_fill(C.byref(spec), gidx.ctypes.data, rowptr.ctypes.data, length.ctypes.data, n_local,
                       iv.ctypes.data if tot else None, threads) != 0:
        raise RuntimeError("yb_synth_fill failed")
    return HostCsr(rowptr, iv[:tot], length[:n_local], gidx[:n_local])
