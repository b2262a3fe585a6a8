"""yacrd_b200 — B200-native detect path of yacrd (coverage pile-up -> bad regions -> Chimeric / NotCovered /
NotBad), behind the reference's Reads2Ovl / BadPart surface. The compute path is hand-written sm_100a CUDA
in libyacrd_b200.so (C ABI: include/yacrd_b200.h); this package is the thin host mirror of the reference's
interface. There is no CPU fallback."""
from ._native import (CHIMERIC, NOT_BAD, NOT_COVERED, SYNTH_ONT, SYNTH_PACBIO_SKEW, NativeLibraryMissing,
                      YacrdError)
from .api import (Context, FromOverlap, FromReport, FullMemory, ReadType, PinnedCsr, extract, filter, get_file_type,
                  scrubbing, split, synth_csr, synth_shard_of, version)

__all__ = ["Context", "FullMemory", "FromOverlap", "FromReport", "ReadType", "PinnedCsr", "get_file_type",
           "synth_csr", "synth_shard_of", "version", "scrubbing", "filter", "extract", "split", "YacrdError", "NativeLibraryMissing", "NOT_BAD", "CHIMERIC",
           "NOT_COVERED", "SYNTH_ONT", "SYNTH_PACBIO_SKEW"]
