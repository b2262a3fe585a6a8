"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol include/yacrd_b200.h
declares, the host-only pieces (synthetic generator, file-type rule) behave, and the product refuses to run
without a CUDA device instead of falling back to a CPU path."""
import os
import re

import numpy as np
import pytest

import yacrd_b200 as yb
from yacrd_b200 import _native as N
from yacrd_b200 import dist as ybd

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "yacrd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(yb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = N.lib()
    names = _declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), "libyacrd_b200.so does not export %s" % n
        assert n in N.SIGNATURES, "no ctypes signature for %s" % n
    assert set(N.SIGNATURES) <= set(names), set(N.SIGNATURES) - set(names)


def test_product_does_not_link_or_import_the_oracle():
    for root, _, files in os.walk(os.path.join(REPO, "yacrd_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", "Makefile")):
                src = open(os.path.join(root, f), errors="ignore").read()
                assert "yacrd_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f


def test_product_library_holds_no_measurement_code():
    """The synthetic workload generator lives in workload/libyacrd_synth.so (both bench arms load it from there); the
    product library exports only what include/yacrd_b200.h declares."""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", N.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.split("\n") if " T " in l}
    yb_syms = {s for s in exported if s.startswith("yb_")}
    assert yb_syms == set(_declared_symbols()), yb_syms ^ set(_declared_symbols())
    assert not [s for s in exported if "synth" in s or "oracle" in s]
    import workload
    W = workload.lib()
    for name in ("yb_synth_count", "yb_synth_plan", "yb_synth_fill", "yb_synth_shard_of", "yb_synth_paf"):
        assert hasattr(W, name)


def test_version_and_names():
    assert yb.version().startswith("1.0.0 Magby")
    L = N.lib()
    assert [L.yb_type_name(i).decode() for i in range(3)] == ["NotBad", "Chimeric", "NotCovered"]
    assert [t.as_str() for t in yb.ReadType] == ["NotBad", "Chimeric", "NotCovered"]


@pytest.mark.parametrize("name,expect", [
    ("reads.paf", "paf"), ("x.m4", "m4"), ("x.mhap", "m4"), ("a.yacrd", "yacrd"), ("r.fastq", "fastq"),
    ("r.fq.gz", "fastq"), ("r.fasta", "fasta"), ("r.fa", "fasta"), ("o.yovl", "yovl"), ("plain.txt", None),
    ("both.paf.m4", "m4"),  # util.rs:39-55 tests .m4/.mhap first
])
def test_file_type_rule(name, expect):
    assert yb.get_file_type(name) == expect


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(yb.YacrdError) as e:
        yb.FullMemory()
    assert e.value.code == -12 and "no CPU fallback" in str(e.value)


def test_synth_wellformed_and_deterministic():
    a = yb.synth_csr(20000, 30)
    b = yb.synth_csr(20000, 30, threads=1)
    assert a.n_reads == 20000 and np.array_equal(a.rowptr, b.rowptr) and np.array_equal(a.iv, b.iv)
    k = np.diff(a.rowptr.astype(np.int64))
    assert k.min() >= 1 and 25 < k.mean() < 35
    lens = np.repeat(a.length, k)
    assert (a.iv[:, 0] < a.iv[:, 1]).all() and (a.iv[:, 1] <= lens).all()
    assert a.length.min() >= 200 and a.length.max() <= 250000
    s = yb.synth_csr(5000, 0, profile=yb.SYNTH_PACBIO_SKEW)
    ks = np.diff(s.rowptr.astype(np.int64))
    assert ks.max() <= 5000 and ks.max() > 256 and s.length.min() >= 500


def test_synth_shards_partition_the_workload():
    n, G = 6000, 4
    full = yb.synth_csr(n, 20)
    sh, local, counts = ybd.shard_layout(n, G)
    assert [yb.synth_shard_of(r, G) for r in range(50)] == list(sh[:50])
    seen = 0
    for g in range(G):
        part = yb.synth_csr(n, 20, shard=g, n_shards=G)
        assert part.n_reads == counts[g]
        assert np.array_equal(part.global_idx, np.nonzero(sh == g)[0])
        for i in (0, part.n_reads // 2, part.n_reads - 1):
            r = int(part.global_idx[i])
            assert part.length[i] == full.length[r]
            assert np.array_equal(part.iv[part.rowptr[i]:part.rowptr[i + 1]], full.iv[full.rowptr[r]:full.rowptr[r + 1]])
        seen += part.n_reads
    assert seen == n


def test_bitmap_pack_roundtrip():
    rng = np.random.default_rng(1)
    for n in (0, 1, 3, 4, 15, 16, 17, 1000):
        cls = rng.integers(0, 3, n).astype(np.uint8)
        bm = ybd.pack_bitmap(cls)
        assert bm.shape[0] == ybd.bitmap_bytes(n)
        assert np.array_equal(ybd.unpack_bitmap(bm, n), cls)
