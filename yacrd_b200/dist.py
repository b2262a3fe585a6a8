"""Multi-GPU plumbing of the detect path: reads are independent, so they are sharded by
read-id hash across the ranks (one process per GPU) and every rank runs the kernels on its own shard with
no data-path collective; the only exchange is ONE all-gather of the 2-bit-per-read class bitmap at the
end, so that every rank can serve filter/extract-style consumers for any read.

torch.distributed is the plumbing (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(z):
    """splitmix64 finaliser on a numpy uint64 array (same as csrc/synth.cpp mix64)."""
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def shard_of(read_idx, n_shards):
    """Shard of each read index: mix64(idx) % n_shards (0 when n_shards <= 1)."""
    read_idx = np.asarray(read_idx, dtype=np.uint64)
    if n_shards <= 1:
        return np.zeros(read_idx.shape, dtype=np.uint32)
    return (mix64(read_idx) % np.uint64(n_shards)).astype(np.uint32)


def shard_layout(n_reads, n_shards):
    """(shard[r], local[r], counts[s]): read r is the local[r]-th read of shard shard[r]."""
    sh = shard_of(np.arange(n_reads, dtype=np.uint64), n_shards)
    counts = np.bincount(sh, minlength=max(1, n_shards)).astype(np.int64)
    local = np.zeros(n_reads, dtype=np.int64)
    for s in range(max(1, n_shards)):
        m = sh == s
        local[m] = np.arange(int(counts[s]))
    return sh, local, counts


def bitmap_bytes(n_reads):
    """Bytes of the 2-bit bitmap of n_reads reads (whole 32-bit words, as the kernels write it)."""
    return ((int(n_reads) + 15) // 16) * 4


def pack_bitmap(classes):
    """u8 class codes -> 2-bit bitmap (read r in byte r//4, bits 2*(r%4)..+1), padded to whole u32 words."""
    classes = np.asarray(classes, dtype=np.uint8)
    n = classes.shape[0]
    pad = np.zeros(bitmap_bytes(n) * 4, dtype=np.uint8)
    pad[:n] = classes
    q = pad.reshape(-1, 4)
    return (q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)).astype(np.uint8)


def unpack_bitmap(bitmap, n_reads):
    b = np.asarray(bitmap, dtype=np.uint8)
    out = np.empty((b.shape[0], 4), dtype=np.uint8)
    for j in range(4):
        out[:, j] = (b >> (2 * j)) & 3
    return out.reshape(-1)[:n_reads]


def allgather_bitmaps(local_slot, gathered, group=None):
    """In-place all-gather: `gathered` is a [world, slot_bytes] uint8 tensor whose row `rank` is
    `local_slot` (the kernels wrote the bitmap straight into it). One collective, no staging copy."""
    import torch.distributed as dist
    dist.all_gather_into_tensor(gathered.view(-1), local_slot, group=group)
    return gathered


def global_classes(gathered, n_reads, n_shards):
    """[n_shards, slot_bytes] gathered bitmaps -> class code of every global read index."""
    g = np.asarray(gathered, dtype=np.uint8).reshape(max(1, n_shards), -1)
    sh, local, counts = shard_layout(n_reads, n_shards)
    out = np.empty(n_reads, dtype=np.uint8)
    for s in range(max(1, n_shards)):
        cl = unpack_bitmap(g[s], int(counts[s]))
        m = sh == s
        out[m] = cl[local[m]]
    return out
