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


class PeerGather:
    """All-gather of the class bitmap fused into the kernels' epilogue over NVLink peer memory (include/yacrd_b200.h,
    "peer-memory all-gather"): every rank allocates one gather buffer [2 x world x slot_bytes] (step s uses half s & 1)
    and one flag buffer, the CUDA IPC handles travel through torch.distributed (any backend), every rank maps every
    other rank's buffers, and from then on `ctx.compute_device()` leaves all ranks' bitmaps in every rank's buffer - no
    separate collective and nothing in a step that waits for the slowest rank. A consumer calls `wait()` (a flag wait on
    the stream) and then reads `current()`."""

    def __init__(self, ctx, slot_bytes, group=None):
        import ctypes as C
        import torch.distributed as dist
        from . import _native as N
        self.ctx, self.slot_bytes = ctx, int(slot_bytes)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 16:
            raise ValueError("PeerGather supports at most 16 ranks")
        L = ctx._L
        hg, hf = C.create_string_buffer(64), C.create_string_buffer(64)
        self._own_gather = L.yb_peer_alloc(ctx._h, 2 * self.world * self.slot_bytes, hg)
        self._own_flags = L.yb_peer_alloc(ctx._h, 128, hf)
        if not self._own_gather or not self._own_flags:
            raise N.YacrdError(-12, L.yb_last_error(ctx._h).decode())
        handles = [None] * self.world
        dist.all_gather_object(handles, (hg.raw, hf.raw), group=group)
        self._mapped = []
        gather, flags = [], []
        for p, (g, f) in enumerate(handles):
            if p == self.rank:
                gather.append(self._own_gather)
                flags.append(self._own_flags)
                continue
            pg, pf = L.yb_peer_open(ctx._h, g), L.yb_peer_open(ctx._h, f)
            if not pg or not pf:
                raise N.YacrdError(-12, L.yb_last_error(ctx._h).decode())
            self._mapped += [pg, pf]
            gather.append(pg)
            flags.append(pf)
        ga = (C.c_void_p * self.world)(*gather)
        fa = (C.c_void_p * self.world)(*flags)
        ctx._ck(L.yb_bind_peers(ctx._h, ga, fa, self.world, self.rank, self.slot_bytes))
        dist.barrier(group=group)  # nobody starts writing before everybody has mapped

    class _Arr:
        def __init__(self, ptr, shape, typestr):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(ptr), False), "version": 2}

    def tensor(self):
        """This rank's gather buffer as a [2, world, slot_bytes] uint8 torch tensor (a view, no copy)."""
        import torch
        return torch.as_tensor(self._Arr(self._own_gather, (2, self.world, self.slot_bytes), "|u1"), device="cuda")

    def flags(self):
        """This rank's flag words (int32 view): [q] = steps rank q has finished, [31] = steps this rank has finished."""
        import torch
        return torch.as_tensor(self._Arr(self._own_flags, (32,), "<i4"), device="cuda")

    def wait(self, stream=None):
        """Enqueues (on `stream`, a raw cudaStream_t, default the context's) the wait for every rank's slot of this
        rank's last finished step."""
        self.ctx._ck(self.ctx._L.yb_peer_wait(self.ctx._h, stream))

    def current(self):
        """[world, slot_bytes] view of the half the last finished step wrote (synchronises to read the step count)."""
        steps = int(self.flags()[31].item())
        return self.tensor()[(steps - 1) & 1]

    def close(self):
        L = self.ctx._L
        if getattr(self, "_own_gather", None) and self.ctx._h:
            L.yb_bind_peers(self.ctx._h, None, None, 0, 0, 0)
            for m in self._mapped:
                L.yb_peer_close(self.ctx._h, m)
            L.yb_peer_free(self.ctx._h, self._own_gather)
            L.yb_peer_free(self.ctx._h, self._own_flags)
        self._own_gather = None
