"""world_size-2 gloo test (CPU) of the multi-GPU host logic: hash sharding, in-place all-gather of the
2-bit class bitmaps, reassembly into global read order. The class codes come from the oracle here (no GPU);
tests/test_gpu_parity.py::test_sharded_* runs the same layout through the kernels."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_READS, MEAN_K, C, N = 3000, 12, 1, 0.4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, REPO)
    import yacrd_b200 as yb
    from oracle import yacrd_oracle as o
    from yacrd_b200 import dist as ybd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    part = yb.synth_csr(N_READS, MEAN_K, shard=rank, n_shards=world)
    cls, _, _ = o.run_csr(part.rowptr, part.iv, part.length, C, N, threads=1)
    _, _, counts = ybd.shard_layout(N_READS, world)
    slot = ybd.bitmap_bytes(int(counts.max()))
    gathered = torch.zeros(world, slot, dtype=torch.uint8)
    bm = ybd.pack_bitmap(cls)
    gathered[rank, : bm.shape[0]] = torch.from_numpy(bm)  # "the kernels wrote into this rank's slot"
    ybd.allgather_bitmaps(gathered[rank], gathered)
    glob = ybd.global_classes(gathered.numpy(), N_READS, world)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), glob)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bitmap_allgather(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, REPO)
    import yacrd_b200 as yb
    from oracle import yacrd_oracle as o
    full = yb.synth_csr(N_READS, MEAN_K)
    want, _, _ = o.run_csr(full.rowptr, full.iv, full.length, C, N, threads=1)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npy" % r))
        assert np.array_equal(got, want)
