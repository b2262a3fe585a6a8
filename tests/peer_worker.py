"""Worker of tests/test_gpu_peer.py: one rank of the peer-memory all-gather (CUDA IPC handles exchanged over gloo)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    rank, world, port, n = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
    import torch
    import torch.distributed as dist
    import yacrd_b200 as yb
    from yacrd_b200 import dist as ybd
    from oracle import yacrd_oracle as o
    ndev = torch.cuda.device_count()
    dev = rank % ndev
    print("rank %d: device %d of %d visible (%s)" % (rank, dev, ndev, "one GPU per rank" if ndev >= world else "ranks share a GPU"))
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % port, rank=rank, world_size=world)
    _, _, counts = ybd.shard_layout(n, world)
    slot = ybd.bitmap_bytes(int(counts.max()))
    csr = yb.synth_csr(n, 40, shard=rank, n_shards=world)
    fm = yb.FullMemory(device=dev)
    fm.bind_csr(csr)
    pg = ybd.PeerGather(fm, slot)
    full = yb.synth_csr(n, 40)
    want, _, _ = o.run_csr(full.rowptr, full.iv, full.length, 4, 0.4)
    for it in range(5):  # several steps: the two halves of the gather buffer alternate, no host barrier in between
        if it < 2:
            fm.upload()
        fm.compute_device(4, 0.4)
        pg.wait()
        fm.synchronize()
        got = pg.current().cpu().numpy()
        assert np.array_equal(ybd.global_classes(got, n, world), want), "rank %d step %d: gathered classes differ" % (rank, it)
    fm.download()
    assert np.array_equal(fm.class_bitmap(), got[rank][: len(fm.class_bitmap())])
    dist.barrier()
    pg.close()
    fm.close()
    dist.destroy_process_group()
    print("rank %d ok" % rank)


if __name__ == "__main__":
    main()
