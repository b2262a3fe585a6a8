"""Peer-memory all-gather of the class bitmap (include/yacrd_b200.h): two ranks, CUDA IPC, no NCCL. Runs on one GPU too
(both ranks on device 0: the peer stores are then ordinary device stores, the flag barrier is the same)."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.gpu
def test_peer_allgather_one_gpu_per_rank():
    """The same over NVLink peer memory: needs two visible GPUs, and says so when it cannot run (a record that only shows
    the shared-GPU variant below has not exercised peer stores)."""
    if _device_count() < 2:
        pytest.skip("only %d GPU visible: the peer-memory all-gather over NVLink is not exercised by this run "
                    "(bench.py --gpus N checks the gathered bitmap against the oracle on multi-GPU boxes)" % _device_count())
    _run_two_ranks(expect="one GPU per rank")


@pytest.mark.gpu
def test_peer_allgather_two_ranks():
    _run_two_ranks()


def _run_two_ranks(expect=None):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "peer_worker.py"), str(r), "2", str(port), "30000"],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and ("rank %d ok" % r) in out, out[-3000:]
        if expect:
            assert expect in out, out[-3000:]
