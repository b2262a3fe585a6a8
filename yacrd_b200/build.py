"""Builds yacrd_b200/libyacrd_b200.so in-tree (nvcc, -gencode arch=compute_100a,code=sm_100a -lineinfo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose=False, clean=False):
    csrc = os.path.join(HERE, "csrc")
    if clean:
        subprocess.run(["make", "-C", csrc, "clean"], check=True, stdout=subprocess.DEVNULL)
    r = subprocess.run(["make", "-C", csrc, "-j", str(os.cpu_count() or 4)], capture_output=not verbose, text=True)
    if r.returncode != 0:
        if not verbose:
            sys.stderr.write(r.stdout[-4000:] + r.stderr[-8000:])
        raise RuntimeError("building libyacrd_b200.so failed")
    return os.path.join(HERE, "libyacrd_b200.so")


if __name__ == "__main__":
    print(build(verbose=True, clean="--clean" in sys.argv))
