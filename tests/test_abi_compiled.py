"""A C program compiled against include/yacrd_b200.h and linked with the in-tree library drives the ABI the way a
cgo / FFI host would (tests/abi_smoke.c): header and library cannot drift apart behind the ctypes signature table."""
import os
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(REPO, "include"),
                    os.path.join(REPO, "tests", "abi_smoke.c"), "-L", os.path.join(REPO, "yacrd_b200"), "-lyacrd_b200",
                    "-Wl,-rpath," + os.path.join(REPO, "yacrd_b200"), "-o", exe], check=True)
    return exe


def test_abi_smoke_compiles_and_links(tmp_path):
    """CPU: the header is valid C11 and every entry the program uses resolves against the library."""
    assert os.path.exists(_build(tmp_path))


@pytest.mark.gpu
def test_abi_smoke_runs(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, str(tmp_path / "out.yacrd")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "abi_smoke ok" in r.stdout, r.stdout + r.stderr
