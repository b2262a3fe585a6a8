"""The generated compare-exchange networks of the row-per-lane tier (yacrd_b200/csrc/sortnets.cuh): the committed header is
what tools/gen_sortnets.py produces, and every network sorts (reduced zero-one / random check; the full check runs when
the header is regenerated)."""
import os
import subprocess
import sys

from tests.conftest import REPO


def test_sortnets_header_is_up_to_date_and_the_networks_sort():
    r = subprocess.run([sys.executable, os.path.join(REPO, "tools", "gen_sortnets.py"), "--check", "--quick"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "up to date" in r.stdout
