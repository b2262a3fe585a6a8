"""The device's crossing formulation == the reference's heap sweep (well-formed intervals). CPU only."""
import random

import pytest

from oracle import yacrd_oracle as o
from tests import kats
from tests.device_model import bad_part_by_crossings, classify


@pytest.mark.parametrize("name,ivs,length,cov,expect", kats.STACK_KATS, ids=[k[0] for k in kats.STACK_KATS])
def test_model_kats(name, ivs, length, cov, expect):
    assert bad_part_by_crossings(ivs, length, cov) == expect


@pytest.mark.parametrize("name,ivs,length,cov,gaps,cls", kats.QUIRK_KATS, ids=[k[0] for k in kats.QUIRK_KATS])
def test_model_quirks(name, ivs, length, cov, gaps, cls):
    assert bad_part_by_crossings(ivs, length, cov) == gaps
    assert classify(length, gaps, 0.8) == cls


def test_model_equals_heap_sweep_fuzz():
    rng = random.Random(7)
    for it in range(40000):
        length = rng.choice([1, 2, 3, 8, 20, 64, 1000, 250000])
        k = rng.choice([0, 1, 2, 3, 4, 6, 10, 20, 45])
        ivs = []
        for _ in range(k):
            if rng.random() < 0.4:
                b = rng.randrange(0, min(length, 4))
            else:
                b = rng.randrange(0, length)
            e = length if rng.random() < 0.3 else rng.randrange(b + 1, length + 1)
            ivs.append((b, e))
        c = rng.choice([0, 0, 0, 1, 2, 3, 4, 5, 9, 100])
        want = o.c_compute_bad_part(ivs, length, c)
        got = bad_part_by_crossings(ivs, length, c)
        assert got == want, (ivs, length, c)
        for n in (0.4, 0.8):
            assert classify(length, got, n) == o.c_type_of_read(length, want, n)
