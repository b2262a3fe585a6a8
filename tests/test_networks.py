"""The compare-exchange networks compiled into the register tier (yacrd_b200/csrc/regtier.cuh), read from the source:
the 60-exchange network sorts every zero-one input of 16 keys (so it sorts everything), Batcher's 16 + 16 odd-even merge
merges every pair of sorted zero-one halves, and their composition (185 exchanges) sorts 32 keys. CPU only."""
import os
import random
import re

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _networks():
    src = open(os.path.join(REPO, "yacrd_b200", "csrc", "regtier.cuh")).read()
    body16 = src[src.index("void sort16_t("):src.index("// the lane's 32 keys")]
    body32 = src[src.index("void sort32_t("):src.index("// bitonic half-cleaners on the slot bits")]
    ce = lambda text: [(int(i), int(j)) for _, i, j in re.findall(r"CE\((\d+), (\d+), (\d+)\)", text)]
    return ce(body16), ce(body32)


def test_sort16_zero_one_principle():
    n16, _ = _networks()
    assert len(n16) == 60
    x = np.arange(1 << 16, dtype=np.uint32)
    bits = [((x >> i) & 1).astype(np.uint8) for i in range(16)]
    for i, j in n16:
        assert i < j
        bits[i], bits[j] = np.minimum(bits[i], bits[j]), np.maximum(bits[i], bits[j])
    assert all((bits[i] <= bits[i + 1]).all() for i in range(15))


def test_merge16x16_and_sort32():
    n16, m32 = _networks()
    assert len(m32) == 65
    for a in range(17):
        for b in range(17):
            v = [0] * (16 - a) + [1] * a + [0] * (16 - b) + [1] * b
            for i, j in m32:
                if v[i] > v[j]:
                    v[i], v[j] = v[j], v[i]
            assert v == sorted(v)
    net = n16 + [(i + 16, j + 16) for i, j in n16] + m32
    assert len(net) == 185
    rng = random.Random(3)
    for _ in range(3000):
        v = [rng.randrange(0, rng.choice([2, 5, 70000])) for _ in range(32)]
        w = sorted(v)
        for i, j in net:
            if v[i] > v[j]:
                v[i], v[j] = v[j], v[i]
        assert v == w
