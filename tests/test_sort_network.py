"""The sorting network the tile sort falls back to when a tile's keys tie on depth (lgm_b200/csrc/direct_bin.cu): bitonic
merges in their all-ascending form — the first step of a merge of size k pairs i with i ^ (k - 1), the later steps i with
i ^ j — applied to n elements that are NOT padded to a power of two: a partner index >= n is skipped, which is the same
as comparing with a virtual +infinity that no comparator moves.  Restated here in Python with the kernel's loop bounds
and checked for every n up to 300 and some larger sizes.  CPU only."""
import numpy as np


def network_sort(a):
    a = list(a)
    n = len(a)
    k = 2
    while (k >> 1) < n:                      # for (k = 2; (k >> 1) < n; k <<= 1)
        j = k >> 1
        while j > 0:                         #   for (j = k >> 1; j > 0; j >>= 1)
            flip = (k - 1) if j == (k >> 1) else j
            for i in range(n):
                p = i ^ flip
                if p > i and p < n and a[p] < a[i]:
                    a[i], a[p] = a[p], a[i]
            j >>= 1
        k <<= 1
    return a


def test_unpadded_bitonic_network_sorts_every_length():
    rng = np.random.default_rng(0)
    for n in list(range(0, 301)) + [511, 512, 513, 1000, 2047, 2049]:
        keys = rng.integers(0, 1 << 62, size=n, dtype=np.int64)
        if n > 3:
            keys[rng.integers(0, n, size=n // 3)] = keys[0]      # ties
        assert network_sort(keys.tolist()) == sorted(keys.tolist()), n


def test_network_is_chosen_only_when_the_rank_loop_costs_more():
    """The criterion of the kernel: sum of squared bucket sizes (compares of the rank loop) > n * L (L + 1) / 2 with
    L = ceil(log2 n) (twice the comparators of the network)."""
    def takes_network(sizes):
        n = sum(sizes)
        lg = max(n - 1, 1).bit_length()
        return sum(c * c for c in sizes) > n * (lg * (lg + 1) // 2)
    assert not takes_network([3] * 4000)                 # evenly filled buckets: rank loop
    assert not takes_network([300] + [3] * 4000)         # one moderately full bucket among many: rank loop
    assert takes_network([15000])                        # every key in one bucket (exact duplicates)
    assert takes_network([1200] + [1] * 50)              # the S-class case of the GPU test
