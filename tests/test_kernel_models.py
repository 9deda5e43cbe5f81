"""CPU models of two device algorithms whose exactness rests on an ordering argument (no GPU, no product code: the arithmetic of
slaf_b200/csrc/kernels.cuh restated in numpy and checked as properties).

* K6b (rank-value tokenization without a sorting network): the two-level bucket function must be NON-DECREASING in the key, the
  fine buckets must fit the counter array, and "bucket order, then the full composite inside a bucket" must equal a sort.
* K2w / K2f (percentile mode): the a*-th smallest distinct value found by dropping set bits / by rounds of "smallest key not below
  the bound" must equal the a*-th element of the sorted distinct values.
"""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

N_COARSE = 256


def two_level_buckets(ik: np.ndarray):
    """fine-bucket index of every inverted key, as tokenize_rank_bucket_kernel computes it (kernels.cuh, steps 1-4)."""
    ik = ik.astype(np.uint64)
    lo, hi = int(ik.min()), int(ik.max())
    rng = hi - lo
    sh1 = (rng.bit_length() - 8) if rng >= N_COARSE else 0
    d = ik - lo
    cb = (d >> sh1).astype(np.int64)
    assert cb.max() < N_COARSE
    cc = np.bincount(cb, minlength=N_COARSE)
    lg = np.array([min(((int(c) - 1) >> 1).bit_length() if c > 2 else 0, sh1) for c in cc], np.int64)
    first = np.concatenate(([0], np.cumsum(1 << lg)[:-1]))
    sub = ((d & ((1 << sh1) - 1)) >> (sh1 - lg[cb]).astype(np.uint64)).astype(np.int64)
    return first[cb] + sub, int(np.sum(1 << lg))


def bucket_order(ik: np.ndarray):
    """order of the rows K6b emits: by fine bucket, inside a bucket by the composite (~key, row)."""
    fine, n_fine = two_level_buckets(ik)
    comp = (ik.astype(np.uint64) << np.uint64(32)) | np.arange(ik.size, dtype=np.uint64)
    return np.lexsort((comp, fine)), fine, n_fine


KEYSETS = {
    "uniform": lambda r, n: r.integers(0, 2**32, n, dtype=np.uint64),
    "clustered": lambda r, n: (np.uint64(0x3F000000) + r.choice([0, 1 << 20, 3 << 21], n).astype(np.uint64) + r.integers(0, 4000, n).astype(np.uint64)),
    "all_equal": lambda r, n: np.full(n, 12345678, np.uint64),
    "two_values": lambda r, n: r.choice([7, 2**32 - 9], n).astype(np.uint64),
    "lognormal_bits": lambda r, n: np.float32(r.lognormal(0.0, 1.5, n)).view(np.uint32).astype(np.uint64),
    "narrow": lambda r, n: np.uint64(1000) + r.integers(0, 100, n).astype(np.uint64),
    "ramp": lambda r, n: np.arange(n, dtype=np.uint64) * np.uint64(977),
}


@pytest.mark.parametrize("kind", sorted(KEYSETS))
@pytest.mark.parametrize("n", [1, 2, 255, 256, 2100, 4096, 8192])
def test_two_level_bucket_order_is_a_sort(kind, n):
    r = np.random.default_rng(n * 7 + len(kind))
    ik = KEYSETS[kind](r, n) & np.uint64(0xFFFFFFFF)
    order, fine, n_fine = bucket_order(ik)
    comp = (ik << np.uint64(32)) | np.arange(n, dtype=np.uint64)
    assert np.array_equal(order, np.argsort(comp, kind="stable"))           # composites are unique: the sort is total
    assert n_fine <= n + 2 * N_COARSE and fine.max() < n_fine                # the counter array of the kernel holds them
    s = np.argsort(ik, kind="stable")
    assert np.all(np.diff(fine[s]) >= 0)                                      # the bucket function is monotone in the key


@settings(max_examples=200, deadline=None)
@given(st.lists(st.integers(0, 2**32 - 1), min_size=1, max_size=600))
def test_two_level_bucket_order_property(keys):
    ik = np.array(keys, np.uint64)
    order, _, _ = bucket_order(ik)
    comp = (ik << np.uint64(32)) | np.arange(ik.size, dtype=np.uint64)
    assert np.array_equal(order, np.argsort(comp, kind="stable"))


def target_asc_rank(D: int, k: int, p: float) -> int:
    """a* = max(D - min(k, D) + 1, ceil(p * D / 100)) with the f64 product first (aggregators.py:184-188), at least 1"""
    t = (p * float(D)) / 100.0
    return max(D - min(k, D) + 1, int(np.ceil(t)), 1)


@settings(max_examples=300, deadline=None)
@given(st.sets(st.integers(0, 8191), min_size=1, max_size=300), st.integers(1, 400), st.sampled_from([0.0, 10.0, 37.5, 50.0, 99.9, 100.0]))
def test_percentile_threshold_from_presence_bits(values, k, p):
    """K2w: presence of the values below 32 in a register mask, of the others in a bitmap; the threshold is the a*-th set bit."""
    vals = sorted(values)
    D = len(vals)
    a = min(target_asc_rank(D, k, p), D)
    m0 = sum(1 << v for v in vals if v < 32)
    words = [0] * 256
    for v in vals:
        if v >= 32:
            words[v >> 5] |= 1 << (v & 31)
    d0 = bin(m0).count("1")
    if a <= d0:
        x = m0
        for _ in range(a - 1):
            x &= x - 1                                                       # drop the lowest set bit
        thr = (x & -x).bit_length() - 1
    else:
        want, thr = a - d0, None
        for wi in range(1, 256):
            c = bin(words[wi]).count("1")
            if want <= c:
                x = words[wi]
                for _ in range(want - 1):
                    x &= x - 1
                thr = (wi << 5) + (x & -x).bit_length() - 1
                break
            want -= c
    assert thr == vals[a - 1]


@settings(max_examples=300, deadline=None)
@given(st.sets(st.floats(0.0, 1e6, width=32, allow_nan=False), min_size=1, max_size=255), st.integers(1, 300), st.sampled_from([0.0, 10.0, 50.0, 100.0]))
def test_percentile_threshold_from_rounds_of_minima(values, k, p):
    """K2f: a* rounds of "smallest ordered key not below the bound" over the (unordered) set of distinct values."""
    keys = np.array(sorted(values), np.float32).view(np.uint32) | np.uint32(0x80000000)        # order-preserving key of a non-negative float
    table = np.random.default_rng(len(values)).permutation(keys)
    a = min(target_asc_rank(len(keys), k, p), len(keys))
    lowb, sel = 0, None
    for _ in range(a):
        cand = table[table >= lowb]
        sel = int(cand.min())
        lowb = sel + 1
    assert sel == int(np.sort(keys)[a - 1])


@settings(max_examples=300, deadline=None)
@given(st.integers(0, 7), st.lists(st.booleans(), min_size=1, max_size=1500), st.integers(1, 1200))
def test_in_place_compaction_with_prefetch_never_reads_overwritten_rows(sh, keep, limit):
    """K2w / K2f percentile mode: the kept genes are compacted IN PLACE in the gene stage, 256 rows (32 lanes x 8) per step, and the
    loads of step t + 1 are issued before the stores of step t.  Model of that schedule: every load must still see the original
    gene of its row, and the result must be the first `limit` kept genes in row order."""
    n = len(keep)
    stage = np.full(sh + n + 8, -1, np.int64)
    stage[sh:sh + n] = np.arange(n) + 1000                                   # gene of row i = 1000 + i
    ngroups = (sh + n + 7) // 8
    def load(step):                                                          # what the 32 lanes read for a step: (group, 8 genes)
        out = []
        for lane in range(32):
            gi = min(step * 32 + lane, ngroups - 1)                          # clamped, as in the kernel
            out.append((step * 32 + lane, stage[gi * 8: gi * 8 + 8].copy()))
        return out
    base, nxt, step = 0, load(0), 0
    while step * 32 < ngroups and base < limit:
        cur, nxt = nxt, load(step + 1)                                       # next step's loads first ...
        for gi, genes in cur:                                                # ... then this step's stores, lane by lane
            if gi >= ngroups:
                continue
            for i in range(8):
                row = gi * 8 + i - sh
                if 0 <= row < n and keep[row]:
                    assert genes[i] == 1000 + row                            # the load saw the original gene
                    if base < limit:
                        stage[sh + base] = genes[i]
                    base += 1
        step += 1
    want = [1000 + i for i in range(n) if keep[i]][:limit]
    got = stage[sh: sh + min(base, limit)].tolist()
    assert got == want[: len(got)] and (len(got) == len(want) or base >= limit)


@settings(max_examples=200, deadline=None)
@given(st.lists(st.integers(0, 2**32 - 2), min_size=1, max_size=1200), st.integers(0, 3))
def test_home_slot_lookups_count_distinct_values_exactly(keys, seed):
    """K2f: the sweep only asks "is the key in its HOME slot"; everything else (a new key, a key that linear probing displaced)
    takes the inserting path, which is idempotent.  The number of successful inserts is the number of distinct keys as long as
    the 512-slot set does not fill up (the kernel stops trusting it beyond 255 entries)."""
    SLOTS = 512
    tab = [None] * SLOTS
    inserted = 0
    home = lambda k: ((k * 2654435761) & 0xFFFFFFFF) >> (32 - 9)
    order = np.random.default_rng(seed).permutation(len(keys))
    for i in order:
        k = keys[i]
        if tab[home(k)] == k:
            continue                                                          # hot path: seen
        h = home(k)                                                           # inserting path (also reached by displaced keys)
        while True:
            if tab[h] is None:
                tab[h] = k
                inserted += 1
                break
            if tab[h] == k:
                break
            h = (h + 1) & (SLOTS - 1)
        if inserted > 255:
            break
    D = len(set(keys))
    assert inserted == D if D <= 255 else inserted > 255
