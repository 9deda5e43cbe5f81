// py_shuffle.h -- host-side, bit-exact restatement of CPython's
//     random.seed(int); random.shuffle(list(range(n)))
// which is what RandomShuffle.apply uses to order the per-cell chunks
// (reference: slaf/ml/samplers.py:89-96, seed from slaf/ml/datasets.py:1016).
//
// The algorithm is CPython's published one (Modules/_randommodule.c: MT19937 seeded with
// init_by_array over the 32-bit limbs of abs(seed); Lib/random.py: Fisher-Yates from the top
// with _randbelow_with_getrandbits = rejection sampling on getrandbits(bit_length(n))).
// tests/test_shuffle.py checks it against the interpreter's own `random` for many (seed, n).
// O(n), ~4 ns per cell: this is control-plane work that stays on the host by design
// (the permutation depends only on (seed, n_cells)); the device applies it as an indirection.
#pragma once
#include <cstddef>
#include <cstdint>

namespace slafb {

class PyRandom {
  public:
    explicit PyRandom(int64_t seed) {
        uint64_t a = seed < 0 ? (uint64_t)(-(seed + 1)) + 1u : (uint64_t)seed;  // abs()
        uint32_t key[2] = {(uint32_t)(a & 0xffffffffu), (uint32_t)(a >> 32)};
        int len = key[1] ? 2 : 1;
        init_by_array(key, len);
    }
    inline uint32_t next_u32() {
        if (idx_ >= N) refill();
        uint32_t y = mt_[idx_++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    // random._randbelow_with_getrandbits(n), n >= 1
    inline uint32_t randbelow(uint32_t n) {
        int k = 32 - __builtin_clz(n);  // n.bit_length()
        uint32_t r = next_u32() >> (32 - k);
        while (r >= n) r = next_u32() >> (32 - k);
        return r;
    }

  private:
    static constexpr int N = 624, M = 397;
    uint32_t mt_[N];
    int idx_ = N;

    void init_genrand(uint32_t s) {
        mt_[0] = s;
        for (int i = 1; i < N; i++) mt_[i] = 1812433253u * (mt_[i - 1] ^ (mt_[i - 1] >> 30)) + (uint32_t)i;
        idx_ = N;
    }
    void init_by_array(const uint32_t *key, int len) {
        init_genrand(19650218u);
        int i = 1, j = 0;
        for (int k = (N > len ? N : len); k; k--) {
            mt_[i] = (mt_[i] ^ ((mt_[i - 1] ^ (mt_[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
            i++; j++;
            if (i >= N) { mt_[0] = mt_[N - 1]; i = 1; }
            if (j >= len) j = 0;
        }
        for (int k = N - 1; k; k--) {
            mt_[i] = (mt_[i] ^ ((mt_[i - 1] ^ (mt_[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            i++;
            if (i >= N) { mt_[0] = mt_[N - 1]; i = 1; }
        }
        mt_[0] = 0x80000000u;
    }
    void refill() {
        static const uint32_t mag01[2] = {0u, 0x9908b0dfu};
        int kk;
        uint32_t y;
        for (kk = 0; kk < N - M; kk++) {
            y = (mt_[kk] & 0x80000000u) | (mt_[kk + 1] & 0x7fffffffu);
            mt_[kk] = mt_[kk + M] ^ (y >> 1) ^ mag01[y & 1u];
        }
        for (; kk < N - 1; kk++) {
            y = (mt_[kk] & 0x80000000u) | (mt_[kk + 1] & 0x7fffffffu);
            mt_[kk] = mt_[kk + (M - N)] ^ (y >> 1) ^ mag01[y & 1u];
        }
        y = (mt_[N - 1] & 0x80000000u) | (mt_[0] & 0x7fffffffu);
        mt_[N - 1] = mt_[M - 1] ^ (y >> 1) ^ mag01[y & 1u];
        idx_ = 0;
    }
};

// perm[j] = index (in first-appearance order) of the cell emitted at output row j.
//
// Same draws as random.shuffle, restructured for speed (this runs once per batch on the host,
// between the CSR build and the tokenizing kernel):
//   phase 1  walk i = n-1 .. 1 grouped by bit_length(i+1) (constant shift inside a group) and
//            record the accepted draw j_i; rejection sampling is done branch-free (the draw is
//            always stored, the index only advances when it is accepted), so the only branch left
//            is the loop condition;
//   phase 2  apply the swaps perm[i] <-> perm[j_i] in a tight loop.
// `scratch` must hold n uint32 (the context keeps one per pipeline slot).
inline void py_shuffle_perm(int64_t seed, int64_t n, int32_t *perm, uint32_t *scratch = nullptr) {
    for (int64_t i = 0; i < n; i++) perm[i] = (int32_t)i;
    if (n < 2) return;
    PyRandom rng(seed);
    uint32_t *jarr = scratch;
    bool own = false;
    if (!jarr) { jarr = new uint32_t[(size_t)n]; own = true; }
    int64_t i = n - 1;
    while (i >= 1) {
        const int k = 32 - __builtin_clz((uint32_t)(i + 1));      // (i+1).bit_length()
        const int64_t lo = ((int64_t)1 << (k - 1)) - 1;           // smallest i with the same bit length of i+1
        const int64_t stop = lo < 1 ? 1 : lo;
        const int sh = 32 - k;
        while (i >= stop) {
            const uint32_t r = rng.next_u32() >> sh;
            jarr[i] = r;
            i -= (r < (uint32_t)(i + 1)) ? 1 : 0;
        }
    }
    for (int64_t t = n - 1; t >= 1; t--) {
        const uint32_t j = jarr[t];
        const int32_t x = perm[t];
        perm[t] = perm[j];
        perm[j] = x;
    }
    if (own) delete[] jarr;
}

}  // namespace slafb
