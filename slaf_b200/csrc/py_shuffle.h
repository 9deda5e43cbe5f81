// py_shuffle.h -- host-side, bit-exact restatement of CPython's
//     random.seed(int); random.shuffle(list(range(n)))
// which is what RandomShuffle.apply uses to order the per-cell chunks
// (reference: slaf/ml/samplers.py:89-96, seed from slaf/ml/datasets.py:1016).
//
// The algorithm is CPython's published one (Modules/_randommodule.c: MT19937 seeded with
// init_by_array over the 32-bit limbs of abs(seed); Lib/random.py: Fisher-Yates from the top
// with _randbelow_with_getrandbits = rejection sampling on getrandbits(bit_length(n))).
// tests/test_abi.py::test_shuffle_matches_cpython_random checks it against the interpreter's own `random` for many (seed, n).
// O(n), ~3 ns per cell: this is control-plane work that stays on the host by design
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
        return out_[idx_++];
    }
    // the tempered outputs still unread in the current block ([*count] of them, at least 1): bulk consumers read them in a
    // tight loop and report how many they used with consume()
    inline const uint32_t *block(int *count) {
        if (idx_ >= N) refill();
        *count = N - idx_;
        return out_ + idx_;
    }
    inline void consume(int used) { idx_ += used; }
    // random._randbelow_with_getrandbits(n), n >= 1
    inline uint32_t randbelow(uint32_t n) {
        int k = 32 - __builtin_clz(n);  // n.bit_length()
        uint32_t r = next_u32() >> (32 - k);
        while (r >= n) r = next_u32() >> (32 - k);
        return r;
    }

  private:
    static constexpr int N = 624, M = 397;
    uint32_t mt_[N + 4];      // state (+ padding so that the vectorised twist may read one vector past the end)
    uint32_t out_[N];         // tempered outputs of the current state block
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
    // one twist of the whole state, then all 624 outputs tempered at once.  Written so that the compiler vectorises it:
    // the conditional xor is arithmetic (-(y & 1) & MATRIX_A) and each loop only reads words that are either still old or
    // were finished at least 227 iterations earlier.
    static inline __attribute__((always_inline)) void twist_and_temper(uint32_t *__restrict mt, uint32_t *__restrict out) {
        constexpr uint32_t A = 0x9908b0dfu, UP = 0x80000000u, LO = 0x7fffffffu;
        for (int kk = 0; kk < N - M; kk++) {
            const uint32_t y = (mt[kk] & UP) | (mt[kk + 1] & LO);
            mt[kk] = mt[kk + M] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
        }
        // kk in [227, 623): reads mt[kk - 227], already new; in chunks shorter than 227 no chunk reads its own writes
        for (int base = N - M; base < N - 1; base += 198) {
            const int end = base + 198 < N - 1 ? base + 198 : N - 1;
            for (int kk = base; kk < end; kk++) {
                const uint32_t y = (mt[kk] & UP) | (mt[kk + 1] & LO);
                mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
            }
        }
        const uint32_t y = (mt[N - 1] & UP) | (mt[0] & LO);
        mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
        for (int kk = 0; kk < N; kk++) {
            uint32_t t = mt[kk];
            t ^= (t >> 11);
            t ^= (t << 7) & 0x9d2c5680u;
            t ^= (t << 15) & 0xefc60000u;
            t ^= (t >> 18);
            out[kk] = t;
        }
    }
#if defined(__x86_64__) && defined(__GNUC__)
    // the same loops compiled for 256-bit vectors, taken when the host CPU has AVX2 (checked once)
    static __attribute__((target("avx2"))) void twist_and_temper_avx2(uint32_t *mt, uint32_t *out) { twist_and_temper(mt, out); }
#endif
    static void twist_and_temper_base(uint32_t *mt, uint32_t *out) { twist_and_temper(mt, out); }
    void refill() {
#if defined(__x86_64__) && defined(__GNUC__)
        static const bool avx2 = __builtin_cpu_supports("avx2");
        if (avx2) twist_and_temper_avx2(mt_, out_);
        else
#endif
            twist_and_temper_base(mt_, out_);
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
            int avail;
            const uint32_t *blk = rng.block(&avail);
            int used = 0;
            while (used < avail && i >= stop) {
                const uint32_t r = blk[used++] >> sh;
                jarr[i] = r;
                i -= (r < (uint32_t)(i + 1)) ? 1 : 0;
            }
            rng.consume(used);
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
