// pack_rows.h -- host side of the packed wire format of expression rows ("bytes not sent").
//
// End to end the hot path is bound by PCIe, so the format spends host cycles ONCE (at ingest / conversion time, or in the
// feeder when it has cores to spare) to halve the bytes every epoch ships: 2 bytes per row instead of 4 (gene + value columns
// on the cell-start-index route) or 8 (the full COO row).  It leans on two facts of SLAF's expression table
// (slaf/data/converter.py:2819,3025): rows are cell-major with gene ids ASCENDING inside a cell, and values are small counts.
//
// One block per cell, 16-byte aligned (so a block is ONE TMA bulk copy into the tokenizing kernel's shared-memory stage):
//
//   header   16 B   u32 n_rows | u32 base_gene (gene id of row 0) | u32 n_exc_g | u32 n_exc_v
//   deltas   n_rows bytes, zero-padded to a multiple of 16:  byte i = min(delta_i, 255), delta_0 = 0,
//                                                             delta_i = gene_i - gene_{i-1} - 1   (>= 0: strictly ascending)
//   values   n_rows bytes, zero-padded to a multiple of 16:  byte i = min(value_i, 255)
//   gene exceptions   n_exc_g x { u32 row, u32 delta }    rows whose delta is >= 255, ascending by row, full delta
//   value exceptions  n_exc_v x { u32 row, u32 value }    rows whose value is >= 255, ascending by row, full value
//   zero padding to a multiple of 16
//
// plus, per batch, blk_off[C + 1]: block offsets in 16-byte units.  Decoding is exact (lossless): gene_i = base_gene +
// sum_{j=1..i} (delta_j + 1), value_i = byte or its exception.  A cell whose genes do not ascend strictly cannot be packed
// (SLAFB_ERR_INVALID: the caller ships such a batch unpacked).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>


namespace slafb_pack {

constexpr uint32_t kHeaderBytes = 16;
inline size_t pad16(size_t n) { return (n + 15u) & ~(size_t)15u; }
inline size_t block_bytes(size_t n_rows, size_t n_exc_g, size_t n_exc_v) {
    return pad16(kHeaderBytes + 2 * pad16(n_rows) + 8 * (n_exc_g + n_exc_v));
}

// ---- one cell: the two byte streams, written straight into the block; returns the exception counts ---------------------
// (16-bit arithmetic for uint16 gene ids: the subtract / min / narrow chain then vectorises 16-32 rows per instruction)
template <typename GeneT>
static inline __attribute__((always_inline)) void pack_cell(const GeneT *g, const uint16_t *v, size_t n, uint8_t *dl, uint8_t *vl, uint32_t *exc_g,
                                                            uint32_t *exc_v, uint32_t *bad) {
    uint32_t eg = 0, ev = 0, b = 0;
    if (n) dl[0] = 0;
    if (sizeof(GeneT) == 2) {
        for (size_t i = 1; i < n; i++) {
            const uint16_t a = (uint16_t)g[i], p = (uint16_t)g[i - 1];
            const uint16_t d = (uint16_t)(a - p - 1u);
            b |= (a <= p) ? 1u : 0u;
            eg += (d >= 255u) ? 1u : 0u;
            dl[i] = (uint8_t)(d < 255u ? d : 255u);
        }
    } else {
        for (size_t i = 1; i < n; i++) {
            const int64_t d = (int64_t)g[i] - (int64_t)g[i - 1] - 1;
            b |= (d < 0) ? 1u : 0u;
            eg += (d >= 255) ? 1u : 0u;
            dl[i] = (uint8_t)(d < 255 ? d : 255);
        }
    }
    for (size_t i = 0; i < n; i++) {
        const uint16_t x = v[i];
        ev += (x >= 255u) ? 1u : 0u;
        vl[i] = (uint8_t)(x < 255u ? x : 255u);
    }
    *exc_g = eg; *exc_v = ev; *bad |= b;
}

template <typename GeneT>
static inline __attribute__((always_inline)) int pack_body(const GeneT *gene, const uint16_t *value, const uint32_t *seg, size_t n_cells, uint8_t *out,
                                                           size_t out_capacity, uint32_t *blk_off, size_t *out_bytes) {
    size_t off = 0;
    uint32_t bad = 0;
    for (size_t c = 0; c < n_cells; c++) {
        const GeneT *g = gene + seg[c];
        const uint16_t *v = value + seg[c];
        const size_t n = (size_t)(seg[c + 1] - seg[c]), np = pad16(n);
        if (off + kHeaderBytes + 2 * np > out_capacity) return -3;
        if ((off >> 4) > 0xffffffffull) return -3;
        blk_off[c] = (uint32_t)(off >> 4);
        uint8_t *blk = out + off, *dl = blk + kHeaderBytes, *vl = dl + np;
        uint32_t eg, ev;
        pack_cell(g, v, n, dl, vl, &eg, &ev, &bad);
        if (bad) return -1;                         // gene ids do not ascend: the exception counts of this cell mean nothing
        memset(dl + n, 0, np - n);
        memset(vl + n, 0, np - n);
        const size_t total = block_bytes(n, eg, ev);
        if (off + total > out_capacity) return -3;
        uint32_t hdr[4] = {(uint32_t)n, n ? (uint32_t)g[0] : 0u, eg, ev};
        memcpy(blk, hdr, 16);
        uint32_t *ex = reinterpret_cast<uint32_t *>(vl + np);
        if (eg) {                                   // rare: the rows again, for the full deltas
            for (size_t i = 1; i < n; i++) {
                const uint32_t d = (uint32_t)g[i] - (uint32_t)g[i - 1] - 1u;
                if (d >= 255u) { *ex++ = (uint32_t)i; *ex++ = d; }
            }
        }
        if (ev) {
            for (size_t i = 0; i < n; i++)
                if (v[i] >= 255) { *ex++ = (uint32_t)i; *ex++ = (uint32_t)v[i]; }
        }
        uint8_t *end = reinterpret_cast<uint8_t *>(ex);
        memset(end, 0, (size_t)(blk + total - end));
        off += total;
    }
    blk_off[n_cells] = (uint32_t)(off >> 4);
    *out_bytes = off;
    return bad ? -1 : 0;
}

// the same loop compiled for AVX2 / AVX-512 as well (the byte packing and the compares vectorise); chosen at run time
template <typename GeneT>
static int pack_base(const GeneT *gene, const uint16_t *value, const uint32_t *seg, size_t n_cells, uint8_t *out, size_t cap, uint32_t *blk_off, size_t *nb) {
    return pack_body(gene, value, seg, n_cells, out, cap, blk_off, nb);
}
#if defined(__x86_64__)
template <typename GeneT>
static __attribute__((target("avx2"))) int pack_avx2(const GeneT *gene, const uint16_t *value, const uint32_t *seg, size_t n_cells, uint8_t *out, size_t cap,
                                                     uint32_t *blk_off, size_t *nb) {
    return pack_body(gene, value, seg, n_cells, out, cap, blk_off, nb);
}
template <typename GeneT>
static __attribute__((target("avx512f,avx512bw,avx512vl"))) int pack_avx512(const GeneT *gene, const uint16_t *value, const uint32_t *seg, size_t n_cells,
                                                                            uint8_t *out, size_t cap, uint32_t *blk_off, size_t *nb) {
    return pack_body(gene, value, seg, n_cells, out, cap, blk_off, nb);
}
#endif

inline int isa_level() {
#if defined(__x86_64__)
    static const int level = [] {
        if (__builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl")) return 2;
        if (__builtin_cpu_supports("avx2")) return 1;
        return 0;
    }();
    return level;
#else
    return 0;
#endif
}

// Packs cells [0, n_cells) of the columns in ONE pass over the rows (single thread: callers pack disjoint cell ranges in parallel
// and hand the results over as separate pieces).  Returns 0, -1 (genes not strictly ascending in some cell) or -3 (out too small).
// blk_off[n_cells + 1] receives the block offsets (16-byte units, relative to `out`); *out_bytes the bytes written.
template <typename GeneT>
int pack(const GeneT *gene, const uint16_t *value, const uint32_t *seg, size_t n_cells, uint8_t *out, size_t out_capacity, uint32_t *blk_off,
         size_t *out_bytes) {
#if defined(__x86_64__)
    const int l = isa_level();
    if (l == 2) return pack_avx512(gene, value, seg, n_cells, out, out_capacity, blk_off, out_bytes);
    if (l == 1) return pack_avx2(gene, value, seg, n_cells, out, out_capacity, blk_off, out_bytes);
#endif
    return pack_base(gene, value, seg, n_cells, out, out_capacity, blk_off, out_bytes);
}

// Reference decoder (host): the exact inverse, used by tests and by callers that want the columns back.
template <typename GeneT>
int unpack(const uint8_t *packed, const uint32_t *blk_off, size_t n_cells, const uint32_t *seg, GeneT *gene, uint16_t *value) {
    for (size_t c = 0; c < n_cells; c++) {
        const uint8_t *blk = packed + (size_t)blk_off[c] * 16u;
        uint32_t hdr[4];
        memcpy(hdr, blk, 16);
        const size_t n = hdr[0];
        if (n != (size_t)(seg[c + 1] - seg[c])) return -1;
        const uint8_t *dl = blk + kHeaderBytes, *vl = dl + pad16(n);
        const uint32_t *ex = reinterpret_cast<const uint32_t *>(vl + pad16(n));
        const uint32_t *exg = ex, *exv = ex + 2 * (size_t)hdr[2];
        size_t ig = 0, iv = 0;
        uint32_t g = hdr[1];
        for (size_t i = 0; i < n; i++) {
            uint32_t d = dl[i];
            if (d == 255u) { if (ig >= hdr[2] || exg[2 * ig] != i) return -1; d = exg[2 * ig + 1]; ig++; }
            if (i) g += d + 1u;
            uint32_t x = vl[i];
            if (x == 255u) { if (iv >= hdr[3] || exv[2 * iv] != i) return -1; x = exv[2 * iv + 1]; iv++; }
            gene[seg[c] + i] = (GeneT)g;
            value[seg[c] + i] = (uint16_t)x;
        }
    }
    return 0;
}

}  // namespace slafb_pack
