// deflate_model.cpp — sequential host run of the encoder's decisions (TEST INFRASTRUCTURE, not the product).
// Compiles compu_b200/csrc/deflate_core.cuh for the host and strings the passes together one after another, so that the
// CUDA kernels' output can be compared byte for byte, and so that ratio tuning does not need a GPU.
#define CZK_COUNT_STEPS
#define __host__
#define __device__
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../compu_b200/csrc/deflate_core.cuh"

using namespace czk;

struct BitWriter {
    uint8_t *out;
    uint64_t cap, nbits;
    bool overflow;
    void put(uint64_t v, uint32_t n) {
        while (n) {
            uint64_t byte = nbits >> 3;
            uint32_t sh = nbits & 7, k = 8 - sh < n ? 8 - sh : n;
            if (byte >= cap) { overflow = true; return; }
            if (sh == 0) out[byte] = 0;
            out[byte] |= (uint8_t)((v & ((1u << k) - 1)) << sh);
            v >>= k; n -= k; nbits += k;
        }
    }
    void align() { if (nbits & 7) put(0, 8 - (nbits & 7)); }
};

// prevd[i] = distance to the most recent earlier position with the same 4-byte hash, if <= 32768, else 0
static void build_chains(const uint8_t *seg, uint32_t n, std::vector<uint16_t> &prevd) {
    prevd.assign(n + 1, 0);
    std::vector<int64_t> head(1u << CZK_HASH_BITS, -1);
    for (uint32_t i = 0; i + 4 <= n; i++) {
        uint32_t h = hash4(load32(seg + i));
        int64_t q = head[h];
        if (q >= 0 && i - q <= CZK_WINDOW) prevd[i] = (uint16_t)(i - q == 32768 ? 0 : i - q);
        head[h] = i;
    }
}

extern "C" long model_deflate_segment(const uint8_t *seg, uint32_t n, uint8_t *out, uint64_t cap, int level, int strategy,
                                      uint32_t *prevd_out, uint32_t *match_out, uint64_t *stats) {
    DeflateTuning t = deflate_tuning(level, strategy);
    std::vector<uint16_t> prevd;
    build_chains(seg, n, prevd);
    std::vector<uint32_t> mt(n + 2, 0);
    if (!t.level0)
        for (uint32_t p = 0; p < n; p++) mt[p] = find_match(seg, n, prevd.data(), p, t);
    if (prevd_out) for (uint32_t i = 0; i < n; i++) prevd_out[i] = prevd[i];
    if (match_out) for (uint32_t i = 0; i < n; i++) match_out[i] = mt[i];
    uint32_t blk_end[80], nb = 0;
    uint32_t nt = parse_segment(seg, n, mt.data(), t, blk_end, 80, &nb);
    BitWriter bw{out, cap, 0, false};
    HuffScratch hs;
    uint64_t n_lit = 0, n_match = 0, types[3] = {0, 0, 0};
    for (uint32_t b = 0; b < nb; b++) {
        BlockPlan bp;
        bp.tok_begin = b * CZK_BLOCK_TOKENS;
        bp.tok_end = (b + 1) * CZK_BLOCK_TOKENS < nt ? (b + 1) * CZK_BLOCK_TOKENS : nt;
        bp.in_begin = b ? blk_end[b - 1] : 0;
        bp.in_end = blk_end[b];
        uint32_t lf[286], df[30];
        memset(lf, 0, sizeof lf);
        memset(df, 0, sizeof df);
        for (uint32_t k = bp.tok_begin; k < bp.tok_end; k++) {
            uint32_t tok = mt[k], len = tok & 0x1ff;
            if (len < 3) { lf[tok >> 9]++; n_lit++; }
            else { lf[len_code(len)]++; df[dist_code(tok >> 9)]++; n_match++; }
        }
        lf[256] = 1;
        plan_block(lf, df, bp, hs, t);
        types[bp.btype]++;
        if (bp.btype == 0) {
            uint32_t pos = bp.in_begin;
            do {
                uint32_t k = bp.in_end - pos > 65535 ? 65535 : bp.in_end - pos;
                bw.put(0, 3);
                bw.align();
                bw.put(k, 16);
                bw.put(k ^ 0xffff, 16);
                for (uint32_t i = 0; i < k; i++) bw.put(seg[pos + i], 8);
                pos += k;
            } while (pos < bp.in_end);
        } else {
            for (uint32_t i = 0; i < bp.hdr_bits; i += 32) {
                uint32_t k = bp.hdr_bits - i < 32 ? bp.hdr_bits - i : 32;
                bw.put(bp.hdr[i >> 5], k);
            }
            uint64_t before = bw.nbits;
            for (uint32_t k = bp.tok_begin; k < bp.tok_end; k++) {
                uint32_t nbits;
                uint64_t v = token_bits(mt[k], bp, &nbits);
                bw.put(v, nbits);
            }
            bw.put(bp.lit_code[256], bp.lit_len[256]);
            if (bw.nbits - before != bp.body_bits && !bw.overflow) return -2;  // cost accounting must match emission
        }
    }
    // full-flush marker: empty stored block, byte aligned
    bw.put(0, 3);
    bw.align();
    bw.put(0, 16);
    bw.put(0xffff, 16);
    if (stats) { stats[0] = nt; stats[1] = n_lit; stats[2] = n_match; stats[3] = nb; stats[4] = types[0]; stats[5] = types[1]; stats[6] = types[2]; stats[7] = czk_step_count; czk_step_count = 0; }
    if (bw.overflow) return -1;
    return (long)(bw.nbits >> 3);
}

// length-limited Huffman code lengths of the shared core (for property tests)
extern "C" void model_huff_lengths(const uint32_t *freq, uint32_t n, uint32_t maxbits, uint8_t *lens) {
    static HuffScratch hs;
    huff_build_lengths(freq, n, maxbits, lens, hs, 1);
}
