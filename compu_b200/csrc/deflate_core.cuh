// deflate_core.cuh — the encoder's per-position / per-segment / per-block algorithms as __host__ __device__ functions.
//
// What compu outsources to L0 `deflate()` behind encode_fn (/root/reference/src/encoder/zlib_ng.rs:90-92 ->
// src/encoder/mod.rs:334-370): LZ77 match finding, lazy parse, block split, dynamic Huffman construction and the
// bit-level block format of RFC 1951. The GPU kernels (deflate_kernels.cuh) call these functions from one thread per
// position (match search), one thread per segment (parse) and one thread per block (Huffman); tests/model compiles the
// SAME functions for the host, so the CUDA output can be checked byte for byte against a sequential run of the same
// decisions. (The model is test infrastructure; the product has no CPU path.)
#pragma once
#include "czk_common.cuh"

namespace czk {

#define CZK_WINDOW 32768u
#define CZK_MIN_MATCH 3u
#define CZK_MAX_MATCH 258u
#ifndef CZK_HASH_BITS
#define CZK_HASH_BITS 13
#endif
#define CZK_BLOCK_TOKENS 16384u   // tokens per deflate block (zlib memLevel 8: lit_bufsize)

struct DeflateTuning {
    uint32_t max_chain;    // candidates examined per position
    uint32_t nice_len;     // stop searching at this length
    uint32_t lazy;         // 1: one-step lazy evaluation (zlib levels >= 4), 0: greedy
    uint32_t min_len_far;  // matches of length 3 farther than 4096 are dropped (zlib TOO_FAR)
    uint32_t huffman_only; // strategy HuffmanOnly: no matches at all
    uint32_t rle_only;     // strategy Rle: distance-1 matches only
    uint32_t fixed_only;   // strategy Fixed: never emit dynamic blocks
    uint32_t level0;       // stored blocks only
};

__host__ __device__ inline DeflateTuning deflate_tuning(int level, int strategy) {
    DeflateTuning t;
    if (level < 0) level = 6;
    if (level > 9) level = 9;
    // chain depth / nice length per level, shaped after zlib's configuration_table (deflate.c)
    // Level 6 is the benchmarked level: (chain, nice) swept with tools/sweep_deflate_ratio.py (size against zlib 1.3 level 6) and
    // timed on a B200 (profiles/r2_notes.md): chain 16 / 12 / 10 / 8 -> 91.9 / 84.1 / 79.8 / 75.2 ms per GiB at 1.005 / 1.010 /
    // 1.013 / 1.018 x zlib's size on Markov text; 10 is the fastest point within 1.5 % on all three synthetic classes.
    const uint32_t chain[10] = {0, 2, 3, 4, 6, 8, 10, 24, 48, 128};
    const uint32_t nice[10] = {0, 8, 16, 32, 16, 32, 64, 128, 258, 258};
    t.max_chain = chain[level];
    t.nice_len = nice[level];
#ifdef CZK_SWEEP_CHAIN  // tests/model ratio sweeps only
    t.max_chain = CZK_SWEEP_CHAIN; t.nice_len = CZK_SWEEP_NICE;
#endif
    t.lazy = level >= 4;
    t.min_len_far = 1;
    t.huffman_only = strategy == 2;
    t.rle_only = strategy == 3;
    t.fixed_only = strategy == 4;
    t.level0 = level == 0;
    if (strategy == 1) { t.max_chain = t.max_chain > 8 ? 8 : t.max_chain; }  // Filtered: shorter searches
    return t;
}

__host__ __device__ inline uint32_t load32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// Little-endian 32-bit load from any byte address. On the device: two ALIGNED word loads and a funnel shift (an aligned
// word is only touched when it holds at least one of the four bytes, so nothing outside the containing words is read).
__host__ __device__ inline uint32_t load32u(const uint8_t *p) {
#ifdef __CUDA_ARCH__
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    const uint32_t lo = __ldg(w);  // (the encoder only ever reads its input: the non-coherent path is safe)
    const uint32_t hi = sh ? __ldg(w + 1) : 0u;
    return __funnelshift_r(lo, hi, sh);
#else
    return load32(p);
#endif
}
__host__ __device__ inline uint32_t ctz32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__ffs((int)x) - 1u;
#else
    return (uint32_t)__builtin_ctz(x);
#endif
}

// hash of the 4 bytes at p
__host__ __device__ inline uint32_t hash4(uint32_t v) { return (v * 2654435761u) >> (32 - CZK_HASH_BITS); }

#ifdef CZK_COUNT_STEPS
static unsigned long long czk_step_count = 0;
#endif
// State of one find_match() call (the current position's first 16 bytes stay in registers: most matches end inside them).
struct MatchScan {
    const uint8_t *cur;
    uint32_t max_len, cur4, curw1, curw2, curw3;
    uint32_t best_len, best_dist, cur_end;  // cur_end = cur[best_len] once there is a match (best_len < max_len then)
};

// Examine one candidate `total` bytes back, whose first word (c4) and byte at best_len (cb) are already loaded.
// Returns true when the search is over (nice length or the longest possible match).
// quick rejects (zlib's order): the byte that would extend the best match, then the first word
__host__ __device__ inline bool match_passes(const MatchScan &m, uint32_t c4, uint32_t cb) {
    return (m.best_len < 4 || cb == m.cur_end) && c4 == m.cur4;
}

// Extend a candidate that passed the quick rejects and update the best match. True when the search is over.
__host__ __device__ inline bool match_extend(MatchScan &m, uint32_t total, const DeflateTuning &t, uint32_t &chain) {
    const uint8_t *cand = m.cur - total;
    const uint32_t max_len = m.max_len;
    uint32_t l = 4;
    bool mism = false;  // stopped at a mismatching word (l is final)
    do {
        // words 1..3 of the current position come from registers; every exit leaves the few lanes that got here early
        if (8 > max_len) break;
        uint32_t x = load32u(cand + 4) ^ m.curw1;
        if (x) { l = 4 + (ctz32(x) >> 3); mism = true; break; }
        l = 8;
        if (12 > max_len) break;
        x = load32u(cand + 8) ^ m.curw2;
        if (x) { l = 8 + (ctz32(x) >> 3); mism = true; break; }
        l = 12;
        if (16 > max_len) break;
        x = load32u(cand + 12) ^ m.curw3;
        if (x) { l = 12 + (ctz32(x) >> 3); mism = true; break; }
        l = 16;
        while (l + 4 <= max_len) {
            x = load32u(cand + l) ^ load32u(m.cur + l);
            if (x) { l += ctz32(x) >> 3; mism = true; break; }
            l += 4;
        }
    } while (0);
    if (!mism)  // fewer than 4 bytes left to compare (or ran to the end): finish byte-wise
        while (l < max_len && cand[l] == m.cur[l]) l++;
    if (l > m.best_len) {
        m.best_len = l;
        m.best_dist = total;
        if (l >= t.nice_len || l == max_len) return true;
        m.cur_end = m.cur[l];
#ifdef CZK_GOOD_LEN  // experiment (tests/model sweeps): zlib's good_length rule
        if (l >= CZK_GOOD_LEN) chain >>= 2;
#endif
    }
    return false;
}

__host__ __device__ inline bool match_try(MatchScan &m, uint32_t total, uint32_t c4, uint32_t cb, const DeflateTuning &t, uint32_t &chain) {
    return match_passes(m, c4, cb) && match_extend(m, total, t, chain);
}

// Best match for position `pos` of a segment: walk the chain of earlier positions with the same 4-byte hash
// (prevd[i] = distance from i to the previous such position, 0 = none), newest first. Returns len | dist << 9, or 0.
// Matches never cross the end of the segment and never look farther back than the segment start. A candidate must agree
// on the first 4 bytes (the hash is over 4 bytes, so the chain holds little else); comparison is word-wise.
// prevd2 (optional): prevd2[i] = distance from i to the SECOND previous position of its chain (0 = none or beyond the
// window). The walk is a chain of dependent loads; with the second link a step fetches two candidates and the links of the
// node after them in one memory round trip. Same candidates, same order, same result as the single-link walk.
__host__ __device__ inline uint32_t find_match(const uint8_t *seg, uint32_t seg_len, const uint16_t *prevd, uint32_t pos,
                                               const DeflateTuning &t, const uint16_t *prevd2 = nullptr) {
    if (pos + CZK_MIN_MATCH > seg_len) return 0;
    uint32_t max_len = seg_len - pos;
    if (max_len > CZK_MAX_MATCH) max_len = CZK_MAX_MATCH;
    if (t.huffman_only) return 0;
    const uint8_t *cur = seg + pos;
    if (t.rle_only) {
        if (pos == 0) return 0;
        uint32_t l = 0;
        while (l < max_len && cur[l] == cur[-1]) l++;
        return l >= CZK_MIN_MATCH ? (l | (1u << 9)) : 0;
    }
    if (max_len < 4) return 0;  // the chains are built from 4-byte hashes: the last 3 positions have no link
    MatchScan m;
    m.cur = cur; m.max_len = max_len;
    m.cur4 = load32u(cur);
    m.curw1 = max_len >= 8 ? load32u(cur + 4) : 0u;
    m.curw2 = max_len >= 12 ? load32u(cur + 8) : 0u;
    m.curw3 = max_len >= 16 ? load32u(cur + 12) : 0u;
    m.best_len = 0; m.best_dist = 0; m.cur_end = 0;
    uint32_t total = 0, chain = t.max_chain;
    if (prevd2) {
        uint32_t d1 = prevd[pos], d2 = prevd2[pos];
        while (d1 && chain) {
            const uint32_t ta = total + d1;
            if (ta > CZK_WINDOW || ta > pos) break;
            const uint32_t tb = total + d2;
            const bool hasb = d2 != 0 && tb <= CZK_WINDOW && tb <= pos;  // (d2 > d1 whenever it is not 0)
            // everything this step needs is requested before anything is used
            const uint32_t c4a = load32u(cur - ta);
            const uint32_t cba = m.best_len >= 4 ? *(cur - ta + m.best_len) : 0u;
            uint32_t c4b = 0, cbb = 0, n1 = 0, n2 = 0;
            if (hasb) {
                c4b = load32u(cur - tb);
                cbb = m.best_len >= 4 ? *(cur - tb + m.best_len) : 0u;
                n1 = prevd[pos - tb];
                n2 = prevd2[pos - tb];
            }
#ifdef CZK_COUNT_STEPS
            czk_step_count++;
#endif
            chain--;
            const uint32_t bl0 = m.best_len;
            if (match_try(m, ta, c4a, cba, t, chain)) break;
            if (!hasb || !chain) break;  // the chain ends after the first candidate (no further link / budget used up)
#ifdef CZK_COUNT_STEPS
            czk_step_count++;
#endif
            chain--;
            if (m.best_len != bl0) cbb = *(cur - tb + m.best_len);  // the first candidate moved the goal
            if (match_try(m, tb, c4b, cbb, t, chain)) break;
            total = tb;
            d1 = n1;
            d2 = n2;
        }
    } else {
        uint32_t d = prevd[pos];
        while (d && chain--) {
#ifdef CZK_COUNT_STEPS
            czk_step_count++;
#endif
            total += d;
            if (total > CZK_WINDOW || total > pos) break;
            // the next link, the candidate's first word and the byte that would extend the best match are requested together:
            // the walk is a chain of dependent loads, and this way one step costs one memory round trip instead of three
            const uint32_t dnext = prevd[pos - total];
            const uint32_t c4 = load32u(cur - total);
            const uint32_t cb = m.best_len >= 4 ? *(cur - total + m.best_len) : 0u;
            if (match_try(m, total, c4, cb, t, chain)) break;
            d = dnext;
        }
    }
    const uint32_t best_len = m.best_len, best_dist = m.best_dist;
    if (best_len < CZK_MIN_MATCH) return 0;
    return best_len | (best_dist << 9);
}

#if (defined(__CUDACC__) || defined(CUSIM)) && defined(CZ_EXPERIMENTS)
// find_match() for 32 positions at once, called by all lanes of a warp (lanes without a position pass valid = false).
// Same candidates in the same order with the same rules per lane, but the warp alternates between two phases instead of
// letting every lane run its own loop: WALK — lanes follow their chains until each has found a candidate that passes the
// quick rejects (or has finished); EXTEND — all those candidates are compared at once. In the per-lane loop the comparison
// code ran for ~4 lanes at a time (profiles/r1_deflate_match_sweep_ncu.md: 41 % of the kernel's instructions at 4 of 32
// lanes), because lanes reach a passing candidate at different steps.
// MEASURED SLOWER (139 ms against 95 ms per GiB, CZ_MATCH_V=6): lanes that have found a candidate wait for the slowest
// walker of the warp, and that costs more than the batched comparison saves. Kept as an experiment, not the default.
__device__ inline uint32_t find_match_warp(const uint8_t *seg, uint32_t seg_len, const uint16_t *prevd, uint32_t pos,
                                           const DeflateTuning &t, bool valid) {
    uint32_t max_len = 0;
    if (valid && pos + CZK_MIN_MATCH <= seg_len) { max_len = seg_len - pos; if (max_len > CZK_MAX_MATCH) max_len = CZK_MAX_MATCH; }
    const uint8_t *cur = seg + pos;
    MatchScan m;
    m.cur = cur; m.max_len = max_len;
    m.cur4 = m.curw1 = m.curw2 = m.curw3 = 0;
    m.best_len = 0; m.best_dist = 0; m.cur_end = 0;
    bool walking = max_len >= 4;
    uint32_t total = 0, d = 0, chain = t.max_chain;
    if (walking) {
        m.cur4 = load32u(cur);
        if (max_len >= 8) m.curw1 = load32u(cur + 4);
        if (max_len >= 12) m.curw2 = load32u(cur + 8);
        if (max_len >= 16) m.curw3 = load32u(cur + 12);
        d = prevd[pos];
    }
    for (;;) {
        bool pend = false;
        uint32_t pend_total = 0;
        while (__any_sync(0xffffffffu, walking)) {
            if (walking) {
                if (d && chain) {
                    chain--;
                    total += d;
                    if (total > CZK_WINDOW || total > pos) walking = false;
                    else {
                        const uint32_t dnext = prevd[pos - total];
                        const uint32_t c4 = load32u(cur - total);
                        const uint32_t cb = m.best_len >= 4 ? *(cur - total + m.best_len) : 0u;
                        if (match_passes(m, c4, cb)) { pend = true; pend_total = total; walking = false; }
                        d = dnext;
                    }
                } else walking = false;
            }
        }
        if (!__any_sync(0xffffffffu, pend)) break;
        if (pend) walking = !match_extend(m, pend_total, t, chain);
    }
    if (m.best_len < CZK_MIN_MATCH) return 0;
    return m.best_len | (m.best_dist << 9);
}
#endif

// ------------------------------------------------------------------------------------------------------------------
// Tokens: literal = byte << 9 (len field 0), match = len | dist << 9.
__host__ __device__ inline uint32_t len_code(uint32_t len) {  // length 3..258 -> symbol 257..285
    uint32_t l = len - 3;
    if (l < 8) return 257 + l;
    if (len == 258) return 285;
    uint32_t e = 29 - (uint32_t)
#ifdef __CUDA_ARCH__
                          __clz((int)l);
#else
                          __builtin_clz(l);
#endif
    // e = floor(log2(l)) - 2  (extra bits), l >= 8
    return 257 + 4 * e + 4 + ((l >> e) & 3);
}
__host__ __device__ inline uint32_t len_extra_bits(uint32_t sym) { return sym < 265 || sym == 285 ? 0 : (sym - 261) >> 2; }
__host__ __device__ inline uint32_t len_base(uint32_t sym) {
    if (sym < 265) return sym - 254;
    if (sym == 285) return 258;
    uint32_t e = (sym - 261) >> 2;
    return 3 + ((4 + ((sym - 265) & 3)) << e);
}
__host__ __device__ inline uint32_t dist_code(uint32_t dist) {  // distance 1..32768 -> symbol 0..29
    uint32_t d = dist - 1;
    if (d < 4) return d;
    uint32_t e = 30 - (uint32_t)
#ifdef __CUDA_ARCH__
                          __clz((int)d);
#else
                          __builtin_clz(d);
#endif
    // e = floor(log2(d)) - 1
    return 2 * e + 2 + ((d >> e) & 1);
}
__host__ __device__ inline uint32_t dist_extra_bits(uint32_t sym) { return sym < 4 ? 0 : (sym >> 1) - 1; }
__host__ __device__ inline uint32_t dist_base(uint32_t sym) {
    if (sym < 4) return sym + 1;
    uint32_t e = (sym >> 1) - 1;
    return ((2 + (sym & 1)) << e) + 1;
}

// ------------------------------------------------------------------------------------------------------------------
// Sequential parse of one segment from the per-position matches (one-step lazy evaluation, as deflate_slow does in
// spirit: a match at p is deferred by one literal when the match at p+1 is strictly longer). Tokens are written in place
// over the match array (token k never lands beyond the position being read). Block boundaries every CZK_BLOCK_TOKENS
// tokens: blk_in_end[b] = input offset where block b ends, returns the number of tokens; *n_blocks is set.
__host__ __device__ inline uint32_t parse_segment(const uint8_t *seg, uint32_t seg_len, uint32_t *match_tok, const DeflateTuning &t,
                                                  uint32_t *blk_in_end, uint32_t max_blocks, uint32_t *n_blocks) {
    uint32_t p = 0, nt = 0, nb = 0;
    while (p < seg_len) {
        uint32_t m = match_tok[p];
        uint32_t len = m & 0x1ff;
        uint32_t tok, adv;
        if (len >= CZK_MIN_MATCH) {
            uint32_t nlen = (t.lazy && p + 1 < seg_len) ? (match_tok[p + 1] & 0x1ff) : 0;
            if (nlen > len) { tok = (uint32_t)seg[p] << 9; adv = 1; }
            else { tok = m; adv = len; }
        } else { tok = (uint32_t)seg[p] << 9; adv = 1; }
        match_tok[nt++] = tok;  // nt <= p always: tokens never overtake the position being read
        p += adv;
        if (nt % CZK_BLOCK_TOKENS == 0 && nb < max_blocks) blk_in_end[nb++] = p;
    }
    if (nt % CZK_BLOCK_TOKENS != 0 && nb < max_blocks) blk_in_end[nb++] = seg_len;  // last, partial block
    *n_blocks = nb;
    return nt;
}

// ------------------------------------------------------------------------------------------------------------------
// Length-limited Huffman code lengths for n symbols (n <= 288) with maximum length `maxbits`.
// Deterministic and identical on host and device: symbols sorted by (freq, index), two-queue Huffman merge, depths from
// parent links, then the overflow repair of zlib's gen_bitlen (move the deepest overflowing leaves up).
// A single used symbol gets length 1 (and, like zlib, a second dummy symbol is given a code so the set is complete).
struct HuffScratch {
    uint16_t order[288];    // symbol indices sorted by (freq, sym)
    uint32_t nodef[576];    // frequencies of leaves (sorted order) then internal nodes
    uint16_t parent[576];
    uint8_t depth[576];
    uint16_t bl_count[16];
};

__host__ __device__ inline void huff_build_lengths(const uint32_t *freq_in, uint32_t n, uint32_t maxbits, uint8_t *lens,
                                                   HuffScratch &s, uint32_t force_two) {
    uint32_t used = 0;
    uint32_t freq_local[2];
    (void)freq_local;
    for (uint32_t i = 0; i < n; i++) { lens[i] = 0; if (freq_in[i]) s.order[used++] = (uint16_t)i; }
    // zlib forces at least two codes of non-zero frequency so that the tree is complete
    uint32_t dummy[2] = {0xffffffffu, 0xffffffffu};
    uint32_t nd = 0;
    if (force_two) {
        for (uint32_t cand = 0; used + nd < 2 && cand < n; cand++)
            if (!freq_in[cand]) dummy[nd++] = cand;
    }
    // sort by (freq, sym): keys are unique, so any correct sort gives the same order; shell sort keeps it O(n^1.3)
    {
        const uint32_t gaps[6] = {132, 57, 23, 10, 4, 1};
        for (int g = 0; g < 6; g++) {
            uint32_t gap = gaps[g];
            for (uint32_t i = gap; i < used; i++) {
                uint16_t v = s.order[i];
                uint32_t kv = (freq_in[v] << 9) | v;
                uint32_t j = i;
                while (j >= gap) {
                    uint16_t w = s.order[j - gap];
                    if (((freq_in[w] << 9) | w) <= kv) break;
                    s.order[j] = w;
                    j -= gap;
                }
                s.order[j] = v;
            }
        }
    }
    // leaves: dummies first (frequency 1, as zlib does), then the used symbols ascending
    uint32_t nleaf = used + nd;
    if (nleaf == 0) return;
    if (nleaf == 1) { lens[nd ? dummy[0] : s.order[0]] = 1; return; }
    // shift order to make room for dummies at the front
    for (uint32_t i = used; i-- > 0;) s.order[i + nd] = s.order[i];
    for (uint32_t i = 0; i < nd; i++) s.order[i] = (uint16_t)dummy[i];
    for (uint32_t i = 0; i < nleaf; i++) s.nodef[i] = i < nd ? 1u : freq_in[s.order[i]];
    // dummies (freq 1) may be larger than real freq-1 symbols? equal, so order stays non-decreasing: real freqs >= 1.
    uint32_t qa = 0, qb = nleaf, next = nleaf;  // leaf queue [qa, nleaf), internal queue [qb, next)
    while ((nleaf - qa) + (next - qb) > 1) {
        uint32_t pick[2];
        for (int k = 0; k < 2; k++) {
            bool take_leaf;
            if (qa >= nleaf) take_leaf = false;
            else if (qb >= next) take_leaf = true;
            else take_leaf = s.nodef[qa] <= s.nodef[qb];
            pick[k] = take_leaf ? qa++ : qb++;
        }
        s.nodef[next] = s.nodef[pick[0]] + s.nodef[pick[1]];
        s.parent[pick[0]] = (uint16_t)next;
        s.parent[pick[1]] = (uint16_t)next;
        next++;
    }
    uint32_t root = next - 1;
    s.depth[root] = 0;
    for (uint32_t b = 0; b <= 15; b++) s.bl_count[b] = 0;
    // depths top-down with clamping, exactly like zlib's gen_bitlen: EVERY node (internal ones too) that would sit below
    // maxbits is clamped and counted, so that overflow / 2 is the number of repair steps that restores the Kraft sum
    int overflow = 0;
    for (uint32_t i = root; i-- > 0;) {
        uint32_t d = (uint32_t)s.depth[s.parent[i]] + 1;
        if (d > maxbits) { d = maxbits; overflow++; }
        s.depth[i] = (uint8_t)d;
        if (i < nleaf) s.bl_count[d]++;
    }
    if (overflow > 0) {
        // zlib gen_bitlen: repair the Kraft sum by moving leaves
        do {
            uint32_t bits = maxbits - 1;
            while (s.bl_count[bits] == 0) bits--;
            s.bl_count[bits]--;
            s.bl_count[bits + 1] += 2;
            s.bl_count[maxbits]--;
            overflow -= 2;
        } while (overflow > 0);
        // reassign lengths: leaves in order of decreasing frequency get the shortest lengths
        uint32_t i = nleaf;
        for (uint32_t bits = 1; bits <= maxbits; bits++) {
            uint32_t c = s.bl_count[bits];
            while (c--) { i--; s.depth[i] = (uint8_t)bits; }
        }
    }
    for (uint32_t i = 0; i < nleaf; i++) lens[s.order[i]] = s.depth[i];
}

// canonical codes, bit-reversed so they can be OR-ed into an LSB-first bit stream
__host__ __device__ inline void huff_codes(const uint8_t *lens, uint32_t n, uint16_t *codes) {
    uint32_t bl_count[16];
    for (int b = 0; b < 16; b++) bl_count[b] = 0;
    for (uint32_t i = 0; i < n; i++) bl_count[lens[i]]++;
    bl_count[0] = 0;
    uint32_t next[16], code = 0;
    next[0] = 0;
    for (int b = 1; b < 16; b++) { code = (code + bl_count[b - 1]) << 1; next[b] = code; }
    for (uint32_t i = 0; i < n; i++) {
        uint32_t l = lens[i];
        if (!l) { codes[i] = 0; continue; }
        uint32_t c = next[l]++;
        uint32_t r = 0;
        for (uint32_t k = 0; k < l; k++) r |= ((c >> k) & 1u) << (l - 1 - k);
        codes[i] = (uint16_t)r;
    }
}

// Per-block plan produced by plan_block(): everything the emit kernel needs.
struct BlockPlan {
    uint8_t lit_len[288];
    uint8_t dist_len[32];
    uint16_t lit_code[288];
    uint16_t dist_code[32];
    // dynamic header, pre-rendered: up to 3+5+5+4 + 19*3 + 316*(7+7) bits < 4600 bits
    uint32_t hdr_bits;        // number of bits in hdr[]
    uint32_t hdr[160];        // header bits LSB-first (includes BFINAL=0/BTYPE)
    uint32_t btype;           // 0 stored, 1 fixed, 2 dynamic
    uint64_t body_bits;       // bits of the token codes + end-of-block (btype 1/2)
    uint32_t tok_begin, tok_end;   // token range inside the segment's token array
    uint32_t in_begin, in_end;     // input byte range of the block inside the segment
    uint64_t bit_off;         // bit offset of the block inside the segment's output (filled by the layout pass)
};

struct BitSink {  // tiny LSB-first writer into a uint32 array (used for headers only)
    uint32_t *w;
    uint32_t nbits;
    __host__ __device__ inline void put(uint32_t v, uint32_t n) {
        if (!n) return;
        uint32_t i = nbits >> 5, sh = nbits & 31;
        w[i] |= v << sh;
        if (sh + n > 32) w[i + 1] |= v >> (32 - sh);
        nbits += n;
    }
};

__host__ __device__ inline void fixed_lengths(uint8_t *lit_len, uint8_t *dist_len) {
    for (uint32_t i = 0; i < 288; i++) lit_len[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
    for (uint32_t i = 0; i < 32; i++) dist_len[i] = 5;
}

// Chooses the block type (stored / fixed / dynamic, smallest wins like zlib's _tr_flush_block) from the block's symbol
// histograms, builds the codes and renders the header. lit_freq[286], dist_freq[30]; lit_freq[256] must be >= 1.
__host__ __device__ inline void plan_block(const uint32_t *lit_freq, const uint32_t *dist_freq, BlockPlan &bp, HuffScratch &hs,
                                           const DeflateTuning &t) {
    uint8_t dl[288 + 32];
    uint32_t in_bytes = bp.in_end - bp.in_begin;
    // ---- dynamic codes
    huff_build_lengths(lit_freq, 286, 15, bp.lit_len, hs, 1);
    huff_build_lengths(dist_freq, 30, 15, bp.dist_len, hs, 1);
    bp.lit_len[286] = bp.lit_len[287] = 0;
    bp.dist_len[30] = bp.dist_len[31] = 0;
    uint32_t nlit = 286, ndist = 30;
    while (nlit > 257 && bp.lit_len[nlit - 1] == 0) nlit--;
    while (ndist > 1 && bp.dist_len[ndist - 1] == 0) ndist--;
    // code length sequence with run-length symbols 16/17/18 (RFC 1951 §3.2.7), greedy like zlib's scan_tree/send_tree
    for (uint32_t i = 0; i < nlit; i++) dl[i] = bp.lit_len[i];
    for (uint32_t i = 0; i < ndist; i++) dl[nlit + i] = bp.dist_len[i];
    uint32_t total = nlit + ndist;
    uint32_t cl_freq[19];
    for (int i = 0; i < 19; i++) cl_freq[i] = 0;
    // first pass: frequencies of the code-length alphabet
    {
        uint32_t i = 0;
        while (i < total) {
            uint32_t v = dl[i], run = 1;
            while (i + run < total && dl[i + run] == v) run++;
            uint32_t r = run;
            if (v == 0) {
                while (r >= 11) { uint32_t k = r > 138 ? 138 : r; cl_freq[18]++; r -= k; }
                if (r >= 3) { cl_freq[17]++; r = 0; }
                cl_freq[0] += r;
            } else {
                cl_freq[v]++; r--;
                while (r >= 3) { uint32_t k = r > 6 ? 6 : r; cl_freq[16]++; r -= k; }
                cl_freq[v] += r;
            }
            i += run;
        }
    }
    uint8_t cl_len[19];
    uint16_t cl_code[19];
    huff_build_lengths(cl_freq, 19, 7, cl_len, hs, 1);
    huff_codes(cl_len, 19, cl_code);
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint32_t ncl = 19;
    while (ncl > 4 && cl_len[order[ncl - 1]] == 0) ncl--;
    uint64_t dyn_hdr_bits = 3 + 5 + 5 + 4 + 3 * ncl;
    for (int i = 0; i < 19; i++) dyn_hdr_bits += (uint64_t)cl_freq[i] * cl_len[i];
    dyn_hdr_bits += 2ull * cl_freq[16] + 3ull * cl_freq[17] + 7ull * cl_freq[18];
    uint64_t dyn_body = 0, fix_body = 0;
    uint8_t fl[288], fd[32];
    fixed_lengths(fl, fd);
    for (uint32_t i = 0; i < 286; i++) {
        uint32_t eb = i >= 257 ? len_extra_bits(i) : 0;
        dyn_body += (uint64_t)lit_freq[i] * (bp.lit_len[i] + eb);
        fix_body += (uint64_t)lit_freq[i] * (fl[i] + eb);
    }
    for (uint32_t i = 0; i < 30; i++) {
        uint32_t eb = dist_extra_bits(i);
        dyn_body += (uint64_t)dist_freq[i] * (bp.dist_len[i] + eb);
        fix_body += (uint64_t)dist_freq[i] * (5 + eb);
    }
    uint64_t dyn_total = dyn_hdr_bits + dyn_body, fix_total = 3 + fix_body;
    // stored: 3 bits + pad (<=7) + 32 bits per 65535-byte piece + data
    uint64_t pieces = in_bytes ? (in_bytes + 65534) / 65535 : 1;
    uint64_t stored_total = pieces * (3 + 7 + 32) + 8ull * in_bytes;
    for (uint32_t i = 0; i < 160; i++) bp.hdr[i] = 0;
    BitSink bs{bp.hdr, 0};
    if (t.level0 || (stored_total <= dyn_total && stored_total <= fix_total)) {
        bp.btype = 0;
        bp.hdr_bits = 0;
        bp.body_bits = 0;
        return;
    }
    if (t.fixed_only || fix_total <= dyn_total) {
        bp.btype = 1;
        for (uint32_t i = 0; i < 288; i++) bp.lit_len[i] = fl[i];
        for (uint32_t i = 0; i < 32; i++) bp.dist_len[i] = fd[i];
        huff_codes(bp.lit_len, 288, bp.lit_code);
        huff_codes(bp.dist_len, 32, bp.dist_code);
        bs.put(0, 1);  // BFINAL = 0: segments never end the stream
        bs.put(1, 2);
        bp.hdr_bits = bs.nbits;
        bp.body_bits = fix_body;
        return;
    }
    bp.btype = 2;
    huff_codes(bp.lit_len, 288, bp.lit_code);
    huff_codes(bp.dist_len, 32, bp.dist_code);
    bs.put(0, 1);
    bs.put(2, 2);
    bs.put(nlit - 257, 5);
    bs.put(ndist - 1, 5);
    bs.put(ncl - 4, 4);
    for (uint32_t i = 0; i < ncl; i++) bs.put(cl_len[order[i]], 3);
    {
        uint32_t i = 0;
        while (i < total) {
            uint32_t v = dl[i], run = 1;
            while (i + run < total && dl[i + run] == v) run++;
            uint32_t r = run;
            if (v == 0) {
                while (r >= 11) { uint32_t k = r > 138 ? 138 : r; bs.put(cl_code[18], cl_len[18]); bs.put(k - 11, 7); r -= k; }
                if (r >= 3) { bs.put(cl_code[17], cl_len[17]); bs.put(r - 3, 3); r = 0; }
                while (r--) bs.put(cl_code[0], cl_len[0]);
            } else {
                bs.put(cl_code[v], cl_len[v]); r--;
                while (r >= 3) { uint32_t k = r > 6 ? 6 : r; bs.put(cl_code[16], cl_len[16]); bs.put(k - 3, 2); r -= k; }
                while (r--) bs.put(cl_code[v], cl_len[v]);
            }
            i += run;
        }
    }
    bp.hdr_bits = bs.nbits;
    bp.body_bits = dyn_body;
}

// bits and bit count of one token under a plan (max 15+5+15+13 = 48 bits)
__host__ __device__ inline uint64_t token_bits(uint32_t tok, const BlockPlan &bp, uint32_t *nbits) {
    uint32_t len = tok & 0x1ff;
    if (len < CZK_MIN_MATCH) {
        uint32_t b = tok >> 9;
        *nbits = bp.lit_len[b];
        return bp.lit_code[b];
    }
    uint32_t dist = tok >> 9;
    uint32_t ls = len_code(len), ds = dist_code(dist);
    uint32_t n = 0;
    uint64_t v = bp.lit_code[ls];
    n = bp.lit_len[ls];
    uint32_t leb = len_extra_bits(ls);
    v |= (uint64_t)(len - len_base(ls)) << n;
    n += leb;
    v |= (uint64_t)bp.dist_code[ds] << n;
    n += bp.dist_len[ds];
    uint32_t deb = dist_extra_bits(ds);
    v |= (uint64_t)(dist - dist_base(ds)) << n;
    n += deb;
    *nbits = n;
    return v;
}

}  // namespace czk
