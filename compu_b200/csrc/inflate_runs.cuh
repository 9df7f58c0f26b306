// inflate_runs.cuh — block-parallel inflate of LONG streams (any producer: zlib, zlib-ng, pigz, ours).
//
// A DEFLATE stream is serial for a decoder: a lane of phase A manages a few MB/s, a whole warp ~20 MB/s. But the blocks of a
// stream are independent for the HUFFMAN part once their first bit is known, and phase A's tokens do not depend on history.
// So a long stream is cut into RUNS of blocks and every run goes through the two-phase path like a small stream of its own:
//
//   1. inflate_candidates_kernel   one warp per chunk of compressed bytes: the first bit offset in the chunk at which a
//                                  complete, valid dynamic-block header parses (BFINAL = 0, BTYPE = 10, HLIT / HDIST in range,
//                                  a COMPLETE code-length code, code lengths that decode to complete literal/length and
//                                  distance codes with an end-of-block symbol). Lane = bit offset, 32 offsets per step.
//   2. phase A in RUN mode         (inflate_tok_kernel, TwoPhaseParams::runs) from every candidate to the first block boundary
//                                  at or after the next candidate — first only counting (sizes, end positions), then, with
//                                  the output and token offsets known, emitting tokens. The host checks that run r ends
//                                  exactly where run r+1 starts; a candidate that is not on the chain is a false positive and
//                                  its range is decoded again from the true boundary (inflate.cu).
//   3. inflate_lz16_kernel         phase B per run into 16-bit symbols: a byte, or a MARKER 0x8000 | w for a byte that lies
//                                  before the run, w = its offset in the 32 KiB window that precedes the run. Copies of
//                                  markers copy the marker, so after this pass every position names its final source.
//   4. inflate_window_kernel       per stream, run after run: the last 32 KiB of run r with its markers replaced through
//                                  the window of run r-1 — 32 K independent look-ups per run, the only serial step.
//   5. inflate_resolve_kernel      every run in parallel: symbols -> bytes through the window in front of the run, 16-byte
//                                  stores into the final output; Adler-32 / CRC-32 per run, folded by the host with the
//                                  combine identities and compared with the container trailer.
//
// The approach follows the published two-stage scheme of parallel gzip decompressors (candidate block search + marker
// replacement); the kernels and the chain verification are written for this library.
#pragma once
#include "inflate_two_phase.cuh"

namespace czk {

// ---------------------------------------------------------------------------------------------------------------
// 1. candidate search
struct CandChunk {
    uint64_t lo_bit, hi_bit;   // search [lo_bit, hi_bit), absolute bit indices into `in`
    uint64_t end_bit;          // end of the stream the chunk belongs to (nothing may be read as data beyond it)
};

// n bits (n <= 25) at absolute bit position pos of `in` (two aligned word loads; the caller keeps pos + n <= end)
__device__ __forceinline__ uint32_t cand_bits(const uint32_t *words, uint64_t pos, uint32_t n) {
    const uint64_t wi = pos >> 5;
    const uint32_t sh = (uint32_t)(pos & 31);
    const uint32_t lo = __ldg(words + wi), hi = __ldg(words + wi + 1);
    return __funnelshift_r(lo, hi, sh) & ((1u << n) - 1u);
}

// 64 bits at absolute bit position pos (three aligned word loads)
__device__ __forceinline__ uint64_t cand_bits64(const uint32_t *words, uint64_t pos) {
    const uint64_t wi = pos >> 5;
    const uint32_t sh = (uint32_t)(pos & 31);
    const uint32_t w0 = __ldg(words + wi), w1 = __ldg(words + wi + 1), w2 = __ldg(words + wi + 2);
    return (uint64_t)__funnelshift_r(w0, w1, sh) | ((uint64_t)__funnelshift_r(w1, w2, sh) << 32);
}

// Kraft weight of three 3-bit code lengths (7-bit code-length code: a length l weighs 128 >> l), for the 512-entry table
__host__ __device__ inline uint32_t cand_kraft9(uint32_t x) {
    uint32_t k = 0;
    for (int f = 0; f < 3; f++) { const uint32_t l = (x >> (3 * f)) & 7; if (l) k += 128u >> l; }
    return k;
}

// The decode table of a COMPLETE code-length code, built by the whole warp into tab[128] (entry = symbol << 3 | length, indexed
// by the next 7 input bits): lane s < 19 owns symbol s — its length comes from the header's 3-bit fields (`clbits`, transmission
// order), the per-length counts from ballots, its canonical code from its rank among the symbols of its length — and fills the
// 2^(7 - length) entries of its code. (The survivors of the second filter are rare, so one lane used to build this alone: ~650
// of the ~950 instructions of a full check, profiles/r2_runs_gzip_cfg4_ncu.md.)
__device__ __forceinline__ void cand_build_cl_tab_warp(uint8_t *tab, uint64_t clbits, uint32_t hclen, uint32_t lane) {
    // transmission index of symbol s (inverse of 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15), 5 bits per symbol
    const uint64_t inv_lo = 3ull | (17ull << 5) | (15ull << 10) | (13ull << 15) | (11ull << 20) | (9ull << 25) | (7ull << 30) |
                            (5ull << 35) | (4ull << 40) | (6ull << 45) | (8ull << 50) | (10ull << 55);      // symbols 0..11
    const uint64_t inv_hi = 12ull | (14ull << 5) | (16ull << 10) | (18ull << 15) | (0ull << 20) | (1ull << 25) | (2ull << 30);  // 12..18
    uint32_t l = 0;
    if (lane < 19) {
        const uint32_t idx = lane < 12 ? (uint32_t)(inv_lo >> (5 * lane)) & 31 : (uint32_t)(inv_hi >> (5 * (lane - 12))) & 31;
        if (idx < hclen) l = (uint32_t)(clbits >> (3 * idx)) & 7;
    }
    uint32_t code = 0, my_code = 0;
#pragma unroll
    for (uint32_t len = 1; len <= 7; len++) {
        const uint32_t m = __ballot_sync(CZK_FULL, l == len);
        if (l == len) my_code = code + __popc(m & ((1u << lane) - 1u));
        code = (code + __popc(m)) << 1;
    }
    if (l) {
        const uint32_t rev = __brev(my_code) >> (32 - l);
        const uint8_t e = (uint8_t)((lane << 3) | l);
        for (uint32_t idx = rev; idx < 128; idx += (1u << l)) tab[idx] = e;
    }
    __syncwarp();
}

// Full check of a dynamic block header at absolute bit `b`: the cheap filters have passed, the code-length code is complete and
// its table is in cl_tab. Decodes the code lengths with the Kraft sums of both alphabets on the fly, so random bits fail after a
// few symbols. Lane-serial, rare.
__device__ __forceinline__ bool cand_full_check(const uint32_t *words, const uint8_t *cl_tab, uint64_t b, uint64_t end_bit, uint32_t hlit, uint32_t hdist,
                                       uint32_t hclen) {
    uint64_t pos = b + 17 + 3ull * hclen;
    const uint32_t nlit = hlit + 257, total = nlit + hdist + 1;
    uint32_t i = 0, prev = 0, lit_sum = 0, dist_sum = 0, eob_len = 0, dist_codes = 0;
    while (i < total) {
        if (pos + 14 > end_bit) return false;
        const uint32_t e = cl_tab[cand_bits(words, pos, 7)];
        pos += e & 7;
        const uint32_t s = e >> 3;
        uint32_t rep = 1, val = s;
        if (s == 16) {
            if (i == 0) return false;
            rep = 3 + cand_bits(words, pos, 2); pos += 2; val = prev;
        } else if (s == 17) { rep = 3 + cand_bits(words, pos, 3); pos += 3; val = 0; }
        else if (s == 18) { rep = 11 + cand_bits(words, pos, 7); pos += 7; val = 0; }
        if (i + rep > total) return false;
        if (val) {
            const uint32_t w = 1u << (15 - val);
            // symbols i .. i+rep-1: those below nlit belong to the literal/length alphabet
            const uint32_t nl = i >= nlit ? 0u : (i + rep <= nlit ? rep : nlit - i);
            lit_sum += nl * w;
            dist_sum += (rep - nl) * w;
            dist_codes += rep - nl;
            if (lit_sum > 32768u || dist_sum > 32768u) return false;  // over-subscribed
            if (i <= 256 && i + rep > 256) eob_len = val;
        }
        i += rep;
        prev = val;
    }
    if (!eob_len) return false;              // "missing end-of-block"
    if (lit_sum != 32768u) return false;     // incomplete literal/length code (a one-symbol code is legal but never worth a cut)
    if (!(dist_sum == 32768u || dist_sum == 0u || (dist_sum == 16384u && dist_codes == 1))) return false;
    return true;
}

// Examines the first min(32, qn) queued offsets (ascending, relative to bit0) with the second and third filter; returns the
// absolute bit of the first one that is a valid header (the rest of the queue no longer matters then), or ~0 after having
// removed them from the queue. Called by the whole warp.
__device__ __forceinline__ uint64_t cand_examine(const uint32_t *words, uint64_t bit0, uint64_t end_bit, uint32_t *q, uint32_t &qn,
                                                 const uint8_t *kraft9, uint8_t *cl_tab, uint32_t lane) {
    const uint32_t take = qn < 32 ? qn : 32;
    bool ok = lane < take;
    uint64_t b = 0;
    uint32_t hlit = 0, hdist = 0, hclen = 0;
    uint64_t clbits = 0;
    if (ok) {
        b = bit0 + q[lane];
        const uint32_t h = cand_bits(words, b, 17);
        hlit = (h >> 3) & 31; hdist = (h >> 8) & 31; hclen = ((h >> 13) & 15) + 4;
        clbits = cand_bits64(words, b + 17) & ((1ull << (3 * hclen)) - 1ull);  // 3 * hclen <= 57
        uint32_t k = 0;
#pragma unroll
        for (int f = 0; f < 7; f++) k += kraft9[(uint32_t)(clbits >> (9 * f)) & 511u];
        ok = k == 128u;  // zlib: an incomplete (or over-subscribed) code-length code is always an error
    }
    // the rare survivors, one after the other: the warp builds the table, the survivor's lane decodes the code lengths
    for (uint32_t pend = __ballot_sync(CZK_FULL, ok); pend; pend &= pend - 1) {
        const int src = __ffs((int)pend) - 1;
        cand_build_cl_tab_warp(cl_tab, __shfl_sync(CZK_FULL, clbits, src), __shfl_sync(CZK_FULL, hclen, src), lane);
        if ((int)lane == src) ok = cand_full_check(words, cl_tab, b, end_bit, hlit, hdist, hclen);
        __syncwarp();
    }
    const uint32_t m = __ballot_sync(CZK_FULL, ok);
    if (m) return bit0 + q[__ffs((int)m) - 1];  // (the queue is ascending: the lowest lane is the first offset)
    __syncwarp();
    const uint32_t rest = qn - take;  // < 32
    const uint32_t moved = lane < rest ? q[take + lane] : 0u;
    __syncwarp();
    if (lane < rest) q[lane] = moved;
    qn = rest;
    __syncwarp();
    return ~0ull;
}

// cand[c] = absolute bit of the first plausible dynamic block header in chunk c, or ~0. Three filters of rising cost per bit
// offset: the 17 header bits (BFINAL = 0, BTYPE = 10, HLIT / HDIST in range: ~11 % of random offsets pass); the Kraft sum of
// the code-length code from a table over 9 bits (three lengths) at a time — it must be complete, which ~0.5 % of those are;
// then the full decode of the code lengths.
// Only ~11 % of the offsets pass the first filter, but with 32 offsets per step some lane nearly always does, and the warp would
// pay for the second filter at every step: the survivors are queued instead and examined 32 at a time, all lanes busy. The scan
// itself runs on 32-bit offsets relative to the word that holds the chunk's first bit, one (warp-uniform) word load per step:
// lane l examines the 17 bits at bit l of the 64-bit window [w0, w1].
__global__ void __launch_bounds__(128) inflate_candidates_kernel(const uint8_t *in, const CandChunk *chunks, uint32_t n_chunks, uint64_t *cand) {
    __shared__ uint8_t kraft9[512];
    __shared__ uint8_t cl_tabs[4][128];  // per warp: the code-length decode table of the header under the full check
    __shared__ uint32_t queue[4][64];    // per warp: offsets that passed the 17-bit filter, ascending
    for (uint32_t i = threadIdx.x; i < 512; i += blockDim.x) kraft9[i] = (uint8_t)cand_kraft9(i);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t c = blockIdx.x * (blockDim.x >> 5) + warp;
    if (c >= n_chunks) return;
    const CandChunk ch = chunks[c];
    const uint32_t *words = (const uint32_t *)in;  // `in` is the (256-byte aligned) base of the device input buffer
    uint32_t *q = queue[warp];
    const uint64_t word0 = ch.lo_bit >> 5;                                   // first word of the scan
    const uint64_t bit0 = word0 << 5;
    const uint32_t first = (uint32_t)(ch.lo_bit & 31);                       // offsets below this lie before the chunk
    uint64_t last64 = ch.hi_bit;                                             // offsets >= last are out: end of the chunk, or too
    if (ch.end_bit < 17 + 57 + 14) last64 = 0;                               // close to the end of the stream for a header
    else if (ch.end_bit - (17 + 57 + 14) + 1 < last64) last64 = ch.end_bit - (17 + 57 + 14) + 1;
    const uint32_t last = last64 > bit0 ? (uint32_t)(last64 - bit0) : 0u;    // (a chunk is far below 2^32 bits)
    const uint32_t span = (uint32_t)(ch.hi_bit - bit0);
    const uint32_t *wp = words + word0;
    uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1);
    uint64_t found = ~0ull;
    uint32_t qn = 0;
    for (uint32_t rel = 0; rel < span; rel += 32) {
        const uint32_t r = rel + lane;
        const uint32_t h = __funnelshift_r(w0, w1, lane) & 0x1ffffu;
        const bool ok = r >= first && r < last && (h & 7u) == 4u && ((h >> 3) & 31) <= 29 && ((h >> 8) & 31) <= 29;  // BFINAL = 0, BTYPE = 10, HLIT / HDIST in range
        w0 = w1;
        wp++;
        w1 = __ldg(wp + 1);   // (the input buffer is padded: a word beyond the stream may be read, never used as data)
        const uint32_t m = __ballot_sync(CZK_FULL, ok);
        if (ok) q[qn + __popc(m & ((1u << lane) - 1u))] = r;
        qn += __popc(m);
        __syncwarp();
        if (qn >= 32) {
            found = cand_examine(words, bit0, ch.end_bit, q, qn, kraft9, cl_tabs[warp], lane);
            if (found != ~0ull) break;
        }
    }
    while (found == ~0ull && qn) found = cand_examine(words, bit0, ch.end_bit, q, qn, kraft9, cl_tabs[warp], lane);
    if (lane == 0) cand[c] = found;
}

// ---------------------------------------------------------------------------------------------------------------
// 2b. phase A of a run by a WHOLE WARP with one decoding lane — for launches with few runs (one long stream).
// In inflate_tok_kernel a warp carries 32 runs and issues every path any of them takes (literal, length, distance, both
// refills, the canonical code-length searches): a lane decodes at the pace of all 32 together, ~800-1 650 cycles per symbol. Here
// lane 0 alone decodes ONE run through look-up tables in shared memory (the tables and their warp-cooperative construction are
// those of inflate_kernel.cuh: 10-bit literal/length, 8-bit distance, canonical search beyond), so only the taken path is issued;
// the other lanes take part in building the tables. Tokens, TokMeta and RunResult are what inflate_tok_kernel writes in run mode.
// The clean outcomes are what matters (CZK_ST_RUN_END / ST_FINISHED): the host leaves a stream whose run reports anything else
// to the serial kernel, which reproduces zlib's partial output and status exactly.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 2) inflate_tokw_kernel(TwoPhaseParams Q) {
    const InflateParams &P = Q.base;
    __shared__ SlotSmem slots[WARPS];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SlotSmem &my = slots[warp];
    for (;;) {
        unsigned long long u64 = 0;
        if (lane == 0) u64 = atomicAdd(P.counter, 1ull);
        u64 = __shfl_sync(CZK_FULL, u64, 0);
        if (u64 >= P.n) break;
        const uint32_t unit = P.ids ? P.ids[u64] : (uint32_t)u64;
        const RunDesc rd = Q.runs[unit];
        const uint8_t *in_base = P.in + rd.in_lo;
        const uint64_t in_len = rd.in_hi - rd.in_lo;
        uint32_t *tp = Q.tok + rd.tok_off;
        uint32_t *const tp0 = tp, *const tp_end = tp + rd.tok_cap;
        const bool run_mid = rd.mid_stream != 0;
        BitReader br;
        br.words = nullptr; br.mis = 0; br.widx = br.wend = 0; br.cnt = 0; br.buf = 0; br.nextw = 0; br.nextw2 = 0; br.total = 0; br.tail_mask = 0xffffffffu;
        uint64_t pos = 0;
        uint32_t lit = 0, nlit = 0, bfinal = 0, overflow = 0;
        int result = 0;
        bool emit = true;
#define CZKW_PUT(x) do { if (emit) *tp = (x); tp++; } while (0)
#define CZKW_FLUSH_LIT() do { if (nlit) { CZKW_PUT(CZK_TOK_LIT | (nlit << 24) | lit); lit = 0; nlit = 0; } } while (0)
        int st = SS_BLOCK;
        if (lane == 0) {
            br.init(in_base, in_len);
            if (run_mid) {
                br.seek(rd.start_bit >> 3);
                br.skip((uint32_t)(rd.start_bit & 7));
            } else {
                int wrap;
                if (P.window_bits < 0) wrap = 0;
                else if (P.window_bits == 47) { br.refill(); wrap = (in_len >= 2 && br.peek(16) == 0x8b1f) ? 2 : 1; }
                else wrap = P.window_bits > 15 ? 2 : 1;
                int r = 0;
                if (wrap == 1) r = parse_zlib_header(br);
                else if (wrap == 2) r = parse_gzip_header(br);
                if (in_len == 0) { result = ST_NEED_OUTPUT; st = SS_FINISH; }
                else if (r != 0) { result = r == 100 ? ST_NEED_INPUT : r; st = SS_FINISH; }
            }
        }
        // block loop: lane 0 owns the state, the warp follows it
        for (;;) {
            uint32_t nl = 0, nd = 0;
            if (lane == 0 && st == SS_BLOCK) {
                if (br.consumed() >= rd.target_bit && br.consumed() > rd.start_bit) { result = CZK_ST_RUN_END; st = SS_FINISH; }
                else {
                    br.refill();
                    bfinal = br.get(1);
                    const uint32_t btype = br.get(2);
                    if (br.overrun()) { result = ST_NEED_INPUT; st = SS_FINISH; }
                    else if (btype == 0) {
                        br.skip((uint32_t)((0 - br.consumed()) & 7));
                        br.refill();
                        const uint32_t len = br.get(16), nlen = br.get(16);
                        if (br.overrun()) { result = ST_NEED_INPUT; st = SS_FINISH; }
                        else if ((len ^ 0xffffu) != nlen) { result = ST_E_DATA; st = SS_FINISH; }
                        else {
                            const uint64_t ipos = br.consumed() >> 3;
                            if (ipos + len > in_len) { result = ST_NEED_INPUT; st = SS_FINISH; }
                            else {
                                if (emit && tp_end - tp < 16) { emit = false; overflow = 1; }
                                if (len > 8) {
                                    CZKW_FLUSH_LIT();
                                    CZKW_PUT(CZK_TOK_STORED | len);
                                    CZKW_PUT(CZK_TOK_STORED | CZK_TOK_TAIL | (uint32_t)(ipos & 0x1fffffffu));
                                    CZKW_PUT(CZK_TOK_STORED | CZK_TOK_TAIL | (uint32_t)(ipos >> 29));
                                } else {
                                    for (uint32_t k = 0; k < len; k++) {
                                        lit |= (uint32_t)in_base[ipos + k] << (8 * nlit);
                                        if (++nlit == 3) { CZKW_PUT(CZK_TOK_LIT | (3u << 24) | lit); lit = 0; nlit = 0; }
                                    }
                                }
                                pos += len;
                                br.seek(ipos + len);
                                if (bfinal) { result = ST_FINISHED; st = SS_FINISH; }
                            }
                        }
                    } else if (btype == 1) {
                        uint8_t *lens = my.u.lens;
                        for (uint32_t i = 0; i < 288; i++) lens[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
                        for (uint32_t i = 0; i < 32; i++) lens[288 + i] = 5;
                        nl = 288; nd = 32;
                        st = SS_BUILD;
                    } else if (btype == 2) {
                        int r = parse_dynamic_header(br, my, nl, nd);
                        if (r == ST_E_DATA && br.consumed() > br.total) r = 100;
                        if (r == 0) st = SS_BUILD;
                        else { result = r == 100 ? ST_NEED_INPUT : r; st = SS_FINISH; }
                    } else { result = ST_E_DATA; st = SS_FINISH; }
                }
            }
            st = __shfl_sync(CZK_FULL, st, 0);
            if (st == SS_FINISH) break;
            if (st == SS_BLOCK) continue;  // a stored block was taken whole
            // SS_BUILD: the warp builds both tables
            nl = __shfl_sync(CZK_FULL, nl, 0);
            nd = __shfl_sync(CZK_FULL, nd, 0);
            __syncwarp();
            int rc = build_one_table(my, my.u.lens, nl, false, lane);
            if (!rc) rc = build_one_table(my, my.u.lens + nl, nd, true, lane);
            __syncwarp();
            if (rc) { if (lane == 0) { result = ST_E_DATA; } st = SS_FINISH; break; }
            if (lane == 0) {
                // ---- the symbols of this block
                st = SS_BLOCK;
                for (;;) {
                    if (emit && tp_end - tp < 8) { emit = false; overflow = 1; }
                    br.refill();
                    uint32_t e = my.lit_tab[br.peek(CZK_LIT_BITS)];
                    if ((e & 0xfff0u) == CZK_L_LONG) e = decode_long_lit(my, br.peek(15));
                    const uint32_t pay = e >> 4;
                    if (pay < 0x100) {  // literal
                        br.skip(e & 15);
                        pos++;
                        lit |= pay << (8 * nlit);
                        if (++nlit == 3) { CZKW_PUT(CZK_TOK_LIT | (3u << 24) | lit); lit = 0; nlit = 0; }
                        // (bits past the end can only have been used once the last word is in the buffer)
                        if (br.widx >= br.wend && br.overrun()) { result = ST_NEED_INPUT; st = SS_FINISH; break; }
                        continue;
                    }
                    if (!(pay & 0x800)) {
                        if (pay == 0x100) {  // end of block
                            br.skip(e & 15);
                            if (br.overrun()) { result = ST_NEED_INPUT; st = SS_FINISH; break; }
                            if (bfinal) { result = ST_FINISHED; st = SS_FINISH; }
                            break;
                        }
                        result = br.consumed() + ((e & 15) ? (e & 15) : 1) > br.total ? ST_NEED_INPUT : ST_E_DATA;  // invalid code
                        st = SS_FINISH;
                        break;
                    }
                    br.skip(e & 15);
                    const uint32_t eb = (pay >> 8) & 7;
                    const uint32_t len = 3 + (pay & 0xff) + br.peek(eb);
                    br.skip(eb);
                    br.refill();
                    uint32_t de = my.dist_tab[br.peek(CZK_DIST_BITS)];
                    if (de & CZK_D_LONG) de = decode_long_dist(my, br.peek(15));
                    if (de & CZK_D_INVALID) { result = br.consumed() + 1 > br.total ? ST_NEED_INPUT : ST_E_DATA; st = SS_FINISH; break; }
                    br.skip(de & 15);
                    const uint32_t deb = (de >> 4) & 15;
                    const uint32_t dist = (((de >> 8) & 3) << deb) + 1 + br.peek(deb);
                    br.skip(deb);
                    if (br.widx >= br.wend && br.overrun()) { result = ST_NEED_INPUT; st = SS_FINISH; break; }
                    if ((uint64_t)dist > pos && !run_mid) { result = ST_E_DATA; st = SS_FINISH; break; }  // "invalid distance too far back"
                    CZKW_FLUSH_LIT();
                    CZKW_PUT(len | (dist << 9));
                    pos += len;
                }
            }
            st = __shfl_sync(CZK_FULL, st, 0);
            if (st == SS_FINISH) break;
        }
        if (lane == 0) {
            if (emit && nlit && tp_end - tp < 2) { emit = false; overflow = 1; }
            CZKW_FLUSH_LIT();
            TokMeta m;
            m.ntok = (uint32_t)(tp - tp0);
            m.status = result;
            m.out_len = pos;
            m.expect = 0;
            m.wrap = 0;
            Q.meta[unit] = m;
            RunResult rr;
            rr.end_bit = br.consumed();
            rr.final_block = result == ST_FINISHED ? 1u : 0u;
            rr.tok_overflow = overflow;
            Q.run_res[unit] = rr;
        }
        __syncwarp();
#undef CZKW_FLUSH_LIT
#undef CZKW_PUT
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 3. phase B of a run into 16-bit symbols (the token-parallel resolution of inflate_lz_kernel, on symbols)
#define CZK_MARK 0x8000u

// symbol at run-relative index idx of S (idx < 0: a byte in front of the run -> its marker)
__device__ __forceinline__ uint32_t run_sym(const uint16_t *S, int64_t idx) {
    return idx < 0 ? (CZK_MARK | (uint32_t)(32768 + idx)) : (uint32_t)S[idx];
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (WARPS * CZK_LZ_MINB <= 32 ? CZK_LZ_MINB : 1)) inflate_lz16_kernel(TwoPhaseParams Q, uint16_t *sym) {
    const InflateParams &P = Q.base;
    const uint32_t lane = threadIdx.x & 31;
    constexpr int SHORT = CZK_LZ_SHORT;
    for (;;) {
        unsigned long long u64 = 0;
        if (lane == 0) u64 = atomicAdd(Q.counter_b, 1ull);
        u64 = __shfl_sync(CZK_FULL, u64, 0);
        if (u64 >= P.n) break;
        const uint32_t unit = P.ids ? P.ids[u64] : (uint32_t)u64;
        const uint64_t o0 = P.out_off[unit];
        const TokMeta m = Q.meta[unit];
        uint16_t *S = sym + (o0 - P.out_off[0]);             // symbols of this run
        const uint8_t *ib = P.in + Q.runs[unit].in_lo;        // stored runs reference the stream's input
        const uint32_t *tok = Q.runs[unit].tok_cap ? Q.tok + Q.runs[unit].tok_off : Q.tok + tok_word_off(o0 - P.out_off[0], unit);
        const uint32_t ntok = m.ntok;
        uint64_t opos = 0;
        uint32_t ti = 0;
        while (ti < ntok) {
            const uint32_t t_raw = ti + lane < ntok ? tok[ti + lane] : CZK_TOK_STORED;
            const uint32_t stopm = __ballot_sync(CZK_FULL, (t_raw >> 30) == 3u);
            const uint32_t nt = stopm ? (uint32_t)__ffs(stopm) - 1u : 32u;
            if (nt) {
                const uint32_t t = lane < nt ? t_raw : 0u;
                const uint32_t tl = (t >> 31) ? ((t >> 24) & 3u) : (t & 0x1ffu);
                uint32_t pos = tl;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    uint32_t v = __shfl_up_sync(CZK_FULL, pos, d);
                    if ((int)lane >= d) pos += v;
                }
                const uint32_t total = __shfl_sync(CZK_FULL, pos, 31);
                pos -= tl;
                const int64_t base = (int64_t)opos;  // run-relative index of this group's first symbol
                const bool is_tok = lane < nt;
                const bool is_lit = is_tok && (t >> 31);
                const uint32_t dist = (t >> 9) & 0xffffu;
                const int ipos = (int)pos;
                const int src0 = ipos - (int)dist;
                int dep_end = 0;
                if (is_tok && !is_lit) { dep_end = src0 + (int)tl; if (dep_end > ipos) dep_end = ipos; }
                const bool coop = is_tok && !is_lit && (tl > (uint32_t)SHORT || dist < tl);
                bool done = !is_tok;
                for (;;) {
                    const uint32_t undone = __ballot_sync(CZK_FULL, !done);
                    if (!undone) break;
                    const int first = __ffs((int)undone) - 1;
                    const int frontier = __shfl_sync(CZK_FULL, ipos, first);
                    const bool ready = !done && dep_end <= frontier;
                    if (ready && !coop) {
                        uint16_t *d = S + base + ipos;
                        if (is_lit) {
                            d[0] = (uint16_t)(t & 0xff);
                            if (tl > 1) d[1] = (uint16_t)((t >> 8) & 0xff);
                            if (tl > 2) d[2] = (uint16_t)((t >> 16) & 0xff);
                        } else {
                            uint32_t bb[SHORT];
#pragma unroll
                            for (int k = 0; k < SHORT; k++) bb[k] = k < (int)tl ? run_sym(S, base + src0 + k) : 0u;
#pragma unroll
                            for (int k = 0; k < SHORT; k++) if (k < (int)tl) d[k] = (uint16_t)bb[k];
                        }
                    }
                    uint32_t cm = __ballot_sync(CZK_FULL, ready && coop);
                    while (cm) {
                        const int sl = __ffs((int)cm) - 1;
                        cm &= cm - 1;
                        const uint32_t Ln = __shfl_sync(CZK_FULL, tl, sl), D = __shfl_sync(CZK_FULL, dist, sl);
                        const int P0 = __shfl_sync(CZK_FULL, ipos, sl);
                        uint16_t *d = S + base + P0;
                        const int64_t s0 = base + P0 - (int64_t)D;  // symbols [s0, s0 + D) are complete
                        if (D >= Ln) {
                            for (uint32_t k = lane; k < Ln; k += 32) d[k] = (uint16_t)run_sym(S, s0 + k);
                        } else if (D >= 32) {
                            for (uint32_t k0 = 0; k0 < Ln; k0 += 32) {
                                const uint32_t k = k0 + lane;
                                if (k < Ln) d[k] = (uint16_t)run_sym(S, s0 + k);
                                __syncwarp();
                            }
                        } else {
                            uint32_t r = lane % D;
                            const uint32_t stepD = 32u % D;
                            for (uint32_t k = lane; k < Ln; k += 32) {
                                d[k] = (uint16_t)run_sym(S, s0 + r);
                                r += stepD;
                                if (r >= D) r -= D;
                            }
                        }
                    }
                    done = done || ready;
                    __syncwarp();
                }
                opos += total;
                ti += nt;
            }
            if (nt < 32 && ti < ntok) {
                const uint32_t w0 = tok[ti], w1 = tok[ti + 1], w2 = tok[ti + 2];
                const uint32_t n = w0 & 0xffffu;
                const uint64_t ipos_in = (uint64_t)(w1 & 0x1fffffffu) | ((uint64_t)(w2 & 0x1fffffffu) << 29);
                for (uint32_t k = lane; k < n; k += 32) S[opos + k] = (uint16_t)ib[ipos_in + k];
                __syncwarp();
                opos += n;
                ti += 3;
            }
        }
        if (lane == 0) {
            P.out_lens[unit] = opos;
            P.statuses[unit] = m.status;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 4. windows: per stream, serially over its runs. win[r] (32 KiB) = the 32 KiB of final bytes that precede run r + 1, i.e.
// the last 32 KiB up to the end of run r (zero-filled where the stream is shorter; such positions are never referenced by a
// valid stream — a marker that points there is reported through `bad`).
// The chain is cut wherever it can be: a run of at least 32 KiB whose last 32 KiB hold no marker has a window that does not
// depend on anything before it (streams with full-flush points — ours, pigz's — break every segment), so the kernel works on
// CHAINS: runs [first_run, first_run + n_runs) of one stream, the first of which starts the stream or is such a breaker.
struct RunStream {
    uint32_t first_run, n_runs;
    uint64_t stream_start;       // offset (symbol space) of the first byte of the stream the chain belongs to
};

// flags[r] = 0 when run r is a chain breaker (>= 32 KiB long, no marker in its last 32 KiB), else 1. One warp per run.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) inflate_tail_markers_kernel(const uint64_t *run_off, uint32_t n_runs, const uint16_t *sym, uint8_t *flags) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (r >= n_runs) return;
    const uint64_t a = run_off[r], e = run_off[r + 1];
    uint32_t any = e - a < 32768 ? 1u : 0u;
    if (!any) {
        // 16-byte loads where the tail is aligned (symbol offsets are even byte addresses; the head up to alignment goes singly)
        uint64_t p = e - 32768;
        while ((p & 7) && p < e) { if (lane == 0) any |= sym[p] >> 15; p++; }
        const uint4 *v = (const uint4 *)(sym + p);
        const uint64_t nv = (e - p) >> 3;
        for (uint64_t k = lane; k < nv; k += 32) {
            const uint4 x = v[k];
            any |= ((x.x | x.y | x.z | x.w) & 0x80008000u) ? 1u : 0u;
        }
        for (uint64_t q = p + nv * 8 + lane; q < e; q += 32) any |= sym[q] >> 15;
    }
    any = __any_sync(CZK_FULL, any != 0);
    if (lane == 0) flags[r] = (uint8_t)any;
}

// run_off[n_runs + 1]: the runs' offsets in the launch's compact symbol space (run r = sym[run_off[r], run_off[r + 1])).
#define CZK_WINDOW_SMEM 65536
__global__ void __launch_bounds__(1024) inflate_window_kernel(const RunStream *streams, const uint64_t *run_off, const uint16_t *sym,
                                                              uint8_t *win, uint32_t *bad) {
    CZ_DYNAMIC_SMEM(smem_raw);
    uint8_t (*W)[32768] = (uint8_t (*)[32768])smem_raw;
    const RunStream rs = streams[blockIdx.x];
    const uint32_t tid = threadIdx.x;
    const uint64_t stream_start = rs.stream_start;
    int cur = 0;
    uint32_t my_bad = 0;
    for (uint32_t k = 0; k < rs.n_runs; k++) {
        const uint32_t r = rs.first_run + k;
        const uint64_t a = run_off[r], e = run_off[r + 1], len = e - a;
        const uint8_t *Wp = W[cur];
        uint8_t *Wn = W[cur ^ 1];
        const uint64_t valid_before = a - stream_start;  // bytes of the stream in front of this run
        for (uint32_t j = tid; j < 32768; j += 1024) {
            // window position j <-> stream position e - 32768 + j
            uint32_t v = 0;
            if (len + j >= 32768) {  // inside this run
                const uint64_t idx = len + j - 32768;
                const uint32_t s = sym[a + idx];
                if (s & CZK_MARK) {
                    const uint32_t w = s & 0x7fffu;
                    if (32768u - w > valid_before) my_bad = 1;  // reaches before the start of the stream
                    else v = Wp[w];
                } else v = s;
            } else {  // still in front of this run: carried over from the previous window
                const uint32_t w = (uint32_t)(j + len);
                v = (k == 0) ? 0u : Wp[w];
            }
            Wn[j] = (uint8_t)v;
        }
        __syncthreads();
        uint4 *dst = (uint4 *)(win + (uint64_t)r * 32768);
        const uint4 *srcv = (const uint4 *)Wn;
        for (uint32_t j = tid; j < 2048; j += 1024) dst[j] = srcv[j];
        cur ^= 1;
        __syncthreads();
    }
    if (my_bad) atomicOr(bad, 1u);
}

// ---------------------------------------------------------------------------------------------------------------
// 5. symbols -> bytes. One CTA per (run, 64 KiB slice): the window in front of the run in shared memory. Run r's symbols are
// sym[run_off[r], run_off[r + 1]) and its bytes go to out[final_off[r] ...). Checks per run are computed afterwards from the bytes.
struct RunSlice {
    uint32_t run;      // run index in the launch
    uint32_t slice;    // 64 KiB slice of the run
};
__global__ void __launch_bounds__(256) inflate_resolve_kernel(const RunSlice *slices, uint32_t n_slices, const uint64_t *run_off,
                                                              const uint64_t *final_off, const uint8_t *run_is_first, const uint16_t *sym,
                                                              const uint8_t *win, uint8_t *out, uint32_t *bad) {
    __shared__ __align__(16) uint8_t W[32768];
    const uint32_t sl = blockIdx.x;
    if (sl >= n_slices) return;
    const uint32_t r = slices[sl].run;
    const uint64_t a = run_off[r], e = run_off[r + 1];
    const uint64_t lo = a + (uint64_t)slices[sl].slice * 65536;
    const uint64_t hi = lo + 65536 < e ? lo + 65536 : e;
    const bool first = run_is_first[r] != 0;
    if (!first) {
        const uint4 *srcv = (const uint4 *)(win + (uint64_t)(r - 1) * 32768);
        uint4 *dstv = (uint4 *)W;
        for (uint32_t j = threadIdx.x; j < 2048; j += 256) dstv[j] = srcv[j];
    }
    __syncthreads();
    uint8_t *ob = out + final_off[r];
    uint32_t my_bad = 0;
    for (uint64_t p = lo + threadIdx.x; p < hi; p += 256) {
        const uint32_t s = sym[p];
        uint32_t v = s;
        if (s & CZK_MARK) {
            if (first) { my_bad = 1; v = 0; }  // nothing precedes the first run of a stream
            else v = W[s & 0x7fffu];
        }
        ob[p - a] = (uint8_t)v;
    }
    if (my_bad) atomicOr(bad, 1u);
}

// Adler-32 and CRC-32 of every run's final bytes: one warp per run. checks[2r] = adler (seed 1), checks[2r+1] = crc (seed 0).
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) inflate_run_check_kernel(const uint64_t *run_off, const uint64_t *final_off, uint32_t n_runs,
                                                                       const uint8_t *out, const CrcTables *crc, int check_kind, uint32_t *checks) {
    __shared__ uint32_t crc_tab[256 + 34];
    for (uint32_t i = threadIdx.x; i < 256 + 34; i += WARPS * 32) crc_tab[i] = i < 256 ? crc->table[i] : crc->pow128[i - 256];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (r >= n_runs) return;
    const uint64_t a = final_off[r], e = a + (run_off[r + 1] - run_off[r]);
    uint32_t adler = 1, c = 0;
    if (check_kind & 1)
        for (uint64_t p = a; p < e; p += 8192) adler = warp_adler32(adler, out + p, (uint32_t)(e - p < 8192 ? e - p : 8192), lane);
    if (check_kind & 2) {
        uint64_t p = a;
        while (e - p >= 128) {
            uint32_t q = (uint32_t)((e - p) >> 7);
            if (q > 32) q = 32;
            c = warp_crc32_pieces(c, out + p, q, crc_tab, crc_tab + 256, lane);
            p += (uint64_t)q * 128;
        }
        if (p < e) {
            uint32_t c2 = 0;
            if (lane == 0) c2 = crc32_serial(c, out + p, (uint32_t)(e - p), crc_tab);
            c = __shfl_sync(CZK_FULL, c2, 0);
        }
    }
    if (lane == 0) { checks[2 * r] = adler; checks[2 * r + 1] = c; }
}

}  // namespace czk
