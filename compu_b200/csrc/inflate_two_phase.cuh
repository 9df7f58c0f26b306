// inflate_two_phase.cuh — batched inflate as TWO kernels (the default path).
//
// Why (profiles/r1_inflate_lc7_ncu.md): with one lane per stream doing everything, 65 536 streams are in flight at once
// and their 32 KiB back-reference windows (2 GiB) thrash the 126 MB L2 — 13x the algorithmic DRAM traffic, every byte
// store its own sector, the warps stalled on L2 misses. Huffman decoding is the only part that is inherently serial per
// stream; LZ77 resolution is not. So the work is split where the dependency structure changes:
//
//   phase A  inflate_tok_kernel   one LANE per stream: container header, block headers, canonical Huffman decode (the 2x15
//                                 code limits in registers, 448 B shared memory per stream, as inflate_lc_kernel.cuh).
//                                 It produces no output bytes: it writes a compact TOKEN stream (sequential 4-byte
//                                 stores) and validates everything that does not need the bytes (distances, capacity,
//                                 block structure, ISIZE). Only sequential traffic: compressed in, tokens out.
//   phase B  inflate_lz_kernel    one WARP per stream, few streams in flight (their windows fit L2): 32 tokens per step,
//                                 a warp scan turns lengths into output positions, bytes are produced 32 at a time,
//                                 sector-aligned, so every global store is one full 32-byte sector; back-references
//                                 inside the round are resolved with shuffles. Adler-32 / CRC-32 are computed by the
//                                 warp from the bytes it has just written and compared with the trailer values that
//                                 phase A parsed.
//
// Token words (uint32):
//   match    bit31 = 0           bits 0-8 length (1..258, may be cut by the output capacity), bits 9-24 distance (1..32768)
//   literals bits 31-30 = 10     bits 24-25 count (1..3), bits 0-23 the bytes, first byte lowest
//   stored   bits 31-30 = 11     three words: word 0 (bit 29 = 0) holds the length in its low 16 bits; words 1 and 2 (bit 29 = 1,
//                                "tail" words) hold the byte offset of the run inside the unit's input, 29 bits each (low, high)
// Worst case one word per output byte (1-byte stored blocks are written as literal tokens), hence the 4 x capacity
// token area per unit.
//
// Contract, status numbering and zlib's error order are those of inflate_kernel.cuh / inflate_lc_kernel.cuh.
#pragma once
#include "inflate_lc_kernel.cuh"

namespace czk {

struct TokMeta {
    uint32_t ntok;       // token words written
    int32_t status;      // preliminary status (final unless a checksum comparison is pending)
    uint64_t out_len;    // bytes the tokens produce
    uint32_t expect;     // trailer check value (Adler-32 or CRC-32) when status == FINISHED and wrap != 0
    uint32_t wrap;       // 0 raw, 1 zlib (Adler-32), 2 gzip (CRC-32)
};

// RUN mode (block-parallel decode of long streams, inflate_runs.cuh): a unit of phase A is a RUN — a range of ONE stream's
// blocks that starts at a block boundary found by the candidate search (or at the start of the stream) and stops at the first
// block boundary at or after its target (where the next run starts). All runs of a stream share the stream's input range, so
// bit positions are relative to the start of the stream.
struct RunDesc {
    uint64_t in_lo, in_hi; // the STREAM's input: bytes [in_lo, in_hi) of the launch's input buffer (shared by all its runs)
    uint64_t start_bit;   // first bit of the run's first block header (run 0 of a stream: 0, the container header comes first)
    uint64_t target_bit;  // stop at the first block boundary >= this (UINT64_MAX: run to the end of the stream)
    uint32_t mid_stream;  // 1: starts at a block header inside the stream (no container header, distances may reach before the run)
    uint32_t pad;
    uint64_t tok_off;     // tok_cap != 0: the run's token area is Q.tok[tok_off, tok_off + tok_cap) — sizes are not known yet when a
    uint64_t tok_cap;     // run is decoded once only, so the area is bounded by the run's COMPRESSED size; a run that needs more
                          // (it ran past a false next candidate) stops emitting, keeps counting and reports tok_overflow
};
struct RunResult {
    uint64_t end_bit;     // where the run stopped (a block boundary, or the end of the final block)
    uint32_t final_block; // 1: the run ended with the stream's final block
    uint32_t tok_overflow; // 1: the token area was too small: sizes and positions are valid, the tokens are not
};
#define CZK_ST_RUN_END 4  // phase A, run mode: stopped at the target boundary (internal status, never reaches the caller)

struct TwoPhaseParams {
    InflateParams base;          // base.counter: phase A work counter
    uint32_t *tok;               // token area: unit u starts at word tok_word_off(out_off[u] - out_off[0], u)
    TokMeta *meta;               // n
    unsigned long long *counter_b;  // phase B work counter, zero before launch
    int32_t count_only;          // phase A only: no tokens are written, TokMeta.out_len / status / in_consumed are the result
    unsigned long long *counter_c;  // work counter of inflate_lz_cta_kernel, zero before launch
    uint32_t cta_tile;           // != 0: units whose output slot is at most this many bytes belong to inflate_lz_cta_kernel
    uint32_t spin_ns;            // inflate_lz_cta_kernel: back-off of a warp that found no ready token
    const RunDesc *runs;         // != null: run mode, one RunDesc per unit
    RunResult *run_res;          // run mode: per unit
    uint32_t lane_step;          // phase A: > 1 = only lanes with lane % lane_step == 0 take units (launches with few units: a warp
                                 // pays for every path any of its lanes takes, so a lane that is alone in its warp decodes its
                                 // stream 2-3 x sooner; the host picks the largest step that still keeps all units in one wave)
};

__host__ __device__ inline uint64_t tok_word_off(uint64_t out_off, uint64_t unit) { return out_off + 8 * unit; }

#define CZK_TOK_LIT 0x80000000u
#define CZK_TOK_STORED 0xC0000000u
#define CZK_TOK_TAIL 0x20000000u
#ifndef CZK_LZ_MINB
#define CZK_LZ_MINB 4  // 64 registers: 32 warps per SM (5: 48 registers, 40 warps: 33.3 ms; 6: spills: 40 ms; measured on cfg2)
#endif
// Phase A reads every stream word by word, one lane per stream: L1 (22 KB beside the slots) cannot keep 448 lines, so each
// load is an L2 access, and a DRAM access whenever its sector is new. The fetch-size hint makes one DRAM access bring the next
// CZK_TOK_L2_FETCH bytes of the lane's stream into L2 (0: plain __ldg).
#ifndef CZK_TOK_L2_FETCH
#define CZK_TOK_L2_FETCH 256
#endif
__device__ __forceinline__ uint32_t czk_ldg_stream(const uint32_t *p) {
#if defined(__CUDA_ARCH__) && CZK_TOK_L2_FETCH == 256
    uint32_t v;
    asm volatile("ld.global.nc.L2::256B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#elif defined(__CUDA_ARCH__) && CZK_TOK_L2_FETCH == 128
    uint32_t v;
    asm volatile("ld.global.nc.L2::128B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

#ifndef CZK_LZ_ABL
#define CZK_LZ_ABL 0
#endif
#ifndef CZK_LZ_WIDE
#define CZK_LZ_WIDE 1  // phase B: short matches read their source as aligned 8-byte words (0: one load per byte)
#endif
#ifndef CZK_LZ_SHORT
#define CZK_LZ_SHORT 12
#endif
// phase B (token-parallel): matches up to this long are copied by their own lane (8/12/16/24/32 measured: 33.9/32.6/34.0/35.5/38.0 ms)

// EMIT = false: counting only (no token is written; sizes, statuses and end positions are the result).
template <int WARPS, bool EMIT = true>
__global__ void __launch_bounds__(WARPS * 32) inflate_tok_kernel(TwoPhaseParams Q) {
    const InflateParams &P = Q.base;
    CZ_DYNAMIC_SMEM(smem_raw);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // layout: [length info 128 B] [per-warp scratch 256 B each] [WARPS*32 slots]
    uint16_t *len_info = (uint16_t *)smem_raw;
    uint32_t *wscr = (uint32_t *)(smem_raw + 128) + warp * 64;
    LcSlot *slots = (LcSlot *)(smem_raw + 128 + WARPS * 256) + (size_t)warp * 32;
    uint32_t my_off = lane * (uint32_t)sizeof(LcSlot);
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+r"(my_off));  // keep the slot address in a register (otherwise it is recomputed from %tid per symbol)
#endif
    LcSlot &my = *(LcSlot *)((uint8_t *)slots + my_off);
    if (threadIdx.x < 32) len_info[threadIdx.x] = (uint16_t)(threadIdx.x < 29 ? lc_len_info(threadIdx.x) : 0);
    __syncthreads();

    BitReader br;
    br.words = nullptr; br.mis = 0; br.widx = br.wend = 0; br.cnt = 0; br.buf = 0; br.nextw = 0; br.nextw2 = 0; br.total = 0; br.tail_mask = 0xffffffffu;
    uint32_t llim[8], dlim[8];
    lc_limits_reset(llim); lc_limits_reset(dlim);
    int st = SS_IDLE;
    uint32_t unit = 0;
    const uint8_t *in_base = nullptr;
    uint64_t in_len = 0;
    uint64_t pos = 0, cap = 0;
    uint32_t *tp = nullptr, *tp0 = nullptr;   // token write pointer / start
    uint32_t lit = 0, nlit = 0;               // pending literal bytes (at most 2 between tokens)
    int result = 0, wrap = 0;
    uint32_t bfinal = 0, nlit_sym = 0, ndist_sym = 0, stored_len = 0, expect = 0;
    uint64_t run_start = 0, run_target = ~0ull;  // run mode
    bool run_mid = false;

    bool emit = EMIT;                 // (run mode switches it off when a run outgrows its token area)
    uint32_t *tp_end = nullptr;       // run mode with a bounded token area: its end
    uint32_t tok_overflow = 0;
#define CZK_PUT(x) do { if (emit) *tp = (x); tp++; } while (0)  // (words are counted even when nothing is stored)
#define CZK_FLUSH_LIT() do { if (nlit) { CZK_PUT(CZK_TOK_LIT | (nlit << 24) | lit); lit = 0; nlit = 0; } } while (0)

    for (;;) {
        // ---- (1) fetch work
        if (st == SS_IDLE && Q.lane_step > 1 && (lane % Q.lane_step) != 0) st = SS_EXIT;
        if (st == SS_IDLE) {
            unsigned long long u = atomicAdd(P.counter, 1ull);
            if (u >= P.n) st = SS_EXIT;
            else {
                unit = P.ids ? P.ids[u] : (uint32_t)u;
                uint64_t i0, i1;
                if (Q.runs) { i0 = Q.runs[unit].in_lo; i1 = Q.runs[unit].in_hi; }
                else { i0 = P.in_off[unit]; i1 = P.in_off[unit + 1]; }
                const uint64_t o0 = P.out_off[unit], o1 = P.out_off[unit + 1];
                in_base = P.in + i0; in_len = i1 - i0;
                cap = o1 - o0; pos = 0;
                emit = EMIT; tok_overflow = 0; tp_end = nullptr;
                tp0 = tp = emit ? Q.tok + tok_word_off(o0 - P.out_off[0], unit) : (uint32_t *)nullptr + 1024;
                if (EMIT && Q.runs && Q.runs[unit].tok_cap) {
                    tp0 = tp = Q.tok + Q.runs[unit].tok_off;
                    tp_end = tp + Q.runs[unit].tok_cap;
                }
                lit = 0; nlit = 0; bfinal = 0; result = 0; expect = 0;
                br.init(in_base, in_len);
                st = SS_HEADER;
                if (Q.runs) {
                    const RunDesc rd = Q.runs[unit];
                    run_start = rd.start_bit; run_target = rd.target_bit; run_mid = rd.mid_stream != 0;
                    if (run_mid) {
                        br.seek(run_start >> 3);
                        br.skip((uint32_t)(run_start & 7));
                        wrap = 0;
                        st = SS_BLOCK;
                    }
                }
            }
        }
        if (__all_sync(CZK_FULL, st == SS_EXIT)) break;

        // ---- (2) container header
        if (st == SS_HEADER) {
            int r = 0;
            if (P.segment_mode || P.window_bits < 0) wrap = 0;
            else if (P.window_bits == 47) {
                br.refill();
                wrap = (in_len >= 2 && br.peek(16) == 0x8b1f) ? 2 : 1;
            } else wrap = P.window_bits > 15 ? 2 : 1;
            if (wrap == 1) r = parse_zlib_header(br);
            else if (wrap == 2) r = parse_gzip_header(br);
            // an empty unit: one inflate() call with avail_in == 0 makes no progress — Z_BUF_ERROR, which compu's glue reports
            // as NeedOutput (/root/reference/src/decoder/mod.rs:481)
            if (in_len == 0 && !P.segment_mode) { result = ST_NEED_OUTPUT; st = SS_FINISH; }
            else if (r == 0) st = SS_BLOCK;
            else { result = r == 100 ? ST_NEED_INPUT : r; st = SS_FINISH; }
        }

        // ---- (3) block header
        if (st == SS_BLOCK) {
            if (P.segment_mode && br.consumed() >= br.total) {
                result = br.consumed() == br.total ? ST_FINISHED : ST_NEED_INPUT;
                st = SS_TRAILER;
            } else if (Q.runs && br.consumed() >= run_target && br.consumed() > run_start) {
                result = CZK_ST_RUN_END;  // the next run starts at (or before) this boundary
                st = SS_FINISH;
            } else {
                br.refill();
                bfinal = br.get(1);
                uint32_t btype = br.get(2);
                if (br.overrun()) { result = ST_NEED_INPUT; st = SS_FINISH; }
                else if (btype == 0) {
                    br.skip((uint32_t)((0 - br.consumed()) & 7));
                    br.refill();
                    uint32_t len = br.get(16);
                    uint32_t nlen = br.get(16);
                    if (br.overrun()) { result = ST_NEED_INPUT; st = SS_FINISH; }
                    else if ((len ^ 0xffffu) != nlen) { result = ST_E_DATA; st = SS_FINISH; }  // "invalid stored block lengths"
                    else { stored_len = len; st = SS_STORED; }
                } else if (btype == 1) {
                    uint8_t *lens = lc_lens(my);
                    for (uint32_t i = 0; i < 288; i++) lens[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
                    for (uint32_t i = 0; i < 32; i++) lens[288 + i] = 5;
                    nlit_sym = 288; ndist_sym = 32;
                    st = SS_BUILD;
                } else if (btype == 2) {
                    int r = lc_parse_dynamic(br, my, nlit_sym, ndist_sym);
                    // a verdict reached with bits past the end of the input is not a verdict: zlib would still be waiting
                    if (r == ST_E_DATA && br.consumed() > br.total) r = 100;
                    if (r == 0) st = SS_BUILD;
                    else { result = r == 100 ? ST_NEED_INPUT : r; st = SS_FINISH; }
                } else { result = ST_E_DATA; st = SS_FINISH; }  // "invalid block type"
            }
        }

        // ---- (4) warp-cooperative table construction, one slot at a time
        {
            uint32_t m = __ballot_sync(CZK_FULL, st == SS_BUILD);
            while (m) {
                int s = __ffs(m) - 1;
                m &= m - 1;
                uint32_t nl = __shfl_sync(CZK_FULL, nlit_sym, s), nd = __shfl_sync(CZK_FULL, ndist_sym, s);
                __syncwarp();
                int r = lc_build(slots[s], nl, nd, wscr, wscr + 16, lane, (uint32_t)s, llim, dlim);
                if ((int)lane == s) {
                    if (r) { result = ST_E_DATA; st = SS_FINISH; }
                    else st = SS_DECODE;
                }
            }
        }

        // ---- (5) decode into tokens, lane-local, CZK_LC_BUDGET symbols per visit
        // One visit of the loop consumes at most 48 bits and produces at most 258 bytes. While three more input words and
        // 258 bytes of capacity remain, none of the end-of-input / end-of-slot tests can fire: they are evaluated only when
        // `slow` is set (the last ~12 bytes of input or the last 258 bytes of the slot).
        if (st == SS_DECODE) {
            int budget = CZK_LC_BUDGET;
            const int64_t pos_safe = (int64_t)cap - 258;
            if (EMIT && tp_end && emit && tp_end - tp < 2 * CZK_LC_BUDGET + 8) { emit = false; tok_overflow = 1; }
            // ---- fast loop. While three more input words and 258 bytes of capacity remain, none of the end-of-input /
            // end-of-slot conditions can occur (one symbol takes at most 48 bits and 258 bytes): no masking of the unit's
            // last word, no overrun tests, 32-bit position arithmetic, and the only exits are the ones below. Everything
            // unusual — the last words of the input, the end of the slot, end of block, an invalid code — is left untouched
            // for the general loop that follows (the symbol is not consumed here).
            if ((int64_t)pos <= pos_safe) {
                uint64_t buf = br.buf;
                uint32_t cnt = br.cnt, widx = br.widx, nextw = br.nextw, nextw2 = br.nextw2;
                const uint32_t wend = br.wend;
                const uint32_t *const words = br.words;
                const uint64_t room64 = (uint64_t)(pos_safe - (int64_t)pos);
                const uint32_t room = room64 > 0x7fff0000u ? 0x7fff0000u : (uint32_t)room64;  // symbols may START while adv <= room
                const uint32_t reach = pos >= 32768u || run_mid ? 32768u : (uint32_t)pos;     // distances up to reach + adv are valid
                uint32_t adv = 0;
                while (budget > 0 && widx + 3 <= wend && adv <= room) {
                    if (cnt <= 32) {
                        buf |= (uint64_t)nextw << cnt;
                        cnt += 32;
                        widx++;
                        nextw = nextw2;
                        nextw2 = widx + 1 < wend ? czk_ldg_stream(words + widx + 1) : 0u;
                    }
                    uint32_t v = __brev((uint32_t)buf) >> 16;
                    const uint32_t cl = lc_code_len(v, llim);
                    if (cl > 15) break;
                    const uint32_t info = my.lit_info[cl];
                    const uint32_t idx = ((v >> (16 - cl)) + info) & 0xffffu;
                    const uint32_t sym = my.lit_sorted[idx < 288 ? idx : 287] | (idx >= (info >> 16) ? 256u : 0u);
                    if (sym < 256) {  // literal
                        buf >>= cl; cnt -= cl;
                        adv++;
                        lit |= sym << (8 * nlit);
                        if (++nlit == 3) { CZK_PUT(CZK_TOK_LIT | (3u << 24) | lit); lit = 0; nlit = 0; }
                        budget--;
                        continue;
                    }
                    if (sym == 256 || sym > 285) break;  // end of block / invalid: the general loop decodes it again
                    // length + distance; the bit reader is only committed once the whole pair is known to be valid
                    uint64_t b2 = buf >> cl;
                    uint32_t c2 = cnt - cl, w2 = widx, n1 = nextw, n2 = nextw2;
                    const uint32_t li = len_info[sym - 257];
                    const uint32_t eb = li >> 9;
                    const uint32_t len = (li & 0x1ff) + ((uint32_t)b2 & ((1u << eb) - 1u));
                    b2 >>= eb; c2 -= eb;
                    if (c2 <= 32) {
                        b2 |= (uint64_t)n1 << c2;
                        c2 += 32;
                        w2++;
                        n1 = n2;
                        n2 = w2 + 1 < wend ? czk_ldg_stream(words + w2 + 1) : 0u;
                    }
                    v = __brev((uint32_t)b2) >> 16;
                    const uint32_t dcl = lc_code_len(v, dlim);
                    if (dcl > 15) break;
                    const uint32_t dsym = my.dist_sorted[((v >> (16 - dcl)) + my.dist_base[dcl]) & 31];
                    if (dsym > 29) break;
                    b2 >>= dcl; c2 -= dcl;
                    const uint32_t deb = dsym < 2 ? 0 : (dsym >> 1) - 1;
                    const uint32_t dist = ((dsym < 2 ? dsym : 2 + (dsym & 1)) << deb) + 1 + ((uint32_t)b2 & ((1u << deb) - 1u));
                    b2 >>= deb; c2 -= deb;
                    if (dist > reach + adv) break;  // "invalid distance too far back": reported by the general loop
                    buf = b2; cnt = c2; widx = w2; nextw = n1; nextw2 = n2;
                    CZK_FLUSH_LIT();
                    CZK_PUT(len | (dist << 9));
                    adv += len;
                    budget--;
                }
                br.buf = buf; br.cnt = cnt; br.widx = widx; br.nextw = nextw; br.nextw2 = nextw2;
                pos += adv;
            }
            while (budget-- > 0) {
                br.refill();
                const bool slow = !(br.widx + 3 <= br.wend && (int64_t)pos <= pos_safe);
                uint32_t v = __brev((uint32_t)br.buf) >> 16;
                uint32_t cl = lc_code_len(v, llim);
                if (cl > 15) {  // no code matches (incomplete set) — or zero bits past a truncated input
                    result = br.consumed() + 1 > br.total ? ST_NEED_INPUT : ST_E_DATA;
                    st = SS_FINISH;
                    break;
                }
                uint32_t info = my.lit_info[cl];
                uint32_t idx = ((v >> (16 - cl)) + info) & 0xffffu;
                uint32_t sym = my.lit_sorted[idx < 288 ? idx : 287] | (idx >= (info >> 16) ? 256u : 0u);
                br.skip(cl);
                if (sym < 256) {  // literal
                    if (slow) {
                        if (br.overrun()) { result = ST_NEED_INPUT; st = SS_FINISH; break; }
                        if (pos >= cap) { result = br.out_full_status(); st = SS_FINISH; break; }
                    }
                    pos++;
                    lit |= sym << (8 * nlit);
                    if (++nlit == 3) { CZK_PUT(CZK_TOK_LIT | (3u << 24) | lit); lit = 0; nlit = 0; }
                    continue;
                }
                if (sym == 256) {  // end of block
                    if (slow && br.overrun()) { result = ST_NEED_INPUT; st = SS_FINISH; break; }
                    if (bfinal) { result = ST_FINISHED; st = SS_TRAILER; } else st = SS_BLOCK;
                    break;
                }
                if (sym > 285) {  // 286/287 only exist in the fixed code and are invalid
                    result = br.overrun() ? ST_NEED_INPUT : ST_E_DATA;
                    st = SS_FINISH;
                    break;
                }
                uint32_t li = len_info[sym - 257];
                uint32_t eb = li >> 9;
                uint32_t len = (li & 0x1ff) + br.peek(eb);
                br.skip(eb);
                br.refill();
                v = __brev((uint32_t)br.buf) >> 16;
                uint32_t dcl = lc_code_len(v, dlim);
                if (dcl > 15) {
                    result = br.consumed() + 1 > br.total ? ST_NEED_INPUT : ST_E_DATA;
                    st = SS_FINISH;
                    break;
                }
                uint32_t dsym = my.dist_sorted[((v >> (16 - dcl)) + my.dist_base[dcl]) & 31];
                br.skip(dcl);
                if (dsym > 29) { result = br.overrun() ? ST_NEED_INPUT : ST_E_DATA; st = SS_FINISH; break; }
                uint32_t deb = dsym < 2 ? 0 : (dsym >> 1) - 1;
                uint32_t dist = ((dsym < 2 ? dsym : 2 + (dsym & 1)) << deb) + 1 + br.peek(deb);
                br.skip(deb);
                uint32_t n = len;
                if (slow) {
                    if (br.overrun()) { result = ST_NEED_INPUT; st = SS_FINISH; break; }
                    // zlib order (inflate.c MATCH): output space first, then "invalid distance too far back"
                    if (pos >= cap) { result = br.out_full_status(); st = SS_FINISH; break; }
                    if (pos + n > cap) n = (uint32_t)(cap - pos);
                }
                if ((uint64_t)dist > pos && !run_mid) { result = ST_E_DATA; st = SS_FINISH; break; }
                CZK_FLUSH_LIT();
                CZK_PUT(n | (dist << 9));
                pos += n;
                if (n < len) { result = br.out_full_status(); st = SS_FINISH; break; }
            }
        }

        // ---- (6) stored blocks: short runs become literal tokens, longer ones a reference into the input
        if (st == SS_STORED) {
            if (EMIT && tp_end && emit && tp_end - tp < 16) { emit = false; tok_overflow = 1; }
            const uint64_t ipos = br.consumed() >> 3;
            int err = -1;
            uint32_t n = stored_len;
            if (ipos + n > in_len) { n = (uint32_t)(in_len - ipos); err = ST_NEED_INPUT; }
            if (pos + n > cap) { n = (uint32_t)(cap - pos); err = ST_NEED_OUTPUT; }
            if (n > 8) {
                CZK_FLUSH_LIT();
                CZK_PUT(CZK_TOK_STORED | n);
                CZK_PUT(CZK_TOK_STORED | CZK_TOK_TAIL | (uint32_t)(ipos & 0x1fffffffu));
                CZK_PUT(CZK_TOK_STORED | CZK_TOK_TAIL | (uint32_t)(ipos >> 29));
            } else {
                for (uint32_t k = 0; k < n; k++) {
                    lit |= (uint32_t)in_base[ipos + k] << (8 * nlit);
                    if (++nlit == 3) { CZK_PUT(CZK_TOK_LIT | (3u << 24) | lit); lit = 0; nlit = 0; }
                }
            }
            pos += n;
            br.seek(ipos + n);
            if (err >= 0) { result = err; st = SS_FINISH; }
            else if (bfinal) { result = ST_FINISHED; st = SS_TRAILER; }
            else st = SS_BLOCK;
        }

        // ---- (7) trailer: the check value is compared by phase B, the length check happens here
        if (st == SS_TRAILER) {
            if (Q.runs) {
                // the container trailer of a stream decoded in runs is checked by the host: it needs all runs' lengths and checks
            } else if (!P.segment_mode && result == ST_FINISHED) {
                br.skip((uint32_t)((0 - br.consumed()) & 7));
                if (wrap == 1) {
                    uint32_t t = 0;
                    for (int i = 0; i < 4; i++) t = (t << 8) | br.get_byte();
                    if (br.overrun()) result = ST_NEED_INPUT;
                    expect = t;
                } else if (wrap == 2) {
                    // zlib judges the CRC as soon as its four bytes are there (inflate.c CHECK), before it asks for ISIZE: a unit
                    // cut inside ISIZE is a data error if the CRC is wrong, NeedInput otherwise — phase B decides (wrap bit 2)
                    uint32_t t = 0, isz = 0;
                    for (int i = 0; i < 4; i++) t |= br.get_byte() << (8 * i);
                    if (br.overrun()) result = ST_NEED_INPUT;
                    else {
                        expect = t;
                        for (int i = 0; i < 4; i++) isz |= br.get_byte() << (8 * i);
                        if (br.overrun()) { result = ST_NEED_INPUT; wrap |= 4; }
                        else if (isz != (uint32_t)pos) result = ST_E_DATA;  // "incorrect length check" (the data check comes
                    }                                                       //  first in zlib, but both map to the same code)
                }
            }
            st = SS_FINISH;
        }

        // ---- (8) hand over to phase B
        if (st == SS_FINISH) {
            if (EMIT && tp_end && emit && nlit && tp_end - tp < 2) { emit = false; tok_overflow = 1; }
            CZK_FLUSH_LIT();
            TokMeta m;
            m.ntok = (uint32_t)(tp - tp0);
            m.status = result;
            m.out_len = pos;
            m.expect = expect;
            m.wrap = (uint32_t)wrap;
            Q.meta[unit] = m;
            if (Q.runs) {
                RunResult rr;
                rr.end_bit = br.consumed();
                rr.final_block = result == ST_FINISHED ? 1u : 0u;
                rr.tok_overflow = tok_overflow;
                Q.run_res[unit] = rr;
            }
            if (P.in_consumed) {
                uint64_t c = (br.consumed() + 7) >> 3;
                P.in_consumed[unit] = c < in_len ? c : in_len;
            }
            st = SS_IDLE;
        }
    }
#undef CZK_FLUSH_LIT
#undef CZK_PUT
}

template <int WARPS>
constexpr size_t inflate_tok_smem_bytes() { return 128 + WARPS * 256 + sizeof(LcSlot) * 32 * (size_t)WARPS; }

// ---------------------------------------------------------------------------------------------------------------
// Phase B: one warp per unit.
template <int WARPS, int H>
__global__ void __launch_bounds__(WARPS * 32, (WARPS * CZK_LZ_MINB <= 32 ? CZK_LZ_MINB : 1)) inflate_lz_kernel(TwoPhaseParams Q) {
    const InflateParams &P = Q.base;
    __shared__ uint32_t crc_tab[256 + 34];
    const uint32_t lane = threadIdx.x & 31;
    if (P.crc)
        for (uint32_t i = threadIdx.x; i < 256 + 34; i += WARPS * 32) crc_tab[i] = i < 256 ? P.crc->table[i] : P.crc->pow128[i - 256];
    __syncthreads();
    const uint32_t *crc_pow = crc_tab + 256;

    for (;;) {
        unsigned long long u64 = 0;
        if (lane == 0) u64 = atomicAdd(Q.counter_b, 1ull);
        u64 = __shfl_sync(CZK_FULL, u64, 0);
        if (u64 >= P.n) break;
        const uint32_t unit = P.ids ? P.ids[u64] : (uint32_t)u64;
        const uint64_t o0 = P.out_off[unit];
        if (Q.cta_tile && P.out_off[unit + 1] - o0 <= Q.cta_tile) continue;  // resolved by inflate_lz_cta_kernel
        const TokMeta m = Q.meta[unit];
        uint8_t *ob = P.out + o0;
        const uint8_t *ib = P.in + P.in_off[unit];
        const uint32_t *tok = Q.tok + tok_word_off(o0 - P.out_off[0], unit);
        const uint32_t ntok = m.ntok;
        const bool by_kind = P.segment_mode || (P.checks && m.wrap == 0);  // raw units asked for checks: pieces of a longer stream
        const bool want_adler = by_kind ? (P.check_kind & 1) : m.wrap == 1;
        const bool want_crc = by_kind ? (P.check_kind & 2) : (m.wrap & 3) == 2;
        const bool want_ck = want_adler || want_crc;
        uint64_t opos = 0, ck_pos = 0;
        uint32_t adler = 1, crc = 0;
        uint32_t ti = 0;
        while (ti < ntok) {
            const uint32_t t_raw = ti + lane < ntok ? tok[ti + lane] : CZK_TOK_STORED;  // past the end: acts as a stop mark
            const uint32_t stopm = __ballot_sync(CZK_FULL, (t_raw >> 30) == 3u);
            const uint32_t nt = stopm ? (uint32_t)__ffs(stopm) - 1u : 32u;  // ordinary tokens before the first stored mark
            if (nt) {
                const uint32_t t = lane < nt ? t_raw : 0u;
                const uint32_t tl = (t >> 31) ? ((t >> 24) & 3u) : (t & 0x1ffu);
                // exclusive scan of lengths
                uint32_t pos = tl;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    uint32_t v = __shfl_up_sync(CZK_FULL, pos, d);
                    if ((int)lane >= d) pos += v;
                }
                const uint32_t total = __shfl_sync(CZK_FULL, pos, 31);
                pos -= tl;
                if constexpr (H <= 0) {
                    constexpr int SHORT = H == 0 ? CZK_LZ_SHORT : -H;
                    // ---- token-parallel resolution: lane = token. A token is copied by its own lane (literals; matches of up to
                    // CZK_LZ_SHORT bytes that do not overlap themselves: all source bytes are loaded first, then stored) or by the
                    // whole warp (long or self-overlapping matches). A match may read bytes that earlier tokens of this group
                    // produce, so the group resolves in rounds: everything below the first unfinished token is complete, and a
                    // token is ready when the part of this group it reads lies below that frontier.
                    uint8_t *obp = ob + opos;
                    const bool is_tok = lane < nt;
                    const bool is_lit = is_tok && (t >> 31);
                    const uint32_t dist = (t >> 9) & 0xffffu;
                    const int ipos = (int)pos;
                    const int src0 = ipos - (int)dist;
                    int dep_end = 0;  // bytes of THIS group the token reads end here (exclusive); <= 0: none
                    if (is_tok && !is_lit) { dep_end = src0 + (int)tl; if (dep_end > ipos) dep_end = ipos; }
                    const bool coop = is_tok && !is_lit && (tl > (uint32_t)SHORT || dist < tl);
                    bool done = !is_tok;
                    for (;;) {
                        const uint32_t undone = __ballot_sync(CZK_FULL, !done);
                        if (!undone) break;
                        const int first = __ffs((int)undone) - 1;
                        const int frontier = __shfl_sync(CZK_FULL, ipos, first);
                        const bool ready = !done && dep_end <= frontier;  // (always true for the first unfinished token)
                        if (ready && !coop) {
                            uint8_t *d = obp + ipos;
                            if (is_lit) {
                                d[0] = (uint8_t)t;
                                if (tl > 1) d[1] = (uint8_t)(t >> 8);
                                if (tl > 2) d[2] = (uint8_t)(t >> 16);
                            } else {
                                const uint8_t *sp = obp + src0;
                                if constexpr (SHORT <= 16 && CZK_LZ_WIDE) {
                                    // The source is read as aligned 8-byte words (two on average 1.5 of them instead of one load per
                                    // byte): the L1 data pipe handles one line per lane and load instruction, and it was 73 % busy
                                    // (profiles/r1_inflate_lz_cfg2_final_ncu.md). A word holds at least one wanted byte, so it lies in
                                    // a mapped page; the bytes around the source that other lanes may be writing are discarded.
                                    const uintptr_t a = (uintptr_t)sp;
                                    const uint32_t o = (uint32_t)a & 7u;
                                    const uint2 *q = (const uint2 *)(a - o);
                                    const uint32_t need = o + tl;
#if CZK_LZ_ABL == 1  // ablation (timing experiments only, wrong output): no source loads
                                    uint2 q0 = make_uint2(o, need), q1 = make_uint2(0u, 0u), q2 = make_uint2(0u, 0u);
                                    (void)q;
#else
                                    uint2 q0 = q[0], q1 = make_uint2(0u, 0u), q2 = make_uint2(0u, 0u);
                                    if (need > 8) q1 = q[1];
                                    if (need > 16) q2 = q[2];
#endif
                                    uint32_t w0 = q0.x, w1 = q0.y, w2 = q1.x, w3 = q1.y, w4 = q2.x, w5 = q2.y;
                                    if (o & 4u) { w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; }
                                    const uint32_t sh = (o & 3u) * 8u;
                                    uint32_t b[4];
                                    b[0] = __funnelshift_r(w0, w1, sh);
                                    b[1] = __funnelshift_r(w1, w2, sh);
                                    b[2] = __funnelshift_r(w2, w3, sh);
                                    b[3] = __funnelshift_r(w3, w4, sh);
#if CZK_LZ_ABL == 2  // ablation: one store per token instead of one per byte
                                    d[0] = (uint8_t)(b[0] ^ b[1] ^ b[2] ^ b[3]);
#else
#pragma unroll
                                    for (int k = 0; k < SHORT; k++) if (k < (int)tl) d[k] = (uint8_t)(b[k >> 2] >> (8 * (k & 3)));
#endif
                                } else {
                                    uint32_t bb[SHORT];
#pragma unroll
                                    for (int k = 0; k < SHORT; k++) bb[k] = k < (int)tl ? sp[k] : 0u;
#pragma unroll
                                    for (int k = 0; k < SHORT; k++) if (k < (int)tl) d[k] = (uint8_t)bb[k];
                                }
                            }
                        }
                        uint32_t cm = __ballot_sync(CZK_FULL, ready && coop);
                        while (cm) {
                            const int sl = __ffs((int)cm) - 1;
                            cm &= cm - 1;
                            const uint32_t Ln = __shfl_sync(CZK_FULL, tl, sl), D = __shfl_sync(CZK_FULL, dist, sl);
                            const int P0 = __shfl_sync(CZK_FULL, ipos, sl);
                            uint8_t *d = obp + P0;
                            const uint8_t *sp = d - D;  // bytes [P0 - D, P0) are complete
                            if (D >= Ln) {
                                for (uint32_t k = lane; k < Ln; k += 32) d[k] = sp[k];
                            } else if (D >= 32) {
                                // overlapping, but a 32-byte step only reads what earlier steps wrote
                                for (uint32_t k0 = 0; k0 < Ln; k0 += 32) {
                                    const uint32_t k = k0 + lane;
                                    if (k < Ln) d[k] = sp[k];
                                    __syncwarp();
                                }
                            } else {
                                // short period: byte k repeats source byte k mod D (incremental modulo)
                                uint32_t r = lane % D;
                                const uint32_t stepD = 32u % D;
                                for (uint32_t k = lane; k < Ln; k += 32) {
                                    d[k] = sp[r];
                                    r += stepD;
                                    if (r >= D) r -= D;
                                }
                            }
                        }
                        done = done || ready;
                        __syncwarp();
                    }
                } else {
                    const uint64_t a0 = (uint64_t)(uintptr_t)ob + opos;   // absolute address of batch byte 0
                    int rel = -(int)(a0 & 31);                             // batch-relative index of this iteration's first byte
                    uint32_t cnt_before = 0;
                    uint8_t *obp = ob + opos;
                    // H sector-aligned 32-byte rounds per iteration. All back-reference loads of the iteration are issued before
                    // the first one is consumed (H loads in flight per lane: the kernel is bound by the latency of these
                    // loads, the windows of the streams in flight exceed L2); sources produced inside the iteration come from
                    // registers by shuffle.
                    for (; rel < (int)total; rel += 32 * H) {
                        uint32_t val[H];
                        int srcl[H], srch[H];  // source lane / source round inside this iteration (srch < 0: value is final)
                        const int lo2 = rel > 0 ? rel : 0;
    #pragma unroll
                        for (int h = 0; h < H; h++) {
                            const int r0 = rel + 32 * h;
                            // which tokens start inside this round? (every token covers at least one byte)
                            uint32_t bit = (lane < nt && (int)pos >= r0 && (int)pos < r0 + 32) ? 1u << ((int)pos - r0) : 0u;
                            uint32_t S = __reduce_or_sync(CZK_FULL, bit);
                            const int j = r0 + (int)lane;
                            const bool active = j >= 0 && j < (int)total;
                            // lane holding the token that covers byte j
                            int src_lane = (int)cnt_before + __popc(S & (0xffffffffu >> (31 - lane))) - 1;
                            cnt_before += __popc(S);
                            uint32_t tk = __shfl_sync(CZK_FULL, t, src_lane & 31);
                            uint32_t tpp = __shfl_sync(CZK_FULL, pos, src_lane & 31);
                            val[h] = 0; srcl[h] = 0; srch[h] = -1;
                            if (active) {
                                uint32_t off = (uint32_t)j - tpp;
                                if (tk >> 31) val[h] = (tk >> (8 * off)) & 0xff;
                                else {
                                    uint32_t dist = (tk >> 9) & 0xffffu;
                                    if (off >= dist) {  // overlapping copy: off mod dist without the integer-division sequence (off < 512)
                                        uint32_t q = (uint32_t)__float2uint_rz(__fdividef((float)off, (float)dist));
                                        uint32_t r = off - q * dist;
                                        if ((int)r < 0) r += dist;
                                        if (r >= dist) r -= dist;
                                        off = r;
                                    }
                                    int src = (int)tpp - (int)dist + (int)off;  // batch-relative source index (< tpp)
                                    if (src >= lo2) { srch[h] = (src - rel) >> 5; srcl[h] = (src - rel) & 31; }
                                    else val[h] = obp[src];                     // bytes of earlier iterations / batches
                                }
                            }
                        }
    #pragma unroll
                        for (int h = 0; h < H; h++) {
    #pragma unroll
                            for (int hp = 0; hp < h; hp++) {  // sources in earlier rounds of this iteration: already final
                                uint32_t v = __shfl_sync(CZK_FULL, val[hp], srcl[h]);
                                if (srch[h] == hp) { val[h] = v; srch[h] = -1; }
                            }
                            bool need = srch[h] == h;
                            uint32_t pend = __ballot_sync(CZK_FULL, need);
                            while (pend) {
                                uint32_t v = __shfl_sync(CZK_FULL, val[h], srcl[h]);
                                bool src_ready = !((pend >> srcl[h]) & 1u);
                                if (need && src_ready) { val[h] = v; need = false; }
                                pend = __ballot_sync(CZK_FULL, need);
                            }
                        }
    #pragma unroll
                        for (int h = 0; h < H; h++) {
                            const int j = rel + 32 * h + (int)lane;
                            if (j >= 0 && j < (int)total) obp[j] = (uint8_t)val[h];
                        }
                        __syncwarp();
                    }
                }
                opos += total;
                ti += nt;
            }
            if (nt < 32 && ti < ntok) {
                // a stored run: three marked words starting at ti
                const uint32_t w0 = tok[ti], w1 = tok[ti + 1], w2 = tok[ti + 2];
                const uint32_t n = w0 & 0xffffu;
                const uint64_t ipos = (uint64_t)(w1 & 0x1fffffffu) | ((uint64_t)(w2 & 0x1fffffffu) << 29);
                for (uint32_t k = lane; k < n; k += 32) ob[opos + k] = ib[ipos + k];
                __syncwarp();
                opos += n;
                ti += 3;
            }
            // ---- checksums over freshly written output (L1/L2 hits), in pieces
            if (want_ck && ((uint32_t)(opos - ck_pos) >= 8192u || ti >= ntok)) {
                uint64_t to = opos;
                if (ti < ntok) to = ck_pos + ((to - ck_pos) & ~(uint64_t)127);  // keep CRC pieces at 128 B until the end
                if (want_adler) {
                    for (uint64_t p = ck_pos; p < to; p += 8192) {
                        uint32_t n = (uint32_t)(to - p < 8192 ? to - p : 8192);
                        adler = warp_adler32(adler, ob + p, n, lane);
                    }
                }
                if (want_crc) {
                    uint64_t p = ck_pos;
                    while (to - p >= 128) {
                        uint32_t q = (uint32_t)((to - p) >> 7);
                        if (q > 32) q = 32;
                        crc = warp_crc32_pieces(crc, ob + p, q, crc_tab, crc_pow, lane);
                        p += (uint64_t)q * 128;
                    }
                    if (p < to) {
                        uint32_t c2 = 0;
                        if (lane == 0) c2 = crc32_serial(crc, ob + p, (uint32_t)(to - p), crc_tab);
                        crc = __shfl_sync(CZK_FULL, c2, 0);
                    }
                }
                ck_pos = to;
            }
        }
        if (lane == 0) {
            int status = m.status;
            if (status == ST_FINISHED && !P.segment_mode) {
                if (m.wrap == 1 && m.expect != adler) status = ST_E_DATA;  // "incorrect data check"
                if (m.wrap == 2 && m.expect != crc) status = ST_E_DATA;
            }
            if (m.wrap == 6 && status == ST_NEED_INPUT && m.expect != crc) status = ST_E_DATA;  // gzip cut inside ISIZE, CRC wrong
            P.out_lens[unit] = opos;
            P.statuses[unit] = status;
            if (P.checks) { P.checks[2 * unit] = adler; P.checks[2 * unit + 1] = crc; }
        }
    }
}

#ifdef CZ_EXPERIMENTS  // phase B variants that measured slower than inflate_lz_kernel<8, 0> (profiles/r1_notes.md)
// ---------------------------------------------------------------------------------------------------------------
// Phase B, one warp per unit, TPL tokens per lane (32 * TPL tokens per step). Same resolution rule as the token-parallel
// branch of inflate_lz_kernel (a token is ready when the part of this step it reads lies below the first unfinished token),
// but a step covers TPL times as many tokens: the token load and the scan are paid once per 32 * TPL tokens, the
// back-reference loads of all ready tokens of the step are in flight together, and fewer warps (= fewer 32 KiB windows
// competing for L2) keep the same number of loads in flight.
template <int WARPS, int TPL, int SHORT, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) inflate_lzw_kernel(TwoPhaseParams Q) {
    const InflateParams &P = Q.base;
    __shared__ uint32_t crc_tab[256 + 34];
    const uint32_t lane = threadIdx.x & 31;
    if (P.crc)
        for (uint32_t i = threadIdx.x; i < 256 + 34; i += WARPS * 32) crc_tab[i] = i < 256 ? P.crc->table[i] : P.crc->pow128[i - 256];
    __syncthreads();
    const uint32_t *crc_pow = crc_tab + 256;

    for (;;) {
        unsigned long long u64 = 0;
        if (lane == 0) u64 = atomicAdd(Q.counter_b, 1ull);
        u64 = __shfl_sync(CZK_FULL, u64, 0);
        if (u64 >= P.n) break;
        const uint32_t unit = P.ids ? P.ids[u64] : (uint32_t)u64;
        const uint64_t o0 = P.out_off[unit];
        if (Q.cta_tile && P.out_off[unit + 1] - o0 <= Q.cta_tile) continue;  // resolved by inflate_lz_cta_kernel
        const TokMeta m = Q.meta[unit];
        uint8_t *ob = P.out + o0;
        const uint8_t *ib = P.in + P.in_off[unit];
        const uint32_t *tok = Q.tok + tok_word_off(o0 - P.out_off[0], unit);
        const uint32_t ntok = m.ntok;
        const bool by_kind = P.segment_mode || (P.checks && m.wrap == 0);
        const bool want_adler = by_kind ? (P.check_kind & 1) : m.wrap == 1;
        const bool want_crc = by_kind ? (P.check_kind & 2) : (m.wrap & 3) == 2;
        uint64_t opos = 0, ck_pos = 0;
        uint32_t adler = 1, crc = 0;
        uint32_t ti = 0;
        while (ti < ntok) {
            uint32_t t[TPL];
            uint32_t nt = 32u * TPL;
            bool stop_found = false;
#pragma unroll
            for (int u = 0; u < TPL; u++) {
                const uint32_t i = ti + 32u * u + lane;
                t[u] = i < ntok ? tok[i] : CZK_TOK_STORED;  // past the end: acts as a stop mark
            }
#pragma unroll
            for (int u = 0; u < TPL; u++) {
                const uint32_t stopm = __ballot_sync(CZK_FULL, (t[u] >> 30) == 3u);
                if (!stop_found && stopm) { nt = 32u * u + (uint32_t)__ffs((int)stopm) - 1u; stop_found = true; }
            }
            if (nt) {
                uint32_t tl[TPL];
                int ipos[TPL], dep_end[TPL];
                bool done[TPL], coop[TPL];
                uint32_t carry = 0;
#pragma unroll
                for (int u = 0; u < TPL; u++) {
                    if (32u * u + lane >= nt) t[u] = 0u;
                    tl[u] = (t[u] >> 31) ? ((t[u] >> 24) & 3u) : (t[u] & 0x1ffu);
                    uint32_t pos = tl[u];
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t v = __shfl_up_sync(CZK_FULL, pos, d);
                        if ((int)lane >= d) pos += v;
                    }
                    const uint32_t tot = __shfl_sync(CZK_FULL, pos, 31);
                    ipos[u] = (int)(carry + pos - tl[u]);
                    carry += tot;
                    const bool is_tok = 32u * u + lane < nt;
                    const bool is_match = is_tok && !(t[u] >> 31);
                    const uint32_t dist = (t[u] >> 9) & 0xffffu;
                    dep_end[u] = 0;
                    if (is_match) { dep_end[u] = ipos[u] - (int)dist + (int)tl[u]; if (dep_end[u] > ipos[u]) dep_end[u] = ipos[u]; }
                    coop[u] = is_match && (tl[u] > (uint32_t)SHORT || dist < tl[u]);
                    done[u] = !is_tok;
                }
                const uint32_t total = carry;
                uint8_t *obp = ob + opos;
                for (;;) {
                    int frontier = (int)total;
                    bool any = false;
#pragma unroll
                    for (int u = 0; u < TPL; u++) {
                        const uint32_t und = __ballot_sync(CZK_FULL, !done[u]);
                        if (!any && und) { frontier = __shfl_sync(CZK_FULL, ipos[u], __ffs((int)und) - 1); any = true; }
                    }
                    if (!any) break;
                    bool ready[TPL];
                    uint32_t bb[TPL][SHORT];
#pragma unroll
                    for (int u = 0; u < TPL; u++) {
                        ready[u] = !done[u] && dep_end[u] <= frontier;
                        if (ready[u] && !coop[u] && !(t[u] >> 31)) {
                            const uint8_t *sp = obp + ipos[u] - (int)((t[u] >> 9) & 0xffffu);
#pragma unroll
                            for (int k = 0; k < SHORT; k++) bb[u][k] = k < (int)tl[u] ? sp[k] : 0u;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < TPL; u++) {
                        if (ready[u] && !coop[u]) {
                            uint8_t *d = obp + ipos[u];
                            if (t[u] >> 31) {
                                d[0] = (uint8_t)t[u];
                                if (tl[u] > 1) d[1] = (uint8_t)(t[u] >> 8);
                                if (tl[u] > 2) d[2] = (uint8_t)(t[u] >> 16);
                            } else {
#pragma unroll
                                for (int k = 0; k < SHORT; k++) if (k < (int)tl[u]) d[k] = (uint8_t)bb[u][k];
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < TPL; u++) {
                        uint32_t cm = __ballot_sync(CZK_FULL, ready[u] && coop[u]);
                        while (cm) {
                            const int sl = __ffs((int)cm) - 1;
                            cm &= cm - 1;
                            const uint32_t Ln = __shfl_sync(CZK_FULL, tl[u], sl), D = (__shfl_sync(CZK_FULL, t[u], sl) >> 9) & 0xffffu;
                            const int P0 = __shfl_sync(CZK_FULL, ipos[u], sl);
                            uint8_t *d = obp + P0;
                            const uint8_t *sp = d - D;  // bytes [P0 - D, P0) are complete
                            if (D >= Ln) {
                                for (uint32_t k = lane; k < Ln; k += 32) d[k] = sp[k];
                            } else if (D >= 32) {
                                for (uint32_t k0 = 0; k0 < Ln; k0 += 32) {
                                    const uint32_t k = k0 + lane;
                                    if (k < Ln) d[k] = sp[k];
                                    __syncwarp();
                                }
                            } else {
                                uint32_t r = lane % D;
                                const uint32_t stepD = 32u % D;
                                for (uint32_t k = lane; k < Ln; k += 32) {
                                    d[k] = sp[r];
                                    r += stepD;
                                    if (r >= D) r -= D;
                                }
                            }
                        }
                        done[u] = done[u] || ready[u];
                    }
                    __syncwarp();
                }
                opos += total;
                ti += nt;
            }
            if (nt < 32u * TPL && ti < ntok) {
                // a stored run: three marked words starting at ti
                const uint32_t w0 = tok[ti], w1 = tok[ti + 1], w2 = tok[ti + 2];
                const uint32_t n = w0 & 0xffffu;
                const uint64_t ipos_in = (uint64_t)(w1 & 0x1fffffffu) | ((uint64_t)(w2 & 0x1fffffffu) << 29);
                for (uint32_t k = lane; k < n; k += 32) ob[opos + k] = ib[ipos_in + k];
                __syncwarp();
                opos += n;
                ti += 3;
            }
            if ((want_adler || want_crc) && (opos - ck_pos >= 8192 || ti >= ntok)) {
                uint64_t to = opos;
                if (ti < ntok) to = ck_pos + ((to - ck_pos) & ~(uint64_t)127);  // keep CRC pieces at 128 B until the end
                if (want_adler) {
                    for (uint64_t p = ck_pos; p < to; p += 8192) {
                        uint32_t n = (uint32_t)(to - p < 8192 ? to - p : 8192);
                        adler = warp_adler32(adler, ob + p, n, lane);
                    }
                }
                if (want_crc) {
                    uint64_t p = ck_pos;
                    while (to - p >= 128) {
                        uint32_t q = (uint32_t)((to - p) >> 7);
                        if (q > 32) q = 32;
                        crc = warp_crc32_pieces(crc, ob + p, q, crc_tab, crc_pow, lane);
                        p += (uint64_t)q * 128;
                    }
                    if (p < to) {
                        uint32_t c2 = 0;
                        if (lane == 0) c2 = crc32_serial(crc, ob + p, (uint32_t)(to - p), crc_tab);
                        crc = __shfl_sync(CZK_FULL, c2, 0);
                    }
                }
                ck_pos = to;
            }
        }
        if (lane == 0) {
            int status = m.status;
            if (status == ST_FINISHED && !P.segment_mode) {
                if (m.wrap == 1 && m.expect != adler) status = ST_E_DATA;  // "incorrect data check"
                if (m.wrap == 2 && m.expect != crc) status = ST_E_DATA;
            }
            if (m.wrap == 6 && status == ST_NEED_INPUT && m.expect != crc) status = ST_E_DATA;  // gzip cut inside ISIZE, CRC wrong
            P.out_lens[unit] = opos;
            P.statuses[unit] = status;
            if (P.checks) { P.checks[2 * unit] = adler; P.checks[2 * unit + 1] = crc; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Phase B for units whose output slot fits a shared-memory tile: one CTA per unit.
//
// The whole output of the unit (<= 64 KiB) is assembled in shared memory, so back-references cost a shared-memory
// access instead of an L2/DRAM round trip, and the finished tile leaves with 16-byte coalesced stores. W warps work
// on one unit at a time, W*K*32 tokens per chunk (lane = token, K tokens per lane):
//   * a block scan of the token lengths gives every token its output position;
//   * literals, stored runs and matches whose source lies below the chunk are copied at once;
//   * a match that reads bytes produced by this chunk waits for exactly the tokens that produce them: every token sets
//     a bit in a done mask when its bytes are in the tile, and a position -> token map with one entry per 8 output bytes
//     turns the source range [lo, e) into a (slightly widened) token range [a, b], b < own index. Dependencies only point
//     to lower token indices, so the lowest unfinished token is always ready and the warps never wait on a barrier while
//     they resolve — they poll the done mask. Deflate text needs ~7 dependent levels per 1 024 tokens this way, against
//     ~23 rounds with a single completed-prefix frontier.
// Adler-32 is computed from the tile, CRC-32 from the freshly stored output; W partial values are folded with the combine
// identities.
#define CZK_LZ_TILE 65536u
#define CZK_LZ_SPAN 16384u  // output bytes one chunk may start tokens in (8 bytes per map entry)

template <int W>
constexpr size_t inflate_lz_cta_smem_bytes() { return CZK_LZ_TILE + 16 + (256 + 34) * 4 + 8 + 3 * 32 * 4 + 4 * 8 * 4 + CZK_LZ_SPAN / 8 * 2; }

__device__ __forceinline__ bool lz_deps_done(const volatile uint32_t *dmask, uint32_t dep) {
    const uint32_t a = dep & 0xffffu, b = dep >> 16;
    const uint32_t wa = a >> 5, wb = b >> 5;
    const uint32_t lo_m = 0xffffffffu << (a & 31), hi_m = 0xffffffffu >> (31 - (b & 31));
    if (wa == wb) { const uint32_t mm = lo_m & hi_m; return (dmask[wa] & mm) == mm; }
    if ((dmask[wa] & lo_m) != lo_m) return false;
    if ((dmask[wb] & hi_m) != hi_m) return false;
    for (uint32_t w = wa + 1; w < wb; w++) if (dmask[w] != 0xffffffffu) return false;
    return true;
}

template <int W, int K>
__global__ void __launch_bounds__(W * 32, 3) inflate_lz_cta_kernel(TwoPhaseParams Q) {
    static_assert(W * K == 32, "the block scan keeps one token group per lane");
    constexpr uint32_t T = W * K * 32;
    constexpr int SHORT = CZK_LZ_SHORT;
    constexpr uint32_t DEP_NONE = 0xffffffffu;
    const InflateParams &P = Q.base;
    CZ_DYNAMIC_SMEM(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t *tile_mem = smem_raw;
    uint32_t *crc_tab = (uint32_t *)(smem_raw + CZK_LZ_TILE + 16);
    unsigned long long *s_u = (unsigned long long *)(crc_tab + 256 + 34);
    volatile uint32_t *gtot = (volatile uint32_t *)(s_u + 1);  // [32] bytes of every token group
    volatile uint32_t *dmask = gtot + 32;                       // [32] done bits, one word per token group
    volatile uint32_t *gtot2 = dmask + 32;                      // [32] clipped chunks: bytes | tokens << 20 of every group
    uint32_t *part = (uint32_t *)(gtot2 + 32);                  // [W][4]: adler, crc (shifted to the end), bytes
    uint16_t *map8 = (uint16_t *)(part + 4 * W);                // [SPAN / 8] token that covers chunk byte 8 q
    if (P.crc)
        for (uint32_t i = tid; i < 256 + 34; i += W * 32) crc_tab[i] = i < 256 ? P.crc->table[i] : P.crc->pow128[i - 256];

    for (;;) {
        __syncthreads();  // the previous unit is finished (tile, s_u, part)
        if (tid == 0) *s_u = atomicAdd(Q.counter_c, 1ull);
        __syncthreads();
        const unsigned long long u64 = *s_u;
        if (u64 >= P.n) break;
        const uint32_t unit = P.ids ? P.ids[u64] : (uint32_t)u64;
        const uint64_t o0 = P.out_off[unit];
        const uint64_t cap = P.out_off[unit + 1] - o0;
        if (cap > CZK_LZ_TILE) continue;  // resolved by inflate_lz_kernel
        const TokMeta m = Q.meta[unit];
        uint8_t *ob = P.out + o0;
        const uint8_t *ib = P.in + P.in_off[unit];
        const uint32_t *tok = Q.tok + tok_word_off(o0 - P.out_off[0], unit);
        const uint32_t ntok = m.ntok;
        const uint32_t mis = (uint32_t)((uintptr_t)ob & 15);
        uint8_t *tile = tile_mem + mis;  // tile byte k and output byte k have the same alignment
        uint32_t opos = 0;
        bool bad = false;
        uint32_t ti0 = 0;
        while (ti0 < ntok) {
            uint32_t t[K], tl[K], rel[K];
            // ---- tokens, lengths, scan inside every group of 32
#pragma unroll
            for (int k = 0; k < K; k++) {
                const uint32_t i = ti0 + (warp * K + k) * 32 + lane;
                t[k] = i < ntok ? tok[i] : (CZK_TOK_STORED | CZK_TOK_TAIL);
                const uint32_t top = t[k] >> 29;  // 0 match, 4/5 literals, 6 stored run, 7 tail word of a stored run / padding
                tl[k] = top < 4 ? (t[k] & 0x1ffu) : top < 6 ? ((t[k] >> 24) & 3u) : top == 6 ? (t[k] & 0xffffu) : 0u;
                uint32_t incl = tl[k];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(CZK_FULL, incl, d);
                    if ((int)lane >= d) incl += v;
                }
                rel[k] = incl - tl[k];
                if (lane == 31) gtot[warp * K + k] = incl;
            }
            __syncthreads();  // (A) every warp has left the previous chunk; group totals are visible
            uint32_t gsum = gtot[lane];
            uint32_t gincl = gsum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(CZK_FULL, gincl, d);
                if ((int)lane >= d) gincl += v;
            }
            const uint32_t gexcl = gincl - gsum;
            uint32_t chunk_bytes = __shfl_sync(CZK_FULL, gincl, 31);
            uint32_t nt = T;
#pragma unroll
            for (int k = 0; k < K; k++) rel[k] += __shfl_sync(CZK_FULL, gexcl, warp * K + k);
            const bool clipped = chunk_bytes > CZK_LZ_SPAN;
            if (clipped) {
                // tokens that start beyond the span wait for the next chunk
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const bool in = rel[k] < CZK_LZ_SPAN;
                    if (!in) tl[k] = 0;
                    const uint32_t cnt = __popc(__ballot_sync(CZK_FULL, in));
                    const uint32_t bytes = __reduce_add_sync(CZK_FULL, tl[k]);
                    if (lane == 0) gtot2[warp * K + k] = bytes | (cnt << 20);
                }
            }
            // ---- position -> token map, done bits of the tokens that produce nothing
#pragma unroll
            for (int k = 0; k < K; k++) {
                const uint32_t j = (warp * K + k) * 32 + lane;
                if (tl[k]) {
                    uint32_t p1 = rel[k] + tl[k];
                    if (p1 > CZK_LZ_SPAN) p1 = CZK_LZ_SPAN;
                    for (uint32_t q = (rel[k] + 7) >> 3; 8 * q < p1; q++) map8[q] = (uint16_t)j;
                }
                const uint32_t dm = __ballot_sync(CZK_FULL, tl[k] == 0);
                if (lane == 0) dmask[warp * K + k] = dm;
            }
            __syncthreads();  // (B) map, done mask (and the clipped totals) are visible
            if (clipped) {
                const uint32_t g2 = gtot2[lane];
                const uint32_t s2 = __reduce_add_sync(CZK_FULL, g2);
                chunk_bytes = s2 & 0xfffffu;
                nt = s2 >> 20;
            }
            if (opos + chunk_bytes > CZK_LZ_TILE) { bad = true; break; }  // cannot happen with tokens of inflate_tok_kernel
            const uint32_t map_end = chunk_bytes < CZK_LZ_SPAN ? chunk_bytes : CZK_LZ_SPAN;

            // ---- dependencies
            uint32_t dep[K];
            uint32_t pend = 0;
#pragma unroll
            for (int k = 0; k < K; k++) {
                dep[k] = DEP_NONE;
                const uint32_t j = (warp * K + k) * 32 + lane;
                if (tl[k]) pend |= 1u << k;
                if (tl[k] && (t[k] >> 31) == 0) {
                    const uint32_t dist = (t[k] >> 9) & 0xffffu;
                    const int s = (int)rel[k] - (int)dist;                 // chunk-relative source start (may be negative)
                    int e = s + (int)tl[k];
                    if (e > (int)rel[k]) e = (int)rel[k];                   // a self-overlapping match repeats its own start
                    if (e > 0) {
                        const uint32_t lo = s > 0 ? (uint32_t)s : 0u;
                        const uint32_t a = map8[lo >> 3];
                        const uint32_t qe = (((uint32_t)e - 1u) >> 3) + 1u;
                        uint32_t b = 8 * qe < map_end ? (uint32_t)map8[qe] : j - 1;
                        if (b > j - 1) b = j - 1;
                        dep[k] = a | (b << 16);
                    }
                }
            }

            // ---- resolution
            uint8_t *obp = tile + opos;
            for (;;) {
                bool progress = false;
#pragma unroll
                for (int k = 0; k < K; k++) {
                    bool ready = false;
                    if ((pend >> k) & 1u) ready = dep[k] == DEP_NONE || lz_deps_done(dmask, dep[k]);
                    const uint32_t rm = __ballot_sync(CZK_FULL, ready);
                    if (!rm) continue;
                    __threadfence_block();  // bytes of the tokens whose done bits were just read
                    const uint32_t tk = t[k], len = tl[k];
                    const uint32_t top = tk >> 29;
                    const uint32_t dist = (tk >> 9) & 0xffffu;
                    const bool is_match = top < 4, is_run = top == 6;
                    const bool coop = ready && ((is_match && (len > (uint32_t)SHORT || dist < len)) || is_run);
                    if (ready && !coop) {
                        uint8_t *d = obp + rel[k];
                        if (!is_match) {
                            d[0] = (uint8_t)tk;
                            if (len > 1) d[1] = (uint8_t)(tk >> 8);
                            if (len > 2) d[2] = (uint8_t)(tk >> 16);
                        } else {
                            const uint8_t *sp = d - dist;
                            uint32_t bb[SHORT];
#pragma unroll
                            for (int x = 0; x < SHORT; x++) bb[x] = x < (int)len ? sp[x] : 0u;
#pragma unroll
                            for (int x = 0; x < SHORT; x++) if (x < (int)len) d[x] = (uint8_t)bb[x];
                        }
                    }
                    uint32_t cm = __ballot_sync(CZK_FULL, coop);
                    while (cm) {
                        const int sl = __ffs((int)cm) - 1;
                        cm &= cm - 1;
                        const uint32_t Ln = __shfl_sync(CZK_FULL, len, sl), D = __shfl_sync(CZK_FULL, dist, sl);
                        const uint32_t P0 = __shfl_sync(CZK_FULL, rel[k], sl);
                        const uint32_t topx = __shfl_sync(CZK_FULL, top, sl);
                        uint8_t *d = obp + P0;
                        if (topx == 6) {
                            // stored run: the bytes come from the unit's input
                            const uint32_t i = ti0 + (warp * K + k) * 32 + (uint32_t)sl;
                            const uint32_t w1 = tok[i + 1], w2 = tok[i + 2];
                            const uint8_t *src = ib + ((uint64_t)(w1 & 0x1fffffffu) | ((uint64_t)(w2 & 0x1fffffffu) << 29));
                            for (uint32_t x = lane; x < Ln; x += 32) d[x] = src[x];
                        } else {
                            const uint8_t *sp = d - D;
                            if (D >= Ln) {
                                for (uint32_t x = lane; x < Ln; x += 32) d[x] = sp[x];
                            } else if (D >= 32) {
                                for (uint32_t x0 = 0; x0 < Ln; x0 += 32) {
                                    const uint32_t x = x0 + lane;
                                    if (x < Ln) d[x] = sp[x];
                                    __syncwarp();
                                }
                            } else {
                                uint32_t r = lane % D;
                                const uint32_t stepD = 32u % D;
                                for (uint32_t x = lane; x < Ln; x += 32) {
                                    d[x] = sp[r];
                                    r += stepD;
                                    if (r >= D) r -= D;
                                }
                            }
                        }
                    }
                    __syncwarp();
                    __threadfence_block();  // the bytes are in the tile before the done bits
                    if (lane == 0) atomicOr((uint32_t *)&dmask[warp * K + k], rm);
                    if (ready) pend &= ~(1u << k);
                    progress = true;
                }
                if (!__any_sync(CZK_FULL, pend != 0)) break;
#if defined(__CUDA_ARCH__)
                if (!progress && Q.spin_ns) __nanosleep(Q.spin_ns);
#endif
            }
            opos += chunk_bytes;
            ti0 += nt;
        }
        __syncthreads();  // the tile is complete

        // ---- tile -> output, 16 bytes per lane where the output address is aligned
        const uint32_t total = opos;
        {
            uint32_t head = mis ? 16u - mis : 0u;
            if (head > total) head = total;
            const uint32_t nvec = (total - head) >> 4;
            const uint4 *sv = (const uint4 *)(tile + head);
            uint4 *dv = (uint4 *)(ob + head);
            for (uint32_t v = tid; v < nvec; v += W * 32) dv[v] = sv[v];
            if (tid < head) ob[tid] = tile[tid];
            for (uint32_t x = head + nvec * 16 + tid; x < total; x += W * 32) ob[x] = tile[x];
        }
        // ---- checksums: W slices, folded by thread 0
        const bool by_kind = P.segment_mode || (P.checks && m.wrap == 0);
        const bool want_adler = by_kind ? (P.check_kind & 1) : m.wrap == 1;
        const bool want_crc = by_kind ? (P.check_kind & 2) : (m.wrap & 3) == 2;
        if (want_adler || want_crc) {
            const uint32_t q = ((total + W - 1) / W + 127) & ~127u;
            const uint32_t b0 = warp * q < total ? warp * q : total;
            const uint32_t n = total - b0 < q ? total - b0 : q;
            uint32_t a = 1, c = 0;
            if (want_adler)
                for (uint32_t o = 0; o < n; o += 8192) a = warp_adler32(a, tile + b0 + o, n - o < 8192 ? n - o : 8192, lane);
            if (want_crc) {
                __syncthreads();  // the output bytes of every warp are stored
                const uint8_t *gp = ob + b0;
                uint32_t o = 0;
                while (n - o >= 128) {
                    uint32_t kk = (n - o) >> 7;
                    if (kk > 32) kk = 32;
                    c = warp_crc32_pieces(c, gp + o, kk, crc_tab, crc_tab + 256, lane);
                    o += kk * 128;
                }
                if (lane == 0) {
                    if (o < n) c = crc32_serial(c, gp + o, n - o, crc_tab);
                    c = crc_mulmod(crc_xpow8n(total - (b0 + n)), c);  // shifted to the end of the unit
                }
            }
            if (lane == 0) { part[4 * warp] = a; part[4 * warp + 1] = c; part[4 * warp + 2] = n; }
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t adler = 1, crc = 0;
            if (want_adler) {
                adler = part[0];
                for (int w = 1; w < W; w++) adler = adler32_combine_u(adler, part[4 * w], part[4 * w + 2]);
            }
            if (want_crc)
                for (int w = 0; w < W; w++) crc ^= part[4 * w + 1];
            int status = bad ? (int)ST_E_DATA : m.status;
            if (status == ST_FINISHED && !P.segment_mode) {
                if (m.wrap == 1 && m.expect != adler) status = ST_E_DATA;  // "incorrect data check"
                if (m.wrap == 2 && m.expect != crc) status = ST_E_DATA;
            }
            if (m.wrap == 6 && status == ST_NEED_INPUT && m.expect != crc) status = ST_E_DATA;  // gzip cut inside ISIZE, CRC wrong
            P.out_lens[unit] = total;
            P.statuses[unit] = status;
            if (P.checks) { P.checks[2 * unit] = adler; P.checks[2 * unit + 1] = crc; }
        }
    }
}

#endif  // CZ_EXPERIMENTS

}  // namespace czk
