// inflate_lane_kernel.cuh — batched inflate, "one lane = one stream" variant.
//
// Measured on B200 (profiles/r1_notes.md): per-stream Huffman decoding is a chain of dependent operations
// (~5 cycles per dependent ALU op, ~30 per shared-memory lookup), so the kernel is LATENCY-bound and the throughput is
// set by how many independent streams are in flight per SM, not by instruction issue. This variant therefore gives
// every lane its own stream and keeps ALL per-symbol work lane-local — decode, literal store, LZ77 copy, running
// Adler-32/CRC-32 — so that one warp instruction advances 32 streams and nothing on the per-symbol path is
// serialised across slots. Only block-level work is warp-cooperative: Huffman table construction and stored-block
// copies (coalesced). The number of streams in flight per SM is bounded by shared memory (the per-stream decode
// tables), hence the small tables: 2^LB litlen + 2^DB distance primary entries and a canonical-code fallback for
// longer codes.
//
// Same contract as inflate_kernel.cuh (same InflateParams, same status numbering and zlib error order).
#pragma once
#include "inflate_kernel.cuh"

namespace czk {

template <int LB, int DB>
struct LaneSlot {
    uint16_t lit_tab[1 << LB];    // during header parsing: [0,128) code-length decode table, lens at byte offset 256
    uint16_t dist_tab[1 << DB];
    uint16_t lit_sorted[288];
    uint16_t lit_first[16], lit_offs[16], lit_count[16];
    uint16_t dist_first[16], dist_offs[16], dist_count[16];
    uint8_t dist_sorted[32];
};

// lens scratch must fit behind the 256-byte code-length table inside lit_tab+dist_tab+lit_sorted
template <int LB, int DB>
__device__ __forceinline__ uint8_t *lane_lens(LaneSlot<LB, DB> &sm) { return (uint8_t *)&sm + 256; }

// Warp-cooperative table build. Code lengths are first pulled into registers (the lens scratch aliases the tables).
template <int LB, int DB>
__device__ inline int lane_build_tables(LaneSlot<LB, DB> &sm, uint32_t nlit, uint32_t ndist, uint32_t *wcnt, uint32_t *wrun,
                                        uint32_t lane) {
    const uint8_t *lens = lane_lens(sm);
    uint32_t ll[9];
#pragma unroll
    for (int r = 0; r < 9; r++) { uint32_t sym = r * 32 + lane; ll[r] = sym < nlit ? lens[sym] : 0; }
    uint32_t dl = lane < ndist ? lens[nlit + lane] : 0;
    __syncwarp();
    int rc = 0;
#pragma unroll 1
    for (int which = 0; which < 2; which++) {
        const bool is_dist = which == 1;
        uint16_t *tab = is_dist ? sm.dist_tab : sm.lit_tab;
        const uint32_t tbits = is_dist ? DB : LB;
        uint16_t *first = is_dist ? sm.dist_first : sm.lit_first;
        uint16_t *offs = is_dist ? sm.dist_offs : sm.lit_offs;
        uint16_t *count = is_dist ? sm.dist_count : sm.lit_count;
        if (lane < 16) { wcnt[lane] = 0; wrun[lane] = 0; }
        __syncwarp();
        if (is_dist) { if (dl) atomicAdd(&wcnt[dl], 1u); }
        else {
#pragma unroll
            for (int r = 0; r < 9; r++) if (ll[r]) atomicAdd(&wcnt[ll[r]], 1u);
        }
        __syncwarp();
        int left = 1;
        uint32_t code = 0, off = 0, maxlen = 0, my_first = 0, my_off = 0, my_cnt = 0;
        bool over = false;
        for (uint32_t len = 1; len <= 15; len++) {
            uint32_t c = wcnt[len];
            left = (left << 1) - (int)c;
            if (left < 0) over = true;
            if (lane == len) { my_first = code; my_off = off; my_cnt = c; }
            code = (code + c) << 1;
            off += c;
            if (c) maxlen = len;
        }
        if (over || (left > 0 && maxlen > 1)) { rc = ST_E_DATA; break; }
        if (lane >= 1 && lane < 16) { first[lane] = (uint16_t)my_first; offs[lane] = (uint16_t)my_off; count[lane] = (uint16_t)my_cnt; }
        const uint32_t inval = is_dist ? CZK_D_INVALID : CZK_L_INVALID;
        for (uint32_t i = lane; i < (1u << tbits); i += 32) tab[i] = (uint16_t)inval;
        __syncwarp();
        const int rounds = is_dist ? 1 : 9;
#pragma unroll 1
        for (int r = 0; r < rounds; r++) {
            uint32_t sym = r * 32 + lane;
            uint32_t l = is_dist ? dl : 0;
            if (!is_dist) {
#pragma unroll
                for (int q = 0; q < 9; q++) if (q == r) l = ll[q];
            }
            uint32_t m = __match_any_sync(CZK_FULL, l);
            uint32_t rank = __popc(m & ((1u << lane) - 1u));
            uint32_t base = wrun[l & 15];
            __syncwarp();
            if (l && (m >> lane) <= 1u) wrun[l] = base + __popc(m);
            if (l) {
                uint32_t c = (uint32_t)first[l] + base + rank;
                uint32_t pos = (uint32_t)offs[l] + base + rank;
                if (is_dist) sm.dist_sorted[pos] = (uint8_t)sym; else sm.lit_sorted[pos] = (uint16_t)sym;
                uint32_t rev = __brev(c) >> (32 - l);
                if (l <= tbits) {
                    uint32_t e = is_dist ? dist_entry(sym, l) : lit_entry(sym, l);
                    for (uint32_t idx = rev; idx < (1u << tbits); idx += (1u << l)) tab[idx] = (uint16_t)e;
                } else {
                    tab[rev & ((1u << tbits) - 1u)] = (uint16_t)(is_dist ? CZK_D_LONG : CZK_L_LONG);
                }
            }
            __syncwarp();
        }
    }
    __syncwarp();
    return rc;
}

template <int LB, int DB>
__device__ __forceinline__ uint32_t lane_long_lit(const LaneSlot<LB, DB> &sm, uint32_t bits15) {
    uint32_t code15 = __brev(bits15) >> 17;
    for (uint32_t len = LB + 1; len <= 15; len++) {
        uint32_t d = (code15 >> (15 - len)) - sm.lit_first[len];
        if (d < sm.lit_count[len]) return lit_entry(sm.lit_sorted[sm.lit_offs[len] + d], len);
    }
    return CZK_L_INVALID;
}
template <int LB, int DB>
__device__ __forceinline__ uint32_t lane_long_dist(const LaneSlot<LB, DB> &sm, uint32_t bits15) {
    uint32_t code15 = __brev(bits15) >> 17;
    for (uint32_t len = DB + 1; len <= 15; len++) {
        uint32_t d = (code15 >> (15 - len)) - sm.dist_first[len];
        if (d < sm.dist_count[len]) return dist_entry(sm.dist_sorted[sm.dist_offs[len] + d], len);
    }
    return CZK_D_INVALID;
}

// Lane-local dynamic header parse; same rules as parse_dynamic_header(), lens go to lane_lens(sm).
template <int LB, int DB>
__device__ inline int lane_parse_dynamic(BitReader &br, LaneSlot<LB, DB> &sm, uint32_t &nlit, uint32_t &ndist) {
    br.refill();
    nlit = br.get(5) + 257;
    ndist = br.get(5) + 1;
    uint32_t ncl = br.get(4) + 4;
    if (nlit > 286 || ndist > 30) return ST_E_DATA;
    const uint64_t order_lo = 16ull | (17ull << 5) | (18ull << 10) | (0ull << 15) | (8ull << 20) | (7ull << 25) |
                              (9ull << 30) | (6ull << 35) | (10ull << 40) | (5ull << 45) | (11ull << 50) | (4ull << 55);
    const uint64_t order_hi = 12ull | (3ull << 5) | (13ull << 10) | (2ull << 15) | (14ull << 20) | (1ull << 25) | (15ull << 30);
    uint64_t cl_lens = 0;
    for (uint32_t i = 0; i < ncl; i++) {
        br.refill();
        uint32_t l = br.get(3);
        uint32_t sym = i < 12 ? (uint32_t)(order_lo >> (5 * i)) & 31 : (uint32_t)(order_hi >> (5 * (i - 12))) & 31;
        cl_lens |= (uint64_t)l << (3 * sym);
    }
    if (br.overrun()) return 100;
    // counts per length packed 5 bits each (<= 19 symbols)
    uint64_t cntp = 0;
    for (uint32_t s = 0; s < 19; s++) {
        uint32_t l = (uint32_t)(cl_lens >> (3 * s)) & 7;
        if (l) cntp += 1ull << (5 * l);
    }
    int left = 1;
    uint64_t nextp = 0;  // next code per length, 8 bits each
    uint32_t code = 0;
    for (uint32_t len = 1; len <= 7; len++) {
        uint32_t c = (uint32_t)(cntp >> (5 * len)) & 31;
        left = (left << 1) - (int)c;
        if (left < 0) return ST_E_DATA;
        nextp |= (uint64_t)code << (8 * len);
        code = (code + c) << 1;
    }
    if (left > 0) return ST_E_DATA;
    uint16_t *cl_tab = sm.lit_tab;  // 128 entries = 256 bytes
    for (uint32_t i = 0; i < 128; i++) cl_tab[i] = 0;
    for (uint32_t s = 0; s < 19; s++) {
        uint32_t l = (uint32_t)(cl_lens >> (3 * s)) & 7;
        if (!l) continue;
        uint32_t c = (uint32_t)(nextp >> (8 * l)) & 0xff;
        nextp += 1ull << (8 * l);
        uint32_t rev = __brev(c) >> (32 - l);
        for (uint32_t idx = rev; idx < 128; idx += (1u << l)) cl_tab[idx] = (uint16_t)((s << 3) | l);
    }
    uint32_t total = nlit + ndist, i = 0, prev = 0;
    uint8_t *lens = lane_lens(sm);
    while (i < total) {
        br.refill();
        uint32_t e = cl_tab[br.peek(7)];
        if (!e) return ST_E_DATA;
        br.skip(e & 7);
        uint32_t s = e >> 3;
        if (s < 16) {
            lens[i++] = (uint8_t)s;
            prev = s;
        } else {
            uint32_t rep, val;
            if (s == 16) {
                if (i == 0) return ST_E_DATA;
                rep = 3 + br.get(2);
                val = prev;
            } else if (s == 17) { rep = 3 + br.get(3); val = 0; }
            else { rep = 11 + br.get(7); val = 0; }
            if (i + rep > total) return ST_E_DATA;
            for (uint32_t k = 0; k < rep; k++) lens[i++] = (uint8_t)val;
            prev = val;
        }
        if (br.overrun()) return 100;
    }
    if (lens[256] == 0) return ST_E_DATA;
    return 0;
}

#define CZK_LANE_BUDGET 192  // symbols decoded per lane between two visits of the block-level phases

template <int LB, int DB, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) inflate_lane_kernel(InflateParams P) {
    CZ_DYNAMIC_SMEM(smem_raw);
    static_assert(sizeof(LaneSlot<LB, DB>) >= 256 + 320, "lens scratch does not fit");
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // layout: [crc table 1 KB] [per-warp scratch 128 B] [WARPS*32 slots]
    uint32_t *crc_tab = (uint32_t *)smem_raw;
    uint32_t *wscr = (uint32_t *)(smem_raw + 1024) + warp * 32;
    typedef LaneSlot<LB, DB> Slot;
    Slot *slots = (Slot *)(smem_raw + 1024 + WARPS * 128) + (size_t)warp * 32;
    Slot &my = slots[lane];
    if (P.crc)
        for (uint32_t i = threadIdx.x; i < 256; i += WARPS * 32) crc_tab[i] = P.crc->table[i];
    __syncthreads();

    BitReader br;
    br.words = nullptr; br.mis = 0; br.widx = br.wend = 0; br.cnt = 0; br.buf = 0; br.nextw = 0; br.nextw2 = 0; br.total = 0; br.tail_mask = 0xffffffffu;
    int st = SS_IDLE;
    uint32_t unit = 0;
    const uint8_t *in_base = nullptr;
    uint64_t in_len = 0;
    uint8_t *out = nullptr;
    uint64_t pos = 0, cap = 0;
    uint32_t s1 = 1, s2 = 0, adl_n = 5552, crc = 0xffffffffu;
    int ckmode = 0;  // bit0 adler, bit1 crc
    int result = 0, wrap = 0;
    uint32_t bfinal = 0, nlit = 0, ndist = 0, stored_len = 0;

    for (;;) {
        // ---- (1) fetch work
        if (st == SS_IDLE) {
            unsigned long long u = atomicAdd(P.counter, 1ull);
            if (u >= P.n) st = SS_EXIT;
            else {
                unit = P.ids ? P.ids[u] : (uint32_t)u;
                uint64_t i0 = P.in_off[unit], i1 = P.in_off[unit + 1], o0 = P.out_off[unit], o1 = P.out_off[unit + 1];
                in_base = P.in + i0; in_len = i1 - i0;
                out = P.out + o0; cap = o1 - o0; pos = 0;
                s1 = 1; s2 = 0; adl_n = 5552; crc = 0xffffffffu; bfinal = 0; result = 0;
                br.init(in_base, in_len);
                st = SS_HEADER;
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
            ckmode = P.segment_mode ? (P.check_kind & 3) : wrap;
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
                    else if ((len ^ 0xffffu) != nlen) { result = ST_E_DATA; st = SS_FINISH; }
                    else { stored_len = len; st = SS_STORED; }
                } else if (btype == 1) {
                    uint8_t *lens = lane_lens(my);
                    for (uint32_t i = 0; i < 288; i++) lens[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
                    for (uint32_t i = 0; i < 32; i++) lens[288 + i] = 5;
                    nlit = 288; ndist = 32;
                    st = SS_BUILD;
                } else if (btype == 2) {
                    int r = lane_parse_dynamic(br, my, nlit, ndist);
                    // a verdict reached with bits past the end of the input is not a verdict: zlib would still be waiting
                    if (r == ST_E_DATA && br.consumed() > br.total) r = 100;
                    if (r == 0) st = SS_BUILD;
                    else { result = r == 100 ? ST_NEED_INPUT : r; st = SS_FINISH; }
                } else { result = ST_E_DATA; st = SS_FINISH; }
            }
        }

        // ---- (4) warp-cooperative table construction, one slot at a time
        {
            uint32_t m = __ballot_sync(CZK_FULL, st == SS_BUILD);
            while (m) {
                int s = __ffs(m) - 1;
                m &= m - 1;
                uint32_t nl = __shfl_sync(CZK_FULL, nlit, s), nd = __shfl_sync(CZK_FULL, ndist, s);
                __syncwarp();
                int r = lane_build_tables(slots[s], nl, nd, wscr, wscr + 16, lane);
                if ((int)lane == s) {
                    if (r) { result = ST_E_DATA; st = SS_FINISH; }
                    else st = SS_DECODE;
                }
            }
        }

        // ---- (5) decode + copy, lane-local, CZK_LANE_BUDGET symbols per visit
        if (st == SS_DECODE) {
            int budget = CZK_LANE_BUDGET;
            while (budget-- > 0) {
                if (adl_n < 260) { s1 %= CZK_ADLER_BASE; s2 %= CZK_ADLER_BASE; adl_n = 5552; }
                br.refill();
                uint32_t e = my.lit_tab[br.peek(LB)];
                if ((e & 0xfff0u) == CZK_L_LONG) e = lane_long_lit(my, br.peek(15));
                uint32_t pay = e >> 4;
                if (pay < 0x100) {  // literal
                    br.skip(e & 15);
                    if (br.overrun()) { result = ST_NEED_INPUT; st = SS_FINISH; break; }
                    if (pos >= cap) { result = br.out_full_status(); st = SS_FINISH; break; }
                    out[pos++] = (uint8_t)pay;
                    if (ckmode & 1) { s1 += pay; s2 += s1; adl_n--; }
                    if (ckmode & 2) crc = (crc >> 8) ^ crc_tab[(crc ^ pay) & 0xff];
                    continue;
                }
                if (!(pay & 0x800)) {
                    if (pay == 0x100) {  // end of block
                        br.skip(e & 15);
                        if (br.overrun()) { result = ST_NEED_INPUT; st = SS_FINISH; break; }
                        if (bfinal) { result = ST_FINISHED; st = SS_TRAILER; } else st = SS_BLOCK;
                        break;
                    }
                    result = br.consumed() + ((e & 15) ? (e & 15) : 1) > br.total ? ST_NEED_INPUT : ST_E_DATA;
                    st = SS_FINISH;
                    break;
                }
                br.skip(e & 15);
                uint32_t eb = (pay >> 8) & 7;
                uint32_t len = 3 + (pay & 0xff) + br.peek(eb);
                br.skip(eb);
                br.refill();
                uint32_t de = my.dist_tab[br.peek(DB)];
                if (de & CZK_D_LONG) de = lane_long_dist(my, br.peek(15));
                if (de & CZK_D_INVALID) {
                    result = br.consumed() + ((de & 15) ? (de & 15) : 1) > br.total ? ST_NEED_INPUT : ST_E_DATA;
                    st = SS_FINISH;
                    break;
                }
                br.skip(de & 15);
                uint32_t deb = (de >> 4) & 15;
                uint32_t dist = (((de >> 8) & 3) << deb) + 1 + br.peek(deb);
                br.skip(deb);
                if (br.overrun()) { result = ST_NEED_INPUT; st = SS_FINISH; break; }
                // zlib order (inflate.c MATCH): output space first, then "invalid distance too far back"
                if (pos >= cap) { result = br.out_full_status(); st = SS_FINISH; break; }
                if ((uint64_t)dist > pos) { result = ST_E_DATA; st = SS_FINISH; break; }
                uint32_t n = len;
                if (pos + n > cap) n = (uint32_t)(cap - pos);
                uint8_t *d = out + pos;
                const uint8_t *s = d - dist;
                uint32_t k = 0;
                if (dist >= 4) {
                    // groups of 4: the loads of a group never depend on its stores
                    for (; k + 4 <= n; k += 4) {
                        uint32_t b0 = s[k], b1 = s[k + 1], b2 = s[k + 2], b3 = s[k + 3];
                        d[k] = (uint8_t)b0; d[k + 1] = (uint8_t)b1; d[k + 2] = (uint8_t)b2; d[k + 3] = (uint8_t)b3;
                        if (ckmode & 1) { s1 += b0; s2 += s1; s1 += b1; s2 += s1; s1 += b2; s2 += s1; s1 += b3; s2 += s1; }
                        if (ckmode & 2) {
                            crc = (crc >> 8) ^ crc_tab[(crc ^ b0) & 0xff]; crc = (crc >> 8) ^ crc_tab[(crc ^ b1) & 0xff];
                            crc = (crc >> 8) ^ crc_tab[(crc ^ b2) & 0xff]; crc = (crc >> 8) ^ crc_tab[(crc ^ b3) & 0xff];
                        }
                    }
                }
                for (; k < n; k++) {
                    uint32_t b = s[k];
                    d[k] = (uint8_t)b;
                    if (ckmode & 1) { s1 += b; s2 += s1; }
                    if (ckmode & 2) crc = (crc >> 8) ^ crc_tab[(crc ^ b) & 0xff];
                }
                adl_n -= n;
                pos += n;
                if (n < len) { result = br.out_full_status(); st = SS_FINISH; break; }
            }
        }

        // ---- (6) stored blocks: warp-cooperative copy; the checksum of the copied bytes is folded by the warp
        {
            uint32_t m = __ballot_sync(CZK_FULL, st == SS_STORED);
            while (m) {
                int s = __ffs(m) - 1;
                m &= m - 1;
                uint32_t len = __shfl_sync(CZK_FULL, stored_len, s);
                const uint8_t *ib = (const uint8_t *)(uintptr_t)__shfl_sync(CZK_FULL, (unsigned long long)(uintptr_t)in_base, s);
                uint8_t *ob = (uint8_t *)(uintptr_t)__shfl_sync(CZK_FULL, (unsigned long long)(uintptr_t)out, s);
                const uint64_t opos = __shfl_sync(CZK_FULL, (unsigned long long)pos, s);
                const uint64_t ocap = __shfl_sync(CZK_FULL, (unsigned long long)cap, s);
                const uint64_t ipos = __shfl_sync(CZK_FULL, (unsigned long long)br.consumed(), s) >> 3;
                const uint64_t ilen = __shfl_sync(CZK_FULL, (unsigned long long)in_len, s);
                const int ck = __shfl_sync(CZK_FULL, ckmode, s);
                int err = -1;
                uint32_t n = len;
                if (ipos + n > ilen) { n = (uint32_t)(ilen - ipos); err = ST_NEED_INPUT; }
                if (opos + n > ocap) { n = (uint32_t)(ocap - opos); err = ST_NEED_OUTPUT; }
                // adler over the block: sum(b) and sum((n-k) b), in pieces of <= 4096 bytes to stay inside 32 bits
                uint32_t a1 = __shfl_sync(CZK_FULL, s1, s) % CZK_ADLER_BASE, a2 = __shfl_sync(CZK_FULL, s2, s) % CZK_ADLER_BASE;
                for (uint32_t base = 0; base < n; base += 4096) {
                    uint32_t m_ = n - base < 4096 ? n - base : 4096;
                    uint32_t p1 = 0, p2 = 0;
                    for (uint32_t k = lane; k < m_; k += 32) {
                        uint32_t b = ib[ipos + base + k];
                        ob[opos + base + k] = (uint8_t)b;
                        p1 += b;
                        p2 += (m_ - k) * b;
                    }
                    if (ck & 1) {
                        p2 %= CZK_ADLER_BASE;
                        p1 = __reduce_add_sync(CZK_FULL, p1);
                        p2 = __reduce_add_sync(CZK_FULL, p2);
                        a2 = (a2 + m_ * a1 + p2) % CZK_ADLER_BASE;
                        a1 = (a1 + p1) % CZK_ADLER_BASE;
                    }
                }
                __syncwarp();
                if ((int)lane == s) {
                    if (ck & 1) { s1 = a1; s2 = a2; adl_n = 5552; }
                    if (ck & 2) {  // CRC of a stored block: serial on the owning lane from the bytes just written
                        for (uint32_t k = 0; k < n; k++) crc = (crc >> 8) ^ crc_tab[(crc ^ ob[opos + k]) & 0xff];
                    }
                    pos = opos + n;
                    br.seek(ipos + n);
                    if (err >= 0) { result = err; st = SS_FINISH; }
                    else if (bfinal) { result = ST_FINISHED; st = SS_TRAILER; }
                    else st = SS_BLOCK;
                }
            }
        }

        // ---- (7) trailer
        if (st == SS_TRAILER) {
            s1 %= CZK_ADLER_BASE; s2 %= CZK_ADLER_BASE; adl_n = 5552;
            if (!P.segment_mode && result == ST_FINISHED) {
                br.skip((uint32_t)((0 - br.consumed()) & 7));
                if (wrap == 1) {
                    uint32_t v = 0;
                    for (int i = 0; i < 4; i++) v = (v << 8) | br.get_byte();
                    if (br.overrun()) result = ST_NEED_INPUT;
                    else if (v != ((s2 << 16) | s1)) result = ST_E_DATA;
                } else if (wrap == 2) {
                    uint32_t v = 0, isz = 0;
                    for (int i = 0; i < 4; i++) v |= br.get_byte() << (8 * i);
                    for (int i = 0; i < 4; i++) isz |= br.get_byte() << (8 * i);
                    if (br.overrun()) result = ST_NEED_INPUT;
                    else if (v != (crc ^ 0xffffffffu)) result = ST_E_DATA;
                    else if (isz != (uint32_t)pos) result = ST_E_DATA;
                }
            }
            st = SS_FINISH;
        }

        // ---- (8) report
        if (st == SS_FINISH) {
            P.out_lens[unit] = pos;
            P.statuses[unit] = result;
            if (P.in_consumed) {
                uint64_t c = (br.consumed() + 7) >> 3;
                P.in_consumed[unit] = c < in_len ? c : in_len;
            }
            if (P.checks) {
                P.checks[2 * unit] = ((s2 % CZK_ADLER_BASE) << 16) | (s1 % CZK_ADLER_BASE);
                P.checks[2 * unit + 1] = crc ^ 0xffffffffu;
            }
            st = SS_IDLE;
        }
    }
}

template <int LB, int DB, int WARPS>
constexpr size_t inflate_lane_smem_bytes() { return 1024 + WARPS * 128 + sizeof(LaneSlot<LB, DB>) * 32 * (size_t)WARPS; }

}  // namespace czk
