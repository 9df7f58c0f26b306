// inflate_lc_kernel.cuh — batched inflate, "one lane = one stream, canonical decode" variant (the default).
//
// What the ncu captures of the earlier variants showed (profiles/r1_inflate_notes.md):
//   * one-decoder-lane-per-warp designs spend ~50 warp instructions per decoded symbol with ~2 of 32 lanes active;
//     they are issue-bound at a few % of what the SM can do;
//   * giving every lane its own stream fixes the issue efficiency, but with a 2^9/2^10-entry lookup table per stream
//     only 3-4 warps fit in shared memory per SM, and a serial decode chain with 3 warps per SM is latency-bound.
// So this variant keeps NO lookup table: Huffman codes are decoded canonically. The 15 per-length code limits of the
// literal/length code and of the distance code live in REGISTERS (30 registers per lane); a symbol is found with 15
// independent compares (code length = 1 + #limits <= the bit-reversed 16-bit window), one 32-byte base-offset lookup
// and one lookup in the canonically sorted symbol array. Per-stream shared memory drops to 448 bytes, so 14 warps =
// 448 concurrent streams run per SM and the per-symbol latency chain is hidden by other warps.
//
// Everything on the per-symbol path is lane-local (decode, literal store, LZ77 copy, running Adler-32 / CRC-32); only
// table construction and stored-block copies are warp-cooperative. Contract, status numbering and zlib's error order
// are identical to inflate_kernel.cuh.
#pragma once
#include "inflate_kernel.cuh"

namespace czk {

struct LcSlot {
    // while a block header is parsed: u8 cl_tab[128] at byte 0, u8 lens[320] at byte 128
    uint8_t lit_sorted[288];  // low 8 bits of the canonically sorted literal/length symbols
    uint32_t lit_info[16];    // per code length: (offs - first_code) mod 2^16 | (first sorted index holding a symbol >= 256) << 16
    uint16_t dist_base[16];   // offs[len] - first_code[len]  (mod 2^16)
    uint8_t dist_sorted[32];
    uint8_t pad[36];          // 452 B = 113 words: an odd word stride, so the 32 lanes' slots start in 32 different banks
};
static_assert(sizeof(LcSlot) == 452, "LcSlot layout");

__device__ __forceinline__ uint8_t *lc_lens(LcSlot &sm) { return (uint8_t *)&sm + 128; }

// length base (9 bits) | extra-bit count << 9 for length symbols 257..285, kept in shared memory per CTA
__host__ __device__ inline uint32_t lc_len_info(uint32_t c) {
    uint32_t e, base;
    if (c < 8) { e = 0; base = 3 + c; }
    else if (c == 28) { e = 0; base = 258; }
    else { e = (c >> 2) - 1; base = 3 + ((4 + (c & 3)) << e); }
    return base | (e << 9);
}

// The 15 code limits of one code (left-justified to 16 bits, non-decreasing; 65536 = "no code can reach this") live in 8
// registers. A limit for length L <= 14 has its low two bits clear, so `v >= limit` can be decided on the 14-bit values
// v >> 2 and limit >> 2; biased by 0x400 these are bit patterns of finite, normal fp16 numbers (0x0400..0x4400) whose
// order as numbers is their order as integers. pk[j] (j < 7) holds the limits of lengths 2j+1 (low half) and 2j+2 (high
// half) in that form: one HSET2 compares two limits, and the FP16 pipe does the counting instead of the integer ALU.
// pk[7] is the limit of length 15 as a plain integer.
#define CZK_LC_PK_NONE 0x44004400u  // both halves "unreachable"
__device__ __forceinline__ void lc_limits_reset(uint32_t (&pk)[8]) {
#pragma unroll
    for (int i = 0; i < 7; i++) pk[i] = CZK_LC_PK_NONE;
    pk[7] = 0x10000u;
}

// Warp-cooperative: builds slot `sm` (sorted symbols + base offsets) from the code lengths in its scratch and hands the
// 2x15 code limits to lane `owner` in the packed form above.
__device__ __forceinline__ int lc_build(LcSlot &sm, uint32_t nlit, uint32_t ndist, uint32_t *wcnt, uint32_t *wrun, uint32_t lane,
                               uint32_t owner, uint32_t (&llim)[8], uint32_t (&dlim)[8]) {
    const uint8_t *lens = lc_lens(sm);
    uint32_t ll[9];
#pragma unroll
    for (int r = 0; r < 9; r++) { uint32_t sym = r * 32 + lane; ll[r] = sym < nlit ? lens[sym] : 0; }
    uint32_t dl = lane < ndist ? lens[nlit + lane] : 0;
    __syncwarp();
    int rc = 0;
#pragma unroll
    for (int which = 0; which < 2; which++) {
        const bool is_dist = which == 1;
        if (lane < 16) { wcnt[lane] = 0; wrun[lane] = 0; }
        if (lane >= 16) wrun[lane] = 0;  // wrun[16..31]: per length, number of symbols < 256
        __syncwarp();
        if (is_dist) { if (dl) atomicAdd(&wcnt[dl], 1u); }
        else {
#pragma unroll
            for (int r = 0; r < 9; r++) if (ll[r]) { atomicAdd(&wcnt[ll[r]], 1u); if (r < 8) atomicAdd(&wrun[16 + ll[r]], 1u); }
        }
        __syncwarp();
        int left = 1;
        uint32_t code = 0, off = 0, maxlen = 0, my_first = 0, my_off = 0;
        bool over = false;
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; i++) pk[i] = 0;
#pragma unroll
        for (uint32_t len = 1; len <= 15; len++) {
            uint32_t c = wcnt[len];
            left = (left << 1) - (int)c;
            if (left < 0) over = true;
            if (lane == len) { my_first = code; my_off = off; }
            uint32_t lim = (code + c) << (16 - len);
            if (lim > 0x10000u) lim = 0x10000u;  // over-subscribed sets are rejected below; keep the packed fields in range
            if (len < 15) pk[(len - 1) >> 1] |= ((lim >> 2) + 0x400u) << (16 * ((len - 1) & 1));
            else pk[7] = lim;
            code = (code + c) << 1;
            off += c;
            if (c) maxlen = len;
        }
        if (lane == owner) {
#pragma unroll
            for (int i = 0; i < 8; i++) { if (is_dist) dlim[i] = pk[i]; else llim[i] = pk[i]; }
        }
        if (over || (left > 0 && maxlen > 1)) rc = ST_E_DATA;  // zlib inflate_table(): over-subscribed / incomplete set
        if (lane >= 1 && lane < 16) {
            if (is_dist) sm.dist_base[lane] = (uint16_t)(my_off - my_first);
            else sm.lit_info[lane] = ((my_off - my_first) & 0xffffu) | ((my_off + wrun[16 + lane]) << 16);
        }
        __syncwarp();
        const int rounds = is_dist ? 1 : 9;
#pragma unroll
        for (int r = 0; r < rounds; r++) {
            uint32_t sym = r * 32 + lane;
            uint32_t l = is_dist ? dl : ll[r];
            uint32_t m = __match_any_sync(CZK_FULL, l);
            uint32_t rank = __popc(m & ((1u << lane) - 1u));
            uint32_t base = wrun[l & 15];
            __syncwarp();
            if (l && (m >> lane) <= 1u) wrun[l] = base + __popc(m);
            if (l && !rc) {
                // position of this symbol in canonical order = offs[l] + (symbols of the same length before it)
                uint32_t offs_l = 0;
#pragma unroll
                for (uint32_t len = 1; len <= 15; len++) offs_l += len < l ? wcnt[len] : 0;
                uint32_t pos = offs_l + base + rank;
                if (is_dist) sm.dist_sorted[pos & 31] = (uint8_t)sym; else if (pos < 288) sm.lit_sorted[pos] = (uint8_t)sym;
            }
            __syncwarp();
        }
    }
    __syncwarp();
    return rc;
}

// code length of the codeword at the top of the 16-bit left-justified window v: 1 + number of limits <= v (16 = invalid)
__device__ __forceinline__ uint32_t lc_code_len(uint32_t v, const uint32_t (&pk)[8]) {
#if defined(__CUDA_ARCH__)
    const uint32_t vv = (v >> 2) * 0x10001u + 0x04000400u;
    const __half2 hv = *reinterpret_cast<const __half2 *>(&vv);
    __half2 c[7];
#pragma unroll
    for (int i = 0; i < 7; i++) c[i] = __hge2(hv, *reinterpret_cast<const __half2 *>(&pk[i]));  // 1.0 where v >= limit
    const uint32_t magic = 0x64006400u;  // (1024, 1024): 1024 + k has the bit pattern 0x6400 + k
    __half2 s = __hadd2(__hadd2(__hadd2(c[0], c[1]), __hadd2(c[2], c[3])),
                        __hadd2(__hadd2(c[4], c[5]), __hadd2(c[6], *reinterpret_cast<const __half2 *>(&magic))));
    const uint32_t sb = *reinterpret_cast<const uint32_t *>(&s);
    return 1u + ((sb + (sb >> 16)) & 0xfu) + (v >= pk[7] ? 1u : 0u);
#else
    uint32_t len = 1;
    const uint32_t v14 = (v >> 2) + 0x400u;
#pragma unroll
    for (int i = 0; i < 14; i++) len += v14 >= ((pk[i >> 1] >> (16 * (i & 1))) & 0xffffu) ? 1u : 0u;
    return len + (v >= pk[7] ? 1u : 0u);
#endif
}

// Lane-local dynamic header parse (RFC 1951 §3.2.7) into lc_lens(sm); 0, ST_E_DATA or 100 (= input exhausted).
__device__ inline int lc_parse_dynamic(BitReader &br, LcSlot &sm, uint32_t &nlit, uint32_t &ndist) {
    br.refill();
    nlit = br.get(5) + 257;
    ndist = br.get(5) + 1;
    uint32_t ncl = br.get(4) + 4;
    if (nlit > 286 || ndist > 30) return ST_E_DATA;  // "too many length or distance symbols"
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
    uint64_t cntp = 0;  // symbols per length, 5 bits each
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
    if (left > 0) return ST_E_DATA;  // zlib: an incomplete code-length code is always an error
    uint8_t *cl_tab = sm.lit_sorted;  // 128 one-byte entries: symbol << 3 | length
    for (uint32_t i = 0; i < 128; i++) cl_tab[i] = 0;
    for (uint32_t s = 0; s < 19; s++) {
        uint32_t l = (uint32_t)(cl_lens >> (3 * s)) & 7;
        if (!l) continue;
        uint32_t c = (uint32_t)(nextp >> (8 * l)) & 0xff;
        nextp += 1ull << (8 * l);
        uint32_t rev = __brev(c) >> (32 - l);
        for (uint32_t idx = rev; idx < 128; idx += (1u << l)) cl_tab[idx] = (uint8_t)((s << 3) | l);
    }
    uint32_t total = nlit + ndist, i = 0, prev = 0;
    uint8_t *lens = lc_lens(sm);
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
                if (i == 0) return ST_E_DATA;  // "invalid bit length repeat"
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
    if (lens[256] == 0) return ST_E_DATA;  // "invalid code -- missing end-of-block"
    return 0;
}

#define CZK_LC_BUDGET 256  // symbols decoded per lane between two visits of the block-level phases

#ifdef CZ_EXPERIMENTS  // the single-kernel lane-per-stream design (77 GB/s on cfg2: 65 536 windows thrash L2); the default path
                       // only uses the slot layout, the table construction and the canonical decode above

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) inflate_lc_kernel(InflateParams P) {
    CZ_DYNAMIC_SMEM(smem_raw);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // layout: [crc table 1 KB] [length info 128 B] [per-warp scratch 256 B each] [WARPS*32 slots]
    uint32_t *crc_tab = (uint32_t *)smem_raw;
    uint16_t *len_info = (uint16_t *)(smem_raw + 1024);
    uint32_t *wscr = (uint32_t *)(smem_raw + 1152) + warp * 64;  // wcnt[16] wrun[16] wnlo[16] pad
    LcSlot *slots = (LcSlot *)(smem_raw + 1152 + WARPS * 256) + (size_t)warp * 32;
    LcSlot &my = slots[lane];
    if (P.crc)
        for (uint32_t i = threadIdx.x; i < 256; i += WARPS * 32) crc_tab[i] = P.crc->table[i];
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
                    else if ((len ^ 0xffffu) != nlen) { result = ST_E_DATA; st = SS_FINISH; }  // "invalid stored block lengths"
                    else { stored_len = len; st = SS_STORED; }
                } else if (btype == 1) {
                    uint8_t *lens = lc_lens(my);
                    for (uint32_t i = 0; i < 288; i++) lens[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
                    for (uint32_t i = 0; i < 32; i++) lens[288 + i] = 5;
                    nlit = 288; ndist = 32;
                    st = SS_BUILD;
                } else if (btype == 2) {
                    int r = lc_parse_dynamic(br, my, nlit, ndist);
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
                uint32_t nl = __shfl_sync(CZK_FULL, nlit, s), nd = __shfl_sync(CZK_FULL, ndist, s);
                __syncwarp();
                int r = lc_build(slots[s], nl, nd, wscr, wscr + 16, lane, (uint32_t)s, llim, dlim);
                if ((int)lane == s) {
                    if (r) { result = ST_E_DATA; st = SS_FINISH; }
                    else st = SS_DECODE;
                }
            }
        }

        // ---- (5) decode + copy, lane-local, CZK_LC_BUDGET symbols per visit
        if (st == SS_DECODE) {
            int budget = CZK_LC_BUDGET;
            while (budget-- > 0) {
                if (adl_n < 260) { s1 %= CZK_ADLER_BASE; s2 %= CZK_ADLER_BASE; adl_n = 5552; }
                br.refill();
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
                    if (br.overrun()) { result = ST_NEED_INPUT; st = SS_FINISH; break; }
                    if (pos >= cap) { result = br.out_full_status(); st = SS_FINISH; break; }
                    out[pos++] = (uint8_t)sym;
                    if (ckmode & 1) { s1 += sym; s2 += s1; adl_n--; }
                    if (ckmode & 2) crc = (crc >> 8) ^ crc_tab[(crc ^ sym) & 0xff];
                    continue;
                }
                if (sym == 256) {  // end of block
                    if (br.overrun()) { result = ST_NEED_INPUT; st = SS_FINISH; break; }
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
                if (br.overrun()) { result = ST_NEED_INPUT; st = SS_FINISH; break; }
                // zlib order (inflate.c MATCH): output space first, then "invalid distance too far back"
                if (pos >= cap) { result = br.out_full_status(); st = SS_FINISH; break; }
                if ((uint64_t)dist > pos) { result = ST_E_DATA; st = SS_FINISH; break; }
                uint32_t n = len;
                if (pos + n > cap) n = (uint32_t)(cap - pos);
                uint8_t *d = out + pos;
                const uint8_t *s = d - dist;
                uint32_t k = 0;
                if (dist >= 8) {
                    // groups of 4 bytes, software-pipelined: the loads of group k+4 are issued before the stores of
                    // group k (legal because dist >= 8: they only read bytes stored by groups < k)
                    uint32_t b0 = s[0], b1 = n > 1 ? s[1] : 0, b2 = n > 2 ? s[2] : 0, b3 = n > 3 ? s[3] : 0;
                    for (; k < n; k += 4) {
                        uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                        if (k + 4 < n) c0 = s[k + 4];
                        if (k + 5 < n) c1 = s[k + 5];
                        if (k + 6 < n) c2 = s[k + 6];
                        if (k + 7 < n) c3 = s[k + 7];
                        uint32_t m_ = n - k;  // bytes of this group that exist: min(4, n-k)
                        d[k] = (uint8_t)b0;
                        if (m_ > 1) d[k + 1] = (uint8_t)b1;
                        if (m_ > 2) d[k + 2] = (uint8_t)b2;
                        if (m_ > 3) d[k + 3] = (uint8_t)b3;
                        if (ckmode & 1) {
                            s1 += b0; s2 += s1;
                            if (m_ > 1) { s1 += b1; s2 += s1; }
                            if (m_ > 2) { s1 += b2; s2 += s1; }
                            if (m_ > 3) { s1 += b3; s2 += s1; }
                        }
                        if (ckmode & 2) {
                            crc = (crc >> 8) ^ crc_tab[(crc ^ b0) & 0xff];
                            if (m_ > 1) crc = (crc >> 8) ^ crc_tab[(crc ^ b1) & 0xff];
                            if (m_ > 2) crc = (crc >> 8) ^ crc_tab[(crc ^ b2) & 0xff];
                            if (m_ > 3) crc = (crc >> 8) ^ crc_tab[(crc ^ b3) & 0xff];
                        }
                        b0 = c0; b1 = c1; b2 = c2; b3 = c3;
                    }
                } else {
                    for (; k < n; k++) {  // overlapping copy: strictly sequential
                        uint32_t b = s[k];
                        d[k] = (uint8_t)b;
                        if (ckmode & 1) { s1 += b; s2 += s1; }
                        if (ckmode & 2) crc = (crc >> 8) ^ crc_tab[(crc ^ b) & 0xff];
                    }
                }
                adl_n -= n;
                pos += n;
                if (n < len) { result = br.out_full_status(); st = SS_FINISH; break; }
            }
        }

        // ---- (6) stored blocks: warp-cooperative copy; Adler-32 of the copied bytes is folded by the warp
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
                    if (ck & 2) {
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
                    uint32_t t = 0;
                    for (int i = 0; i < 4; i++) t = (t << 8) | br.get_byte();
                    if (br.overrun()) result = ST_NEED_INPUT;
                    else if (t != ((s2 << 16) | s1)) result = ST_E_DATA;  // "incorrect data check"
                } else if (wrap == 2) {
                    uint32_t t = 0, isz = 0;
                    for (int i = 0; i < 4; i++) t |= br.get_byte() << (8 * i);
                    for (int i = 0; i < 4; i++) isz |= br.get_byte() << (8 * i);
                    if (br.overrun()) result = ST_NEED_INPUT;
                    else if (t != (crc ^ 0xffffffffu)) result = ST_E_DATA;  // "incorrect data check"
                    else if (isz != (uint32_t)pos) result = ST_E_DATA;       // "incorrect length check"
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

template <int WARPS>
constexpr size_t inflate_lc_smem_bytes() { return 1152 + WARPS * 256 + sizeof(LcSlot) * 32 * (size_t)WARPS; }
#endif  // CZ_EXPERIMENTS

}  // namespace czk
