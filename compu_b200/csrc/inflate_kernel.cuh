// inflate_kernel.cuh — batched DEFLATE / zlib / gzip inflate for sm_100a.
//
// Replaces what compu outsources to L0 `inflate()` behind decode_fn
// (/root/reference/src/decoder/zlib_ng.rs:94-96 -> src/decoder/mod.rs:459-486): header parse, Huffman table build,
// symbol decode, LZ77 copy, Adler-32 / CRC-32 verification — for MANY independent streams (or full-flush segments)
// per launch. Formats per RFC 1950/1951/1952; error numbering follows zlib (-3 data error, 3 = need dictionary).
//
// Mapping to the hardware
//   * A warp owns D "slots"; lane s (< D) is the sequential Huffman decoder of slot s and keeps the slot's bit
//     reader in registers. Each slot has its own decode tables in shared memory (~3.7 KB, SlotSmem).
//     D lanes decode D different streams in the same instruction stream (SIMT), which is where the throughput
//     over a one-decoder-per-warp design comes from.
//   * The decoder lanes only produce 32-bit tokens (literal | length,distance) into a shared-memory queue.
//     All 32 lanes then resolve the queue of ONE slot at a time: a warp scan turns token lengths into output
//     positions and the bytes are produced 32 at a time, sector-aligned, so every global store is a fully
//     coalesced 32-byte sector write; back-references are coalesced loads of bytes this warp wrote earlier, or are
//     resolved in registers with shuffles when the source lies inside the same 32-byte round.
//   * Dynamic-block headers are parsed by the slot's lane; table construction is warp-cooperative.
//   * Adler-32 / CRC-32 of the output are computed by the warp from the bytes it has just written (L1/L2 hits), in
//     chunks, so the checksum costs no extra HBM pass.
//   * Slots pull work from a global atomic counter (persistent warps), so ragged stream sizes balance themselves.
#pragma once
#include "czk_common.cuh"

namespace czk {

#ifndef CZK_LIT_BITS
#define CZK_LIT_BITS 10
#endif
#ifndef CZK_DIST_BITS
#define CZK_DIST_BITS 8
#endif
#define CZK_TOKENS 32

// Resumable decoding of ONE long-lived stream (the streaming Decoder, cz_decode): a unit may stop anywhere zlib's inflate()
// may return — out of input in the middle of a header, a code or a stored run; output slot full in the middle of a match —
// and the next launch picks up exactly there. What carries over is what zlib keeps in its inflate_state: the bit position, the
// current block's code lengths, a token that did not fit yet, the running check values, and the last 32 KiB of output (which the
// host keeps right in front of the unit's output slot).
enum ResumePhase : uint32_t {
    RP_HEADER = 0,   // container header not parsed yet (restart at bit_pos; the header is re-read whole)
    RP_BLOCK = 1,    // a block header comes next
    RP_CODES = 2,    // inside a fixed / dynamic block: lens[] holds its code lengths
    RP_STORED = 3,   // inside a stored block: stored_left bytes still to copy
    RP_TRAILER = 4,  // all blocks done, the container trailer comes next
    RP_DONE = 5
};
struct ResumeState {
    uint64_t bit_pos;     // in: where decoding restarts, in bits from the start of the unit's input; out: where to restart next
    uint64_t total_out;   // bytes produced by earlier launches (ISIZE check)
    uint32_t phase;       // ResumePhase
    uint32_t bfinal;      // BFINAL of the current block
    uint32_t nlit, ndist; // RP_CODES: alphabet sizes of lens[]
    uint32_t stored_left; // RP_STORED
    uint32_t wrap;        // container kind once the header has been parsed (0 raw, 1 zlib, 2 gzip)
    uint32_t hist_len;    // valid history bytes right in front of the output slot (<= 32768)
    uint32_t adler, crc;  // running check values over everything produced so far
    uint32_t pend_len;    // != 0: a decoded token that has not been (completely) written: pend_len bytes still to produce ...
    uint32_t pend_dist;   // ... by copying from pend_dist back (a match), or, with pend_dist == 0, the literal byte pend_lit
    uint32_t pend_lit;
    uint8_t lens[320];
};

struct InflateParams {
    const uint8_t *in;
    const uint64_t *in_off;   // n+1
    uint8_t *out;
    const uint64_t *out_off;  // n+1 (slot i capacity = out_off[i+1]-out_off[i])
    uint64_t *out_lens;       // n
    int32_t *statuses;        // n
    uint64_t *in_consumed;    // n or null
    uint32_t *checks;         // 2n {adler32, crc32} or null
    unsigned long long *counter;  // work counter, zero before launch
    const CrcTables *crc;     // needed when a CRC is computed
    const uint32_t *ids;      // optional: the n units to process are ids[0..n) (indices into the offset arrays); null = 0..n-1
    uint32_t n;
    int32_t window_bits;      // -15 raw, 15 zlib, 31 gzip, 47 auto
    int32_t segment_mode;     // 1: raw full-flush segments: end of input at a block boundary is success
    int32_t check_kind;       // segment_mode only: bit0 adler, bit1 crc into `checks`
    int32_t count_only;       // inflate_kernel only: produce no output bytes, just sizes / statuses / consumed (out may be null)
    int32_t serial_only;      // inflate_kernel<1, W> only: 1 = no speculative decode by the idle lanes (experiments)
    ResumeState *resume;      // inflate_kernel only, optional: per-unit resume state (in/out), see ResumeState
};

// litlen table entry (u16): bits 0-3 code length, bits 4-15 payload
//   payload < 0x100          literal byte
//   payload & 0x800          length: bits 0-7 base-3, bits 8-10 extra-bit count
//   payload == 0x100 EOB, 0x200 long code (> CZK_LIT_BITS bits), 0x300 invalid
// dist table entry (u16): bits 0-3 code length, 4-7 extra-bit count, 8-9 mantissa m (dist = (m<<e)+1+extra),
//   bit 14 long code, bit 15 invalid
#define CZK_L_EOB (0x100u << 4)
#define CZK_L_LONG (0x200u << 4)
#define CZK_L_INVALID (0x300u << 4)
#define CZK_D_LONG 0x4000u
#define CZK_D_INVALID 0x8000u

struct SlotSmem {
    uint16_t lit_tab[1 << CZK_LIT_BITS];
    uint16_t dist_tab[1 << CZK_DIST_BITS];
    uint16_t lit_sorted[288];
    uint16_t lit_first[16], lit_offs[16], lit_count[16];
    uint16_t dist_first[16], dist_offs[16], dist_count[16];
    uint8_t dist_sorted[32];
    struct {
        uint8_t lens[320];            // code lengths of the current block (kept: a resumable unit hands them to its next launch)
        uint32_t tokens[CZK_TOKENS];  // token queue while the block is decoded
    } u;
    uint32_t cnt[16];
    uint32_t run[16];
};

__host__ __device__ inline uint32_t lit_entry(uint32_t sym, uint32_t len) {
    if (sym < 256) return (sym << 4) | len;
    if (sym == 256) return CZK_L_EOB | len;
    if (sym > 285) return CZK_L_INVALID | len;
    uint32_t c = sym - 257, e, base;
    if (c < 8) { e = 0; base = 3 + c; }
    else if (c == 28) { e = 0; base = 258; }
    else { e = (c >> 2) - 1; base = 3 + ((4 + (c & 3)) << e); }
    return ((0x800u | (e << 8) | (base - 3)) << 4) | len;
}
__host__ __device__ inline uint32_t dist_entry(uint32_t sym, uint32_t len) {
    if (sym > 29) return CZK_D_INVALID | len;
    uint32_t e = sym < 2 ? 0 : (sym >> 1) - 1;
    uint32_t m = sym < 2 ? sym : 2 + (sym & 1);
    return len | (e << 4) | (m << 8);
}

// Lane-local sequential bit reader over one compressed unit (LSB-first, RFC 1951 §3.1.1).
// The position is not tracked per symbol: bits consumed = 32*widx - 8*mis - cnt, computed on demand. The next input
// word is always prefetched into `nextw`, so the refill on the decode critical path is two ALU ops, not a load.
struct BitReader {
    const uint32_t *words;  // 4-byte aligned base (<= first byte)
    uint32_t mis;           // first byte = (uint8_t*)words + mis
    uint32_t widx, wend;    // words merged into buf so far / number of words that hold stream bytes
    uint32_t tail_mask;     // the bytes of the last word that belong to the unit
    uint32_t cnt;           // valid bits in buf
    uint32_t nextw;         // words[widx], already loaded (0 past the end)
    uint32_t nextw2;        // words[widx + 1], already loaded: a refill consumes a word that was requested two refills ago, so
                            // two refills in a row (length code + distance code) do not wait for the memory round trip
    uint64_t buf;
    uint64_t total;         // bits in the unit

    // Bits past the end of the unit read as ZERO whatever follows it in the packed batch: the last word is masked (a
    // truncated stream must behave the same alone and between neighbours: its tail decides NEED_INPUT vs data error).
    // (the mask is applied when a word enters the bit buffer, not when it is loaded: the look-ahead words must not be
    // touched before they are needed, or every refill waits for the load it has just issued)
    __device__ __forceinline__ uint32_t load(uint32_t i) const { return i < wend ? __ldg(words + i) : 0u; }
    __device__ __forceinline__ uint32_t masked(uint32_t w, uint32_t i) const { return i + 1 == wend ? (w & tail_mask) : w; }
    __device__ __forceinline__ void seek(uint64_t byte_pos) {
        uint64_t a = (uint64_t)mis + byte_pos;
        uint32_t wi = (uint32_t)(a >> 2);
        uint32_t sh = (uint32_t)(a & 3) * 8;
        buf = (uint64_t)(masked(load(wi), wi) >> sh);
        cnt = 32 - sh;
        widx = wi + 1;
        nextw = load(widx);
        nextw2 = load(widx + 1);
    }
    __device__ __forceinline__ void init(const uint8_t *p, uint64_t len) {
        mis = (uint32_t)((uintptr_t)p & 3);
        words = (const uint32_t *)(p - mis);
        wend = (uint32_t)(((uint64_t)mis + len + 3) >> 2);
        const uint32_t tb = (uint32_t)(((uint64_t)mis + len) & 3);
        tail_mask = tb ? (1u << (8 * tb)) - 1u : 0xffffffffu;
        total = len * 8;
        seek(0);
    }
    // after refill(): cnt >= 33
    __device__ __forceinline__ void refill() {
        if (cnt <= 32) {
            buf |= (uint64_t)masked(nextw, widx) << cnt;
            cnt += 32;
            widx++;
            nextw = nextw2;
            nextw2 = load(widx + 1);
        }
    }
    __device__ __forceinline__ uint32_t peek(uint32_t n) const { return (uint32_t)buf & ((1u << n) - 1u); }
    __device__ __forceinline__ void skip(uint32_t n) {
        buf >>= n;
        cnt -= n;
    }
    __device__ __forceinline__ uint32_t get(uint32_t n) {  // n <= 16, caller guarantees cnt >= n
        uint32_t v = peek(n);
        skip(n);
        return v;
    }
    __device__ __forceinline__ uint64_t consumed() const { return (uint64_t)widx * 32 - 8 * mis - cnt; }
    // true when bits beyond the end of the unit have been consumed; cheap unless the reader is in the last words
    __device__ __forceinline__ bool overrun() const { return widx >= wend && consumed() > total; }
    // Status when a decoded symbol does not fit the output slot. zlib stops at that symbol having pulled exactly the input
    // bytes its bits needed, and compu's glue reports Z_OK with avail_in == 0 as NeedInput, avail_in > 0 as NeedOutput
    // (/root/reference/src/decoder/mod.rs:459-486): the symbol's last bit in the last input byte => NeedInput.
    __device__ __forceinline__ int out_full_status() const {
        return ((consumed() + 7) >> 3) >= (total >> 3) ? (int)ST_NEED_INPUT : (int)ST_NEED_OUTPUT;
    }
    __device__ __forceinline__ uint32_t get_byte() {  // byte-wise header parsing
        refill();
        return get(8);
    }
};

enum SlotState : int {
    SS_IDLE = 0,
    SS_HEADER,
    SS_BLOCK,
    SS_BUILD,
    SS_DECODE,
    SS_STORED,
    SS_TRAILER,
    SS_FINISH,
    SS_EXIT
};

// ---------------------------------------------------------------------------------------------------------------
// Warp-cooperative construction of the decode tables of one slot from code lengths in sm.u.lens
// (literal/length lengths at [0, nlit), distance lengths at [nlit, nlit+ndist)). Returns 0 or ST_E_DATA.
// Completeness rules follow zlib's inflate_table(): over-subscribed sets are errors; incomplete sets are errors
// unless the longest code is 1 bit (or the set is empty).
__device__ inline int build_one_table(SlotSmem &sm, const uint8_t *lens, uint32_t n, bool is_dist, uint32_t lane) {
    uint16_t *tab = is_dist ? sm.dist_tab : sm.lit_tab;
    const uint32_t tbits = is_dist ? CZK_DIST_BITS : CZK_LIT_BITS;
    uint16_t *first = is_dist ? sm.dist_first : sm.lit_first;
    uint16_t *offs = is_dist ? sm.dist_offs : sm.lit_offs;
    uint16_t *count = is_dist ? sm.dist_count : sm.lit_count;

    if (lane < 16) { sm.cnt[lane] = 0; sm.run[lane] = 0; }
    __syncwarp();
    for (uint32_t i = lane; i < n; i += 32) {
        uint32_t l = lens[i];
        if (l) atomicAdd(&sm.cnt[l], 1u);
    }
    __syncwarp();
    // every lane derives the canonical-code parameters (cheap, avoids a broadcast)
    int left = 1;
    uint32_t code = 0, off = 0, maxlen = 0;
    uint32_t my_first = 0, my_off = 0, my_cnt = 0;
    bool over = false;
    for (uint32_t len = 1; len <= 15; len++) {
        uint32_t c = sm.cnt[len];
        left = (left << 1) - (int)c;
        if (left < 0) over = true;
        if (lane == len) { my_first = code; my_off = off; my_cnt = c; }
        code = (code + c) << 1;
        off += c;
        if (c) maxlen = len;
    }
    if (over) return ST_E_DATA;
    if (left > 0 && maxlen > 1) return ST_E_DATA;  // incomplete set (zlib: allowed only when max == 1 or empty)
    if (lane >= 1 && lane < 16) { first[lane] = (uint16_t)my_first; offs[lane] = (uint16_t)my_off; count[lane] = (uint16_t)my_cnt; }
    const uint32_t inval = is_dist ? CZK_D_INVALID : CZK_L_INVALID;
    for (uint32_t i = lane; i < (1u << tbits); i += 32) tab[i] = (uint16_t)inval;
    __syncwarp();
    for (uint32_t r = 0; r < n; r += 32) {
        uint32_t sym = r + lane;
        uint32_t l = sym < n ? lens[sym] : 0;
        uint32_t m = __match_any_sync(CZK_FULL, l);
        uint32_t rank = __popc(m & ((1u << lane) - 1u));
        uint32_t base = sm.run[l & 15];
        __syncwarp();
        if (l && (m >> lane) <= 1u) sm.run[l] = base + __popc(m);  // highest lane of the group
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
    return 0;
}

// Slow path for codes longer than the primary table: canonical search over lengths tbits+1..15.
// Returns the table entry (same format as the primary table) or the invalid marker.
__device__ __forceinline__ uint32_t decode_long_lit(const SlotSmem &sm, uint32_t bits15) {
    uint32_t code15 = __brev(bits15) >> 17;
    for (uint32_t len = CZK_LIT_BITS + 1; len <= 15; len++) {
        uint32_t d = (code15 >> (15 - len)) - sm.lit_first[len];
        if (d < sm.lit_count[len]) return lit_entry(sm.lit_sorted[sm.lit_offs[len] + d], len);
    }
    return CZK_L_INVALID;
}
__device__ __forceinline__ uint32_t decode_long_dist(const SlotSmem &sm, uint32_t bits15) {
    uint32_t code15 = __brev(bits15) >> 17;
    for (uint32_t len = CZK_DIST_BITS + 1; len <= 15; len++) {
        uint32_t d = (code15 >> (15 - len)) - sm.dist_first[len];
        if (d < sm.dist_count[len]) return dist_entry(sm.dist_sorted[sm.dist_offs[len] + d], len);
    }
    return CZK_D_INVALID;
}

// ---------------------------------------------------------------------------------------------------------------
// Lane-local: parse a dynamic block header (RFC 1951 §3.2.7) into sm.u.lens. Returns 0, ST_E_DATA, or
// ST_NEED_INPUT+100 (= truncated input) encoded as 100.
__device__ inline int parse_dynamic_header(BitReader &br, SlotSmem &sm, uint32_t &nlit, uint32_t &ndist) {
    br.refill();
    nlit = br.get(5) + 257;
    ndist = br.get(5) + 1;
    uint32_t ncl = br.get(4) + 4;
    if (nlit > 286 || ndist > 30) return ST_E_DATA;  // "too many length or distance symbols"
    // code-length code lengths, permuted order 16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15 packed 5 bits each
    const uint64_t order_lo = 16ull | (17ull << 5) | (18ull << 10) | (0ull << 15) | (8ull << 20) | (7ull << 25) |
                              (9ull << 30) | (6ull << 35) | (10ull << 40) | (5ull << 45) | (11ull << 50) | (4ull << 55);
    const uint64_t order_hi = 12ull | (3ull << 5) | (13ull << 10) | (2ull << 15) | (14ull << 20) | (1ull << 25) | (15ull << 30);
    uint64_t cl_lens = 0;  // 19 x 3 bits, indexed by symbol
    for (uint32_t i = 0; i < ncl; i++) {
        br.refill();
        uint32_t l = br.get(3);
        uint32_t sym = i < 12 ? (uint32_t)(order_lo >> (5 * i)) & 31 : (uint32_t)(order_hi >> (5 * (i - 12))) & 31;
        cl_lens |= (uint64_t)l << (3 * sym);
    }
    if (br.overrun()) return 100;
    // canonical code for the 19-symbol alphabet, <= 7 bits: build a 128-entry table in the (currently unused)
    // head of lit_tab: entry = (sym << 3) | len, 0 = invalid
    uint32_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t s = 0; s < 19; s++) cnt[(cl_lens >> (3 * s)) & 7]++;
    int left = 1;
    uint32_t next[8], code = 0, maxlen = 0;
    cnt[0] = 0;
    for (uint32_t len = 1; len <= 7; len++) {
        left = (left << 1) - (int)cnt[len];
        if (left < 0) return ST_E_DATA;
        next[len] = code;
        code = (code + cnt[len]) << 1;
        if (cnt[len]) maxlen = len;
    }
    if (left > 0) return ST_E_DATA;  // zlib: incomplete code-length set is always an error (type CODES)
    (void)maxlen;
    uint16_t *cl_tab = sm.lit_tab;
    for (uint32_t i = 0; i < 128; i++) cl_tab[i] = 0;
    for (uint32_t s = 0; s < 19; s++) {
        uint32_t l = (uint32_t)(cl_lens >> (3 * s)) & 7;
        if (!l) continue;
        uint32_t c = next[l]++;
        uint32_t rev = __brev(c) >> (32 - l);
        for (uint32_t idx = rev; idx < 128; idx += (1u << l)) cl_tab[idx] = (uint16_t)((s << 3) | l);
    }
    uint32_t total = nlit + ndist, i = 0, prev = 0;
    uint8_t *lens = sm.u.lens;
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
            } else if (s == 17) {
                rep = 3 + br.get(3);
                val = 0;
            } else {
                rep = 11 + br.get(7);
                val = 0;
            }
            if (i + rep > total) return ST_E_DATA;
            for (uint32_t k = 0; k < rep; k++) lens[i++] = (uint8_t)val;
            prev = val;
        }
        if (br.overrun()) return 100;
    }
    if (lens[256] == 0) return ST_E_DATA;  // "invalid code -- missing end-of-block"
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Lane-local: container headers. Return 0 ok, 100 truncated, or a status code (<0 / ST_NEED_DICT).
__device__ inline int parse_zlib_header(BitReader &br) {
    if (br.total < 16) return 100;
    uint32_t cmf = br.get_byte(), flg = br.get_byte();
    if (((cmf << 8) | flg) % 31) return ST_E_DATA;  // "incorrect header check"
    if ((cmf & 15) != 8) return ST_E_DATA;          // "unknown compression method"
    if ((cmf >> 4) > 7) return ST_E_DATA;           // "invalid window size"
    if (flg & 0x20) return ST_NEED_DICT;
    return 0;
}

__device__ inline uint32_t crc32_bitwise(uint32_t crc, uint32_t byte) {
    crc ^= byte;
    for (int k = 0; k < 8; k++) crc = (crc >> 1) ^ (CZK_CRC_POLY & (0u - (crc & 1u)));
    return crc;
}

__device__ inline int parse_gzip_header(BitReader &br) {
    // zlib's order with a partial header (inflate.c HEAD / FLAGS / TIME / OS): the magic is judged as soon as two bytes are
    // there, method and flags with four, and only then are the remaining six bytes needed
    uint32_t hcrc = 0xffffffffu;
    uint32_t h[10];
    if (br.total < 16) return 100;
    for (int i = 0; i < 2; i++) { h[i] = br.get_byte(); hcrc = crc32_bitwise(hcrc, h[i]); }
    if (h[0] != 0x1f || h[1] != 0x8b) return ST_E_DATA;  // "incorrect header check"
    if (br.total < 32) return 100;
    for (int i = 2; i < 4; i++) { h[i] = br.get_byte(); hcrc = crc32_bitwise(hcrc, h[i]); }
    if (h[2] != 8) return ST_E_DATA;                     // "unknown compression method"
    uint32_t flg = h[3];
    if (flg & 0xe0) return ST_E_DATA;  // "unknown header flags set"
    if (br.total < 80) return 100;
    for (int i = 4; i < 10; i++) { h[i] = br.get_byte(); hcrc = crc32_bitwise(hcrc, h[i]); }
    if (flg & 4) {                     // FEXTRA
        uint32_t a = br.get_byte(), b = br.get_byte();
        hcrc = crc32_bitwise(crc32_bitwise(hcrc, a), b);
        uint32_t xlen = a | (b << 8);
        for (uint32_t i = 0; i < xlen; i++) {
            hcrc = crc32_bitwise(hcrc, br.get_byte());
            if (br.overrun()) return 100;
        }
    }
    for (int f = 0; f < 2; f++) {  // FNAME (8), FCOMMENT (16): zero-terminated
        if (flg & (8u << f)) {
            for (;;) {
                uint32_t c = br.get_byte();
                if (br.overrun()) return 100;
                hcrc = crc32_bitwise(hcrc, c);
                if (!c) break;
            }
        }
    }
    if (flg & 2) {  // FHCRC
        uint32_t a = br.get_byte(), b = br.get_byte();
        if (br.overrun()) return 100;
        if ((a | (b << 8)) != ((hcrc ^ 0xffffffffu) & 0xffff)) return ST_E_DATA;  // "header crc mismatch"
    }
    if (br.overrun()) return 100;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-cooperative checksum of out[from, to). adler/crc are running (finalised-form) values, uniform across the warp.
// sum of the four bytes of w, and sum of j * byte j (j = 0..3)
__device__ __forceinline__ void adler_word(uint32_t w, uint32_t &s, uint32_t &ws) {
#if defined(__CUDA_ARCH__)
    s = __dp4a(w, 0x01010101u, 0u);
    ws = __dp4a(w, 0x03020100u, 0u);
#else
    const uint32_t b0 = w & 0xff, b1 = (w >> 8) & 0xff, b2 = (w >> 16) & 0xff, b3 = w >> 24;
    s = b0 + b1 + b2 + b3;
    ws = b1 + 2 * b2 + 3 * b3;
#endif
}

// Adler-32 of n <= 8192 bytes on top of `adler`, by the whole warp: a1 = sum b[k], a2 = sum (n - k) b[k]. The 16-byte aligned
// body is read as uint4 (16 bytes per lane and load, byte sums on the DP4A pipe), the unaligned head and the tail byte-wise.
__device__ inline uint32_t warp_adler32(uint32_t adler, const uint8_t *p, uint32_t n, uint32_t lane) {
    uint32_t a1 = 0, a2 = 0;
    uint32_t head = (uint32_t)((0 - (uintptr_t)p) & 15);
    if (head > n) head = n;
    const uint32_t body = (n - head) & ~15u;
    if (lane < head) { const uint32_t b = p[lane]; a1 = b; a2 = (n - lane) * b; }
    for (uint32_t k = head + 16 * lane; k < head + body; k += 512) {
        const uint4 v = *reinterpret_cast<const uint4 *>(p + k);
        uint32_t s0, s1, s2, s3, w0, w1, w2, w3;
        adler_word(v.x, s0, w0); adler_word(v.y, s1, w1); adler_word(v.z, s2, w2); adler_word(v.w, s3, w3);
        const uint32_t S = s0 + s1 + s2 + s3;
        a1 += S;
        // sum over the 16 bytes of (n - k - j) b[j] = (n - k) S - sum j b[j]
        a2 += (n - k) * S - (w0 + w1 + w2 + w3 + 4 * s1 + 8 * s2 + 12 * s3);
    }
    {
        const uint32_t k = head + body + lane;
        if (k < n) { const uint32_t b = p[k]; a1 += b; a2 += (n - k) * b; }  // (fewer than 16 bytes are left)
    }
    a2 %= CZK_ADLER_BASE;
    a1 = __reduce_add_sync(CZK_FULL, a1);
    a2 = __reduce_add_sync(CZK_FULL, a2);
    uint32_t s1 = adler & 0xffff, s2 = adler >> 16;
    s2 = (s2 + (n % CZK_ADLER_BASE) * s1 + a2) % CZK_ADLER_BASE;
    s1 = (s1 + a1) % CZK_ADLER_BASE;
    return (s2 << 16) | s1;
}

// n = 128*q (q <= 32) bytes: lane l < q computes the CRC of piece l, pieces are folded with x^(1024 k) multipliers.
__device__ inline uint32_t warp_crc32_pieces(uint32_t crc, const uint8_t *p, uint32_t q, const uint32_t *tab,
                                             const uint32_t *pow128, uint32_t lane) {
    uint32_t c = 0;
    if (lane < q) {
        const uint8_t *s = p + 128 * lane;
        uint32_t r = 0xffffffffu;
        for (int i = 0; i < 128; i++) r = (r >> 8) ^ tab[(r ^ s[i]) & 0xff];
        c = crc_mulmod(pow128[q - 1 - lane], r ^ 0xffffffffu);
    }
    // fold the running value; done by an otherwise idle lane when there is one
    const uint32_t folder = q < 32 ? 31u : 0u;
    if (lane == folder) c ^= crc_mulmod(pow128[q], crc);
    return __reduce_xor_sync(CZK_FULL, c);
}

__device__ inline uint32_t crc32_serial(uint32_t crc, const uint8_t *p, uint32_t n, const uint32_t *tab) {
    uint32_t r = crc ^ 0xffffffffu;
    for (uint32_t i = 0; i < n; i++) r = (r >> 8) ^ tab[(r ^ p[i]) & 0xff];
    return r ^ 0xffffffffu;
}

// ---------------------------------------------------------------------------------------------------------------
template <int D, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) inflate_kernel(InflateParams P) {
    CZ_DYNAMIC_SMEM(smem_raw);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // layout: [CrcTables (1 KB + pow128)] [WARPS][D] SlotSmem
    uint32_t *crc_tab = (uint32_t *)smem_raw;
    uint32_t *crc_pow = crc_tab + 256;
    SlotSmem *slots = (SlotSmem *)(smem_raw + 1280) + (size_t)warp * D;
    const bool use_crc = P.crc != nullptr;
    if (use_crc) {
        for (uint32_t i = threadIdx.x; i < 256 + 34; i += WARPS * 32)
            crc_tab[i] = i < 256 ? P.crc->table[i] : P.crc->pow128[i - 256];
    }
    __syncthreads();

    // ---- per-slot registers (meaningful on lanes < D)
    BitReader br;
    br.words = nullptr; br.mis = 0; br.widx = br.wend = 0; br.cnt = 0; br.buf = 0; br.nextw = 0; br.nextw2 = 0; br.total = 0; br.tail_mask = 0xffffffffu;
    int st = lane < D ? SS_IDLE : SS_EXIT;
    uint32_t unit = 0;
    const uint8_t *in_base = nullptr;
    uint64_t in_len = 0;
    uint8_t *out_base = nullptr;
    uint64_t out_pos = 0, out_cap = 0, ck_pos = 0;
    uint32_t adler = 1, crc = 0;
    int result = 0;          // status to report at SS_FINISH
    int wrap = 0;            // 0 raw, 1 zlib, 2 gzip (resolved per stream)
    uint32_t bfinal = 0, ntok = 0, nlit = 0, ndist = 0, stored_len = 0;
    int after_tokens = SS_DECODE;  // state to enter once the token queue has been drained
    int full_status = -1;          // >= 0: the decoder stopped at the first token that does not fit the slot; its status
    // resumable units (P.resume): where and in which phase the next launch restarts, history in front of the slot, bytes of
    // earlier launches, and a decoded token that did not fit (completely) yet
    uint64_t rbits = 0, prior_out = 0;
    uint32_t rphase = RP_HEADER, hist = 0, pend_len = 0, pend_dist = 0, pend_lit = 0;
    const bool resumable = P.resume != nullptr;
    SlotSmem &my = slots[lane < D ? lane : 0];

    for (;;) {
        // ---- (1) fetch work
        if (st == SS_IDLE) {
            unsigned long long u = atomicAdd(P.counter, 1ull);
            if (u >= P.n) st = SS_EXIT;
            else {
                unit = P.ids ? P.ids[u] : (uint32_t)u;
                uint64_t i0 = P.in_off[unit], i1 = P.in_off[unit + 1], o0 = P.out_off[unit], o1 = P.out_off[unit + 1];
                in_base = P.in + i0; in_len = i1 - i0;
                out_base = P.out + o0; out_cap = o1 - o0; out_pos = 0; ck_pos = 0;
                adler = 1; crc = 0; ntok = 0; bfinal = 0; result = 0; full_status = -1;
                br.init(in_base, in_len);
                st = SS_HEADER;
                if (resumable) {
                    const ResumeState &R = P.resume[unit];
                    hist = R.hist_len; prior_out = R.total_out; adler = R.adler; crc = R.crc; bfinal = R.bfinal; wrap = (int)R.wrap;
                    pend_len = R.pend_len; pend_dist = R.pend_dist; pend_lit = R.pend_lit;
                    rbits = R.bit_pos; rphase = R.phase;
                    br.seek(rbits >> 3);
                    br.skip((uint32_t)(rbits & 7));
                    if (rphase == RP_BLOCK) st = SS_BLOCK;
                    else if (rphase == RP_CODES) {
                        nlit = R.nlit; ndist = R.ndist;
                        for (uint32_t i = 0; i < nlit + ndist && i < 320; i++) my.u.lens[i] = R.lens[i];
                        st = SS_BUILD;
                    } else if (rphase == RP_STORED) { stored_len = R.stored_left; st = SS_STORED; }
                    else if (rphase == RP_TRAILER) { result = ST_FINISHED; st = SS_TRAILER; }
                    else if (rphase == RP_DONE) { result = ST_FINISHED; st = SS_FINISH; }
                }
            }
        }
        if (__all_sync(CZK_FULL, st == SS_EXIT)) break;

        // ---- (2) container header
        if (st == SS_HEADER) {
            int r = 0;
            if (P.segment_mode || P.window_bits < 0) wrap = 0;
            else if (P.window_bits == 47) {
                // auto: gzip magic, else zlib
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
                // a full-flush segment ends byte-aligned exactly at the end of its input
                result = br.consumed() == br.total ? ST_FINISHED : ST_NEED_INPUT;
                st = SS_TRAILER;
            } else {
                if (resumable) { rphase = RP_BLOCK; rbits = br.consumed(); }  // out of input inside the header: back to its start
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
                    uint8_t *lens = my.u.lens;
                    for (uint32_t i = 0; i < 288; i++) lens[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
                    for (uint32_t i = 0; i < 32; i++) lens[288 + i] = 5;
                    nlit = 288; ndist = 32;
                    st = SS_BUILD;
                } else if (btype == 2) {
                    int r = parse_dynamic_header(br, my, nlit, ndist);
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
                SlotSmem &sm = slots[s];
                __syncwarp();
                int r = build_one_table(sm, sm.u.lens, nl, false, lane);
                __syncwarp();
                int r2 = build_one_table(sm, sm.u.lens + nl, nd, true, lane);
                __syncwarp();
                if ((int)lane == s) {
                    if (r || r2) { result = ST_E_DATA; st = SS_FINISH; }
                    else { st = SS_DECODE; ntok = 0; rphase = RP_CODES; }
                }
            }
        }

        // ---- (5p) D == 1 only: the warp decodes ONE stream, so the 31 lanes that would idle during (5) decode speculatively.
        // Lane l assumes a literal/length code starts at bit (position + l) and (position + 32 + l) and decodes the whole token
        // there (code, extra bits, distance code, extra bits: two table look-ups); then the true chain of token starts is
        // followed through those 64 results with one shuffle per token. A serial lane needs ~250 cycles per token (dependent
        // look-ups); a round of this costs about as much and yields ~5 tokens on text. Anything unusual — a code longer than the
        // primary tables, an invalid code, the last bytes of the input — stops the round and is left to the serial decoder (5).
        // MEASURED (16 zlib-made streams of 8 MiB, tools/big_stream_probe.py): 596 ms against 607 ms serial. A lone warp is bound
        // by the ~5-cycle dependent issue distance, and a round (two look-ups for 64 offsets, ~5 shuffle hops of the chase,
        // compaction) is ~35 dependent instructions per token against ~50 for the serial lane. OFF by default (CZ_PAR_DECODE=1).
        bool par_eob = false;
#ifdef CZ_EXPERIMENTS
        if constexpr (D == 1) {
            const bool can = !P.serial_only && !resumable && __shfl_sync(CZK_FULL, (int)(st == SS_DECODE), 0) != 0;
            if (can) {
                const uint8_t *ib0 = (const uint8_t *)(uintptr_t)__shfl_sync(CZK_FULL, (unsigned long long)(uintptr_t)in_base, 0);
                uint64_t cpos = __shfl_sync(CZK_FULL, (unsigned long long)br.consumed(), 0);
                const uint64_t tbits = __shfl_sync(CZK_FULL, (unsigned long long)br.total, 0);
                uint32_t nq = __shfl_sync(CZK_FULL, ntok, 0);
                const uint16_t *lt = my.lit_tab, *dt = my.dist_tab;
                uint32_t *tokq = my.u.tokens;
                int stop = 0;  // 1: end of block consumed, 2: next token needs the serial decoder
                bool did = false;
                while (!stop && nq < CZK_TOKENS && cpos + 240 <= tbits) {
                    uint32_t meta[2], tokv[2];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const uint64_t b = cpos + 32u * h + lane;
                        const uintptr_t a = (uintptr_t)(ib0 + (b >> 3));
                        const uint32_t *wp = (const uint32_t *)(a & ~(uintptr_t)3);
                        const uint32_t sh = (uint32_t)(a & 3) * 8 + (uint32_t)(b & 7);  // 0..31
                        const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);
                        const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
                        uint64_t w = (uint64_t)lo | ((uint64_t)hi << 32);  // the 64 bits that start at bit b
                        uint32_t e = lt[(uint32_t)w & ((1u << CZK_LIT_BITS) - 1u)];
                        if ((e & 0xfff0u) == CZK_L_LONG) e = decode_long_lit(my, (uint32_t)w & 0x7fffu);
                        uint32_t pay = e >> 4, tl = e & 15, kind = 0, tv = 0;
                        if ((e & 0xfff0u) == CZK_L_INVALID) kind = 2;
                        else if (pay < 0x100) tv = 0x80000000u | (pay << 9) | 1u;
                        else if (!(pay & 0x800)) kind = pay == 0x100 ? 1 : 2;
                        else {
                            w >>= tl;
                            const uint32_t eb = (pay >> 8) & 7;
                            const uint32_t len = 3 + (pay & 0xff) + ((uint32_t)w & ((1u << eb) - 1u));
                            w >>= eb;
                            uint32_t de = dt[(uint32_t)w & ((1u << CZK_DIST_BITS) - 1u)];
                            if (de & CZK_D_LONG) de = decode_long_dist(my, (uint32_t)w & 0x7fffu);
                            if (de & CZK_D_INVALID) kind = 2;
                            else {
                                const uint32_t dl = de & 15, deb = (de >> 4) & 15;
                                w >>= dl;
                                const uint32_t dist = (((de >> 8) & 3) << deb) + 1 + ((uint32_t)w & ((1u << deb) - 1u));
                                tl += eb + dl + deb;
                                tv = (dist << 9) | len;
                            }
                        }
                        meta[h] = tl | (kind << 6);
                        tokv[h] = tv;
                    }
                    // follow the chain of real token starts
                    uint32_t o = 0, on0 = 0, on1 = 0, cnt = 0;
                    while (o < 64 && nq + cnt < CZK_TOKENS) {
                        const uint32_t mm = __shfl_sync(CZK_FULL, o < 32 ? meta[0] : meta[1], o & 31);
                        const uint32_t kind = mm >> 6;
                        if (kind == 2) { stop = 2; break; }
                        if (kind == 1) { o += mm & 63; stop = 1; break; }
                        if (o < 32) on0 |= 1u << o; else on1 |= 1u << (o - 32);
                        cnt++;
                        o += mm & 63;
                    }
                    const uint32_t lt_mask = (1u << lane) - 1u;
                    if ((on0 >> lane) & 1u) tokq[nq + __popc(on0 & lt_mask)] = tokv[0];
                    if ((on1 >> lane) & 1u) tokq[nq + __popc(on0) + __popc(on1 & lt_mask)] = tokv[1];
                    nq += cnt;
                    cpos += o;
                    did = did || o != 0;
                }
                __syncwarp();
                if (did && lane == 0) {
                    ntok = nq;
                    br.seek(cpos >> 3);
                    br.skip((uint32_t)(cpos & 7));
                    if (stop == 1) {
                        par_eob = true;
                        after_tokens = bfinal ? SS_TRAILER : SS_BLOCK;
                        if (bfinal) result = ST_FINISHED;
                        if (!ntok) st = after_tokens;
                    }
                }
            }
        }

#endif  // CZ_EXPERIMENTS
        // ---- (5) Huffman decode: D lanes, each its own stream, up to CZK_TOKENS tokens
        if (st == SS_DECODE && !par_eob) {
            uint32_t *tok = my.u.tokens;
            after_tokens = SS_DECODE;
            // zlib stops at the first symbol that does not fit the output slot, with exactly that symbol's bits consumed;
            // the queue must not run ahead of it (the status and the consumed-byte count depend on where it stops)
            uint64_t room = out_cap - out_pos;
            const bool track = ntok == 0;  // (tokens queued by the speculative path (5p) are not accounted for)
            bool stop = false;
            if (resumable && pend_len) {  // the token the previous launch could not finish comes first
                const uint32_t n = pend_len < room ? pend_len : (uint32_t)room;
                if (n) { tok[ntok++] = pend_dist ? ((pend_dist << 9) | n) : (0x80000000u | (pend_lit << 9) | 1u); room -= n; pend_len -= n; }
                if (pend_len) { result = ST_NEED_OUTPUT; after_tokens = SS_FINISH; stop = true; }
            }
            while (!stop && ntok < CZK_TOKENS) {
                if (resumable) rbits = br.consumed();  // out of input inside this symbol: the next launch re-reads it
                br.refill();
                uint32_t e = my.lit_tab[br.peek(CZK_LIT_BITS)];
                if ((e & 0xfff0u) == CZK_L_LONG) e = decode_long_lit(my, br.peek(15));
                uint32_t pay = e >> 4;
                if (pay < 0x100) {  // literal
                    br.skip(e & 15);
                    if (br.overrun()) { result = ST_NEED_INPUT; after_tokens = SS_FINISH; break; }
                    if (resumable) {
                        // zlib has taken the symbol's bits when it finds no room for its byte (inflate.c LIT): so does the state
                        if (room == 0) { pend_len = 1; pend_dist = 0; pend_lit = pay; rbits = br.consumed(); result = ST_NEED_OUTPUT; after_tokens = SS_FINISH; break; }
                        room--;
                        tok[ntok++] = 0x80000000u | (pay << 9) | 1u;
                        continue;
                    }
                    tok[ntok++] = 0x80000000u | (pay << 9) | 1u;
                    if (track) {
                        if (room == 0) { full_status = br.out_full_status(); after_tokens = SS_FINISH; break; }
                        room--;
                    }
                    continue;
                }
                if (!(pay & 0x800)) {
                    if (pay == 0x100) {  // end of block
                        br.skip(e & 15);
                        if (br.overrun()) { result = ST_NEED_INPUT; after_tokens = SS_FINISH; break; }
                        after_tokens = bfinal ? SS_TRAILER : SS_BLOCK;
                        if (bfinal) result = ST_FINISHED;
                        break;
                    }
                    // invalid code; zero bits past a truncated input can land here too
                    if (br.consumed() + ((e & 15) ? (e & 15) : 1) > br.total) result = ST_NEED_INPUT; else result = ST_E_DATA;
                    after_tokens = SS_FINISH;
                    break;
                }
                // length
                br.skip(e & 15);
                uint32_t eb = (pay >> 8) & 7;
                uint32_t len = 3 + (pay & 0xff) + br.peek(eb);
                br.skip(eb);
                br.refill();
                uint32_t de = my.dist_tab[br.peek(CZK_DIST_BITS)];
                if (de & CZK_D_LONG) de = decode_long_dist(my, br.peek(15));
                if (de & CZK_D_INVALID) {
                    if (br.consumed() + ((de & 15) ? (de & 15) : 1) > br.total) result = ST_NEED_INPUT; else result = ST_E_DATA;
                    after_tokens = SS_FINISH;
                    break;
                }
                br.skip(de & 15);
                uint32_t deb = (de >> 4) & 15;
                uint32_t dist = (((de >> 8) & 3) << deb) + 1 + br.peek(deb);
                br.skip(deb);
                if (br.overrun()) { result = ST_NEED_INPUT; after_tokens = SS_FINISH; break; }
                if (resumable) {
                    // the part of the match that fits is written, the rest waits in the state (inflate.c MATCH: state->length)
                    const uint32_t n = len < room ? len : (uint32_t)room;
                    if (n) tok[ntok++] = (dist << 9) | n;
                    room -= n;
                    if (n < len) { pend_len = len - n; pend_dist = dist; rbits = br.consumed(); result = ST_NEED_OUTPUT; after_tokens = SS_FINISH; break; }
                    continue;
                }
                tok[ntok++] = (dist << 9) | len;
                if (track) {
                    if (len > room) { full_status = br.out_full_status(); after_tokens = SS_FINISH; break; }
                    room -= len;
                }
            }
            st = after_tokens == SS_DECODE ? SS_DECODE : (ntok ? SS_DECODE : after_tokens);
        }

        // ---- (6) LZ77 resolution of each slot's token queue, all 32 lanes
        {
            uint32_t m = __ballot_sync(CZK_FULL, ntok > 0);
            while (m) {
                int s = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t nt = __shfl_sync(CZK_FULL, ntok, s);
                uint8_t *ob = (uint8_t *)(uintptr_t)__shfl_sync(CZK_FULL, (unsigned long long)(uintptr_t)out_base, s);
                const uint64_t opos = __shfl_sync(CZK_FULL, (unsigned long long)out_pos, s);
                const uint64_t ocap = __shfl_sync(CZK_FULL, (unsigned long long)out_cap, s);
                const uint32_t hist_s = __shfl_sync(CZK_FULL, hist, s);
                __syncwarp();
                uint32_t t = lane < nt ? slots[s].u.tokens[lane] : 0;
                uint32_t tl = t & 0x1ff;
                // exclusive scan of lengths
                uint32_t pos = tl;
                for (int d = 1; d < 32; d <<= 1) {
                    uint32_t v = __shfl_up_sync(CZK_FULL, pos, d);
                    if ((int)lane >= d) pos += v;
                }
                uint32_t total = __shfl_sync(CZK_FULL, pos, 31);
                pos -= tl;
                // distance validity: dist <= bytes produced before the token ("invalid distance too far back")
                bool bad = lane < nt && !(t >> 31) && (uint64_t)(t >> 9) > opos + pos + hist_s;
                uint32_t badm = __ballot_sync(CZK_FULL, bad);
                int err = 0;
                if (badm) {
                    int fb = __ffs(badm) - 1;
                    total = __shfl_sync(CZK_FULL, pos, fb);  // keep everything before the bad token
                    // zlib looks at the output space before the distance (inflate.c MATCH: `if (left == 0) goto inf_leave`)
                    err = opos + total >= ocap ? ST_NEED_OUTPUT : ST_E_DATA;
                }
                if (opos + total > ocap) {  // output slot full: deliver the prefix that fits
                    total = (uint32_t)(ocap - opos);
                    err = ST_NEED_OUTPUT;
                }
                const uint64_t a0 = (uint64_t)(uintptr_t)ob + opos;   // absolute address of batch byte 0
                int rel = -(int)(a0 & 31);                             // batch-relative index of this round's lane 0
                uint32_t cnt_before = 0;
                uint8_t *obp = ob + opos;
                if (P.count_only) rel = (int)total;  // sizes only: the tokens' lengths are all that is needed
                for (; rel < (int)total; rel += 32) {
                    // which tokens start inside this round?
                    uint32_t bit = (lane < nt && (int)pos >= rel && (int)pos < rel + 32) ? 1u << ((int)pos - rel) : 0u;
                    uint32_t S = __reduce_or_sync(CZK_FULL, bit);
                    int j = rel + (int)lane;
                    bool active = j >= 0 && j < (int)total;
                    int ti = (int)cnt_before + __popc(S & (0xffffffffu >> (31 - lane))) - 1;
                    cnt_before += __popc(S);
                    uint32_t tk = __shfl_sync(CZK_FULL, t, ti & 31);
                    uint32_t tp = __shfl_sync(CZK_FULL, pos, ti & 31);
                    uint32_t val = 0;
                    bool need = false;
                    int srcl = 0;
                    if (active) {
                        if (tk >> 31) val = (tk >> 9) & 0xff;
                        else {
                            uint32_t dist = tk >> 9, off = (uint32_t)j - tp;
                            if (off >= dist) off %= dist;
                            int src = (int)tp - (int)dist + (int)off;  // batch-relative source index (< tp)
                            int lo = rel > 0 ? rel : 0;
                            if (src >= lo) { need = true; srcl = src - rel; }
                            else val = obp[src];                        // bytes of earlier rounds / batches
                        }
                    }
                    uint32_t pend = __ballot_sync(CZK_FULL, need);
                    while (pend) {
                        uint32_t v = __shfl_sync(CZK_FULL, val, srcl);
                        bool src_ready = !((pend >> srcl) & 1u);
                        if (need && src_ready) { val = v; need = false; }
                        pend = __ballot_sync(CZK_FULL, need);
                    }
                    if (active) obp[j] = (uint8_t)val;
                    __syncwarp();
                }
                if ((int)lane == s) {
                    out_pos = opos + total;
                    ntok = 0;
                    if (err) { result = (err == ST_NEED_OUTPUT && full_status >= 0) ? full_status : err; st = SS_FINISH; }
                    else st = after_tokens;
                }
            }
        }

        // ---- (7) stored blocks: warp-cooperative byte copy
        {
            uint32_t m = __ballot_sync(CZK_FULL, st == SS_STORED);
            while (m) {
                int s = __ffs(m) - 1;
                m &= m - 1;
                uint32_t len = __shfl_sync(CZK_FULL, stored_len, s);
                const uint8_t *ib = (const uint8_t *)(uintptr_t)__shfl_sync(CZK_FULL, (unsigned long long)(uintptr_t)in_base, s);
                uint8_t *ob = (uint8_t *)(uintptr_t)__shfl_sync(CZK_FULL, (unsigned long long)(uintptr_t)out_base, s);
                const uint64_t opos = __shfl_sync(CZK_FULL, (unsigned long long)out_pos, s);
                const uint64_t ocap = __shfl_sync(CZK_FULL, (unsigned long long)out_cap, s);
                const uint64_t ipos = __shfl_sync(CZK_FULL, (unsigned long long)br.consumed(), s) >> 3;
                const uint64_t ilen = __shfl_sync(CZK_FULL, (unsigned long long)in_len, s);
                int err = -1;  // none
                uint32_t n = len;
                if (ipos + n > ilen) { n = (uint32_t)(ilen - ipos); err = ST_NEED_INPUT; }
                if (opos + n > ocap) { n = (uint32_t)(ocap - opos); err = ST_NEED_OUTPUT; }
                if (!P.count_only)
                    for (uint32_t k = lane; k < n; k += 32) ob[opos + k] = ib[ipos + k];
                __syncwarp();
                if ((int)lane == s) {
                    out_pos = opos + n;
                    if (err >= 0) {
                        result = err; st = SS_FINISH; br.seek(ipos + n);
                        if (resumable) { rphase = RP_STORED; rbits = (ipos + n) * 8; stored_len = len - n; }
                    }
                    else {
                        br.seek(ipos + n);
                        if (bfinal) { result = ST_FINISHED; st = SS_TRAILER; } else st = SS_BLOCK;
                    }
                }
            }
        }

        // ---- (8) checksums over freshly written output
        {
            const bool fin = st == SS_TRAILER || st == SS_FINISH;
            bool want_adler = false, want_crc = false;
            if (lane < D && st != SS_EXIT && st != SS_IDLE) {
                // raw units asked for checks (pieces of a longer stream): both, like segment mode
                const bool by_kind = P.segment_mode || (P.checks && wrap == 0);
                want_adler = by_kind ? (P.check_kind & 1) : wrap == 1;
                want_crc = by_kind ? (P.check_kind & 2) : wrap == 2;
            }
            uint64_t pending = out_pos - ck_pos;
            bool go = !P.count_only && (want_adler || want_crc) && (fin ? pending > 0 : pending >= 4096);
            uint32_t m = __ballot_sync(CZK_FULL, go);
            while (m) {
                int s = __ffs(m) - 1;
                m &= m - 1;
                const uint8_t *ob = (const uint8_t *)(uintptr_t)__shfl_sync(CZK_FULL, (unsigned long long)(uintptr_t)out_base, s);
                uint64_t from = __shfl_sync(CZK_FULL, (unsigned long long)ck_pos, s);
                uint64_t to = __shfl_sync(CZK_FULL, (unsigned long long)out_pos, s);
                const bool sfin = __shfl_sync(CZK_FULL, (int)fin, s);
                const bool wa = __shfl_sync(CZK_FULL, (int)want_adler, s), wc = __shfl_sync(CZK_FULL, (int)want_crc, s);
                uint32_t a = __shfl_sync(CZK_FULL, adler, s), c = __shfl_sync(CZK_FULL, crc, s);
                if (!sfin) to = from + ((to - from) & ~(uint64_t)127);  // keep CRC pieces at 128 B until the end
                __syncwarp();
                if (wa) {
                    for (uint64_t p = from; p < to; p += 8192) {
                        uint32_t n = (uint32_t)(to - p < 8192 ? to - p : 8192);
                        a = warp_adler32(a, ob + p, n, lane);
                    }
                }
                if (wc) {
                    uint64_t p = from;
                    while (to - p >= 128) {
                        uint32_t q = (uint32_t)((to - p) >> 7);
                        if (q > 32) q = 32;
                        c = warp_crc32_pieces(c, ob + p, q, crc_tab, crc_pow, lane);
                        p += (uint64_t)q * 128;
                    }
                    if (p < to) {
                        uint32_t c2 = 0;
                        if (lane == 0) c2 = crc32_serial(c, ob + p, (uint32_t)(to - p), crc_tab);
                        c = __shfl_sync(CZK_FULL, c2, 0);
                    }
                }
                if ((int)lane == s) { adler = a; crc = c; ck_pos = to; }
            }
        }

        // ---- (9) trailer
        if (st == SS_TRAILER) {
            if (!P.segment_mode && result == ST_FINISHED) {
                // back to a byte boundary, then the container trailer
                br.skip((uint32_t)((0 - br.consumed()) & 7));
                if (resumable) { rphase = RP_TRAILER; rbits = br.consumed(); }
                if (wrap == 1) {
                    uint32_t v = 0;
                    for (int i = 0; i < 4; i++) v = (v << 8) | br.get_byte();
                    if (br.overrun()) result = ST_NEED_INPUT;
                    else if (v != adler) result = ST_E_DATA;  // "incorrect data check"
                } else if (wrap == 2) {
                    // zlib judges the CRC as soon as its four bytes are there (inflate.c CHECK), before it asks for ISIZE
                    uint32_t v = 0, isz = 0;
                    for (int i = 0; i < 4; i++) v |= br.get_byte() << (8 * i);
                    if (br.overrun()) result = ST_NEED_INPUT;
                    else if (v != crc) result = ST_E_DATA;                 // "incorrect data check"
                    else {
                        for (int i = 0; i < 4; i++) isz |= br.get_byte() << (8 * i);
                        if (br.overrun()) result = ST_NEED_INPUT;
                        else if (isz != (uint32_t)(prior_out + out_pos)) result = ST_E_DATA;  // "incorrect length check"
                    }
                }
            }
            st = SS_FINISH;
        }

        // ---- (10) report
        if (st == SS_FINISH) {
            if (resumable) {
                ResumeState &R = P.resume[unit];
                R.bit_pos = result == ST_FINISHED ? br.consumed() : rbits;
                R.phase = result == ST_FINISHED ? (uint32_t)RP_DONE : rphase;
                R.total_out = prior_out + out_pos;
                R.bfinal = bfinal; R.nlit = nlit; R.ndist = ndist; R.stored_left = stored_len; R.wrap = (uint32_t)wrap;
                const uint64_t h = (uint64_t)hist + out_pos;
                R.hist_len = h > 32768 ? 32768u : (uint32_t)h;
                R.adler = adler; R.crc = crc;
                R.pend_len = pend_len; R.pend_dist = pend_dist; R.pend_lit = pend_lit;
                if (rphase == RP_CODES)
                    for (uint32_t i = 0; i < nlit + ndist && i < 320; i++) R.lens[i] = my.u.lens[i];
            }
            P.out_lens[unit] = out_pos;
            P.statuses[unit] = result;
            if (P.in_consumed) {
                uint64_t c = (br.consumed() + 7) >> 3;
                P.in_consumed[unit] = c < in_len ? c : in_len;
            }
            if (P.checks) { P.checks[2 * unit] = adler; P.checks[2 * unit + 1] = crc; }
            st = SS_IDLE;
        }
    }
}

template <int D, int WARPS>
constexpr size_t inflate_smem_bytes() { return 1280 + sizeof(SlotSmem) * (size_t)D * WARPS; }

}  // namespace czk
