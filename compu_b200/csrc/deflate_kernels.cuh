// deflate_kernels.cuh — segmented DEFLATE encoder for sm_100a as a chain of data-parallel passes.
//
// Replaces what compu outsources to L0 `deflate()` behind encode_fn (/root/reference/src/encoder/zlib_ng.rs:90-92 ->
// src/encoder/mod.rs:334-370). The input of every unit (buffer to compress) is cut into segments of at most 1 MiB; each
// segment is compressed with no history from its neighbours and ends with an empty stored block (what Z_FULL_FLUSH
// emits), so the concatenation  header | seg0 | seg1 | ... | 03 00 | trailer  is ONE valid zlib/gzip/raw stream
//  and every segment can later be inflated on its own.
//
// Data layout in HBM (all offsets relative to seg_off[0]; "B" = bytes of input in the launch):
//   in        B bytes        segments are contiguous: segment s = in[seg_off[s] .. seg_off[s+1])
//   prevd     2B bytes       hash-chain links (distance to the previous position with the same 4-byte hash)
//   match     4B bytes       best match per position (len | dist << 9); overwritten in place by the token list
//   blocks    (B >> 14) + nseg + 1 slots: segment s owns slots [blk_first(s), blk_first(s+1)); a block = 16384 tokens
//             per slot: blk_end (input offset where the block ends), freqs (320 counters), BlockPlan (codes + header)
//
// Passes (all intermediates stay in HBM; the match search K2 dominates):
//   K0 checksum   warp per segment      Adler-32 / CRC-32 of the segment's input (combined per unit in K5c)
//   K1 chains     warp per segment      prevd[]: 32 positions per step, __match_any_sync inside the step, a 16 KB
//                                       shared-memory head table across steps (result == sequential insertion)
//   K2 match      thread per position   best (length, distance) by walking the chain (deflate_core.cuh find_match)
//   K3 parse      warp per segment      one-step-lazy parse into tokens, in place (tile staged in shared memory, one
//                                       lane walks the serial chain, all lanes emit); cuts blocks of 16384 tokens
//   K4a histogram CTA per block         literal/length and distance symbol counts (shared-memory atomics)
//   K4b plan      thread per block      length-limited Huffman codes, stored/fixed/dynamic choice, rendered header
//   K5a layout    thread per segment    bit offset of every block, size of the segment
//   K5c sizes     thread per unit       size of the unit's stream, combined checksums, capacity check
//   K5s scan      one CTA               (packed output only) exclusive scan of the unit sizes -> output positions
//   K5d frame     thread per unit       container header, flush markers, final block, trailer; segment positions
//   K5b zero      thread per block      clears the bytes that two blocks share (they are written with atomicOr)
//   K6 emit       CTA per block         token -> code bits, block-wide prefix scan of the bit lengths, bits assembled in
//                                       shared memory and written out as whole bytes
#pragma once
#include "deflate_core.cuh"
#include "inflate_kernel.cuh"  // warp_adler32 / warp_crc32_pieces / crc32_serial

namespace czk {

#define CZK_SEG_MAX (1u << 20)  // largest segment
#define CZK_FREQ_STRIDE 320u    // 288 literal/length slots + 32 distance slots per block

struct SegState {
    uint32_t ntok, nblk;
    uint32_t adler, crc;
    uint64_t out_bytes;  // bytes of the segment including its flush marker
    uint64_t out_off;    // absolute byte offset of the segment in the output buffer (~0: unit not written)
    uint64_t body_bits;  // bits of all blocks (before the flush marker)
};

struct DeflateParams {
    const uint8_t *in;             // in + seg_off[s] is the first byte of segment s
    uint8_t *out;
    const uint64_t *seg_off;       // nseg+1
    uint32_t nseg, n_units, n_slots;
    const uint32_t *unit_seg;      // n_units+1: first segment of each unit; null => unit u is segment u
    const uint64_t *unit_out_off;  // n_units+1: output slots (capacity of unit u = unit_out_off[u+1]-unit_out_off[u])
    uint64_t *unit_out_pos;        // null: unit u is written at its slot; else at out + unit_out_pos[u] (packed, by K5s)
    uint64_t *unit_out_len;        // size of the unit's stream (also when it did not fit)
    int32_t *unit_status;
    uint32_t *unit_checks;         // 2 per unit {adler32, crc32} of the unit's input, or null
    uint64_t *seg_out_bytes;       // per segment compressed size (incl. flush marker) or null — the side index
    uint64_t *total_out;           // packed mode: total bytes written (1 value)
    SegState *st;                  // nseg
    uint16_t *prevd;
    uint16_t *prevd2;              // second link of every position (K1b), or null
    uint32_t *match;
    uint32_t *blk_end;             // n_slots
    uint32_t *freqs;               // n_slots * CZK_FREQ_STRIDE
    BlockPlan *plans;              // n_slots
    const CrcTables *crc;
    DeflateTuning tune;
    int32_t window_bits;  // -15 raw, 15 zlib, 31 gzip
    int32_t level;
    int32_t piece_mode;   // 1: units are raw pieces of a longer stream: no container, no final block
    int32_t check_kind;   // bit0 Adler-32, bit1 CRC-32
};

__device__ __forceinline__ uint64_t seg_base(const DeflateParams &P, uint32_t seg) { return P.seg_off[seg] - P.seg_off[0]; }
__device__ __forceinline__ uint32_t seg_len(const DeflateParams &P, uint32_t seg) { return (uint32_t)(P.seg_off[seg + 1] - P.seg_off[seg]); }
// first block slot of a segment: a closed form that leaves every segment ceil(len / 16384) + 1 > #blocks slots
__device__ __forceinline__ uint32_t blk_first(const DeflateParams &P, uint32_t seg) { return (uint32_t)(seg_base(P, seg) >> 14) + seg; }

// slot -> (segment, block index inside it); false if the slot is unused
__device__ inline bool slot_to_block(const DeflateParams &P, uint32_t slot, uint32_t &seg, uint32_t &b) {
    uint32_t lo = 0, hi = P.nseg;  // last segment with blk_first <= slot
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (blk_first(P, mid) <= slot) lo = mid; else hi = mid;
    }
    seg = lo;
    b = slot - blk_first(P, lo);
    return b < P.st[lo].nblk;
}

__device__ __forceinline__ uint32_t lds32u(const uint8_t *sb, uint32_t off) {  // unaligned little-endian word from shared memory
    const uint32_t *w = (const uint32_t *)sb + (off >> 2);
    return __funnelshift_r(w[0], w[1], (off & 3u) * 8u);
}

// ------------------------------------------------------------------ K0
// One CTA (4 warps) per segment: every warp checksums a quarter, lane 0 folds the four partial values with the combine
// identities (the same ones that fold segments into units in K5c).
__global__ void __launch_bounds__(128) deflate_checksum_kernel(DeflateParams P) {
    __shared__ uint32_t crc_tab[256 + 34];
    __shared__ uint32_t part_a[4], part_c[4], part_n[4];
    for (uint32_t i = threadIdx.x; i < 256 + 34; i += 128) crc_tab[i] = i < 256 ? P.crc->table[i] : P.crc->pow128[i - 256];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t seg = blockIdx.x; seg < P.nseg; seg += gridDim.x) {
        const uint32_t n_all = seg_len(P, seg);
        const uint32_t q = ((n_all + 3) / 4 + 127) & ~127u;  // quarter, a multiple of the 128-byte CRC piece
        const uint32_t b0 = warp * q < n_all ? warp * q : n_all;
        const uint32_t n = n_all - b0 < q ? n_all - b0 : q;
        const uint8_t *p = P.in + P.seg_off[seg] + b0;
        uint32_t a = 1, c = 0;
        if (P.check_kind & 1) {
            for (uint32_t o = 0; o < n; o += 8192) a = warp_adler32(a, p + o, n - o < 8192 ? n - o : 8192, lane);
        }
        if (P.check_kind & 2) {
            uint32_t o = 0;
            while (n - o >= 128) {
                uint32_t k = (n - o) >> 7;
                if (k > 32) k = 32;
                c = warp_crc32_pieces(c, p + o, k, crc_tab, crc_tab + 256, lane);
                o += k * 128;
            }
            if (o < n) {
                uint32_t c2 = 0;
                if (lane == 0) c2 = crc32_serial(c, p + o, n - o, crc_tab);
                c = __shfl_sync(CZK_FULL, c2, 0);
            }
        }
        if (lane == 0) { part_a[warp] = a; part_c[warp] = c; part_n[warp] = n; }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t A = part_a[0], C = part_c[0];
            for (int w = 1; w < 4; w++) {
                if (P.check_kind & 1) A = adler32_combine_u(A, part_a[w], part_n[w]);
                if (P.check_kind & 2) C = crc32_combine_u(C, part_c[w], part_n[w]);
            }
            P.st[seg].adler = A;
            P.st[seg].crc = C;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ K1
// One warp per segment. Step k handles positions 32k..32k+31. Result == inserting the positions one by one.
// head[h] holds the low 16 bits of the most recent position with hash h. To keep "age < 65536" true for every entry (so
// that the 16-bit difference IS the distance) the table is swept every 16384 positions: entries 32768 or more behind are
// re-stamped to exactly 32768 behind, which can never look recent before the next sweep. No global re-check is needed.
#define CZK_CHAIN_SWEEP 16384u
#define CZK_CHAIN_CHUNK 1024u
__global__ void __launch_bounds__(32) deflate_chain_kernel(DeflateParams P, uint32_t seg_begin, uint32_t seg_end) {
    __shared__ uint16_t head[1u << CZK_HASH_BITS];
    __shared__ __align__(16) uint8_t chunk[CZK_CHAIN_CHUNK + 8];
    const uint32_t lane = threadIdx.x;
    for (uint32_t seg = seg_begin + blockIdx.x; seg < seg_end; seg += gridDim.x) {
        const uint8_t *s = P.in + P.seg_off[seg];
        const uint32_t n = seg_len(P, seg);
        uint16_t *pd = P.prevd + seg_base(P, seg);
        __syncwarp();
        for (uint32_t i = lane; i < (1u << CZK_HASH_BITS); i += 32) head[i] = 32768;  // "32768 behind position 0"
        __syncwarp();
        for (uint32_t cb = 0; cb < n; cb += CZK_CHAIN_CHUNK) {
            // stage the next 1 KiB (+3 bytes of lookahead) in shared memory: one memory round trip per 32 steps
            const uint32_t cn = n - cb < CZK_CHAIN_CHUNK + 3 ? n - cb : CZK_CHAIN_CHUNK + 3;
            __syncwarp();
            for (uint32_t i = lane; i < cn; i += 32) chunk[i] = s[cb + i];
            for (uint32_t i = cn + lane; i < CZK_CHAIN_CHUNK + 8; i += 32) chunk[i] = 0;
            __syncwarp();
            const uint32_t cend = cb + CZK_CHAIN_CHUNK < n ? cb + CZK_CHAIN_CHUNK : n;
            // Four steps per iteration. What does not depend on the head table (bytes, hash, the grouping of equal hashes
            // inside the step) is computed for all four first, so that those latencies overlap; only the table lookups
            // and updates run one step after the other.
            for (uint32_t base4 = cb; base4 < cend; base4 += 128) {
                if (base4 && (base4 % CZK_CHAIN_SWEEP) == 0) {
                    for (uint32_t i = lane; i < (1u << CZK_HASH_BITS); i += 32)
                        if (((base4 - (uint32_t)head[i]) & 0xffffu) >= CZK_WINDOW) head[i] = (uint16_t)((base4 - CZK_WINDOW) & 0xffffu);
                    __syncwarp();
                }
                uint32_t hh[4], din[4];
                bool val[4], last[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t pos = base4 + 32u * u + lane;
                    val[u] = pos + 4 <= n;
                    const uint32_t v = lds32u(chunk, pos - cb);
                    // lanes without 4 bytes left get a private pseudo-hash so they never group with real ones
                    hh[u] = val[u] ? hash4(v) : (0x10000u + lane);
                    const uint32_t grp = __match_any_sync(CZK_FULL, hh[u]);
                    const uint32_t lower = grp & ((1u << lane) - 1u);
                    din[u] = lower ? lane - (31u - (uint32_t)__clz((int)lower)) : 0u;  // nearest earlier lane with the same hash
                    last[u] = (grp >> lane) <= 1u;                                     // highest lane of its group
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t pos = base4 + 32u * u + lane;
                    uint32_t d = din[u];
                    if (val[u] && !d) {
                        const uint32_t dd = (pos - (uint32_t)head[hh[u]]) & 0xffffu;
                        if (dd >= 1 && dd < CZK_WINDOW && dd <= pos) d = dd;
                    }
                    if (!val[u]) d = 0;
                    if (pos < n) pd[pos] = (uint16_t)d;
                    __syncwarp();
                    if (val[u] && last[u]) head[hh[u]] = (uint16_t)(pos & 0xffffu);
                    __syncwarp();
                }
            }
        }
    }
}

#ifdef CZ_EXPERIMENTS
// ------------------------------------------------------------------ K1b
// Second link of every chain node: prevd2[i] = prevd[i] + prevd[i - prevd[i]] (0 = none, or farther than the window).
// With it the match search fetches two candidates per memory round trip (find_match).
__global__ void __launch_bounds__(256) deflate_chain2_kernel(DeflateParams P, uint64_t total_bytes) {
    __shared__ uint32_t s_seg;
    const uint64_t g0 = (uint64_t)blockIdx.x * 256;
    if (threadIdx.x == 0) {
        uint32_t lo = 0, hi = P.nseg;  // last segment with base <= g0
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (seg_base(P, mid) <= g0) lo = mid; else hi = mid;
        }
        s_seg = lo;
    }
    __syncthreads();
    const uint64_t g = g0 + threadIdx.x;
    if (g >= total_bytes) return;
    uint32_t seg = s_seg;
    while (seg + 1 < P.nseg && seg_base(P, seg + 1) <= g) seg++;
    const uint64_t base = seg_base(P, seg);
    const uint32_t pos = (uint32_t)(g - base);
    const uint16_t *pd = P.prevd + base;
    const uint32_t d1 = pd[pos];
    uint32_t d2 = 0;
    if (d1 && d1 <= pos) {
        const uint32_t dd = pd[pos - d1];
        if (dd && d1 + dd <= CZK_WINDOW) d2 = d1 + dd;
    }
    P.prevd2[g] = (uint16_t)d2;
}

#endif  // CZ_EXPERIMENTS

// ------------------------------------------------------------------ K2
// 256 consecutive input bytes per CTA; a tile may straddle segment boundaries.
__global__ void __launch_bounds__(256) deflate_match_kernel(DeflateParams P, uint64_t total_bytes) {
    __shared__ uint32_t s_seg;
    const uint64_t g0 = (uint64_t)blockIdx.x * 256;
    if (threadIdx.x == 0) {
        uint32_t lo = 0, hi = P.nseg;  // last segment with base <= g0
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (seg_base(P, mid) <= g0) lo = mid; else hi = mid;
        }
        s_seg = lo;
    }
    __syncthreads();
    const uint64_t g = g0 + threadIdx.x;
    if (g >= total_bytes) return;
    uint32_t seg = s_seg;
    while (seg + 1 < P.nseg && seg_base(P, seg + 1) <= g) seg++;
    const uint64_t base = seg_base(P, seg);
    const uint32_t pos = (uint32_t)(g - base);
    P.match[g] = P.tune.level0 ? 0u : find_match(P.in + P.seg_off[seg], seg_len(P, seg), P.prevd + base, pos, P.tune, P.prevd2 ? P.prevd2 + base : nullptr);
}

// ------------------------------------------------------------------ K2 (sweep)
// The same per-position search, different geometry: a persistent CTA of THREADS threads sweeps a contiguous chunk of the
// input, THREADS positions at a time, so that everything its chain walks touch — the 32 KiB of input and the 64 KiB of
// links behind the sweep — stays in the SM's L1 (with 256-position CTAs dealt round robin, the CTAs resident on one SM
// work ~37 KB apart and their neighbourhoods do not fit: 78 % L1 hit rate, profiles/r1_deflate_match_v2_ncu.md).
// MODE 0: every lane runs find_match() for its position, the warp reconverges after each position (lanes with short chains
//         idle while their neighbours walk 16 links: 12.8 of 32 lanes active);
// MODE 2: the walk is flattened — one loop whose body is ONE chain step; a lane that finishes its position takes its next
//         one (positions threadIdx.x, + THREADS, ... of the chunk) in the same loop, so the lanes of a warp work on
//         different positions and stay busy. Same result per position (same candidates, order and rules as find_match).
//         MEASURED SLOWER (159 ms against 92 ms per GiB): every iteration pays for the set-up, step and extension paths;
// MODE 1: the warp-synchronous walk/extend alternation of find_match_warp (experiment, slower).
template <int THREADS, int MINB, int MODE, bool LINKS2 = false>
__global__ void __launch_bounds__(THREADS, MINB) deflate_match_sweep_kernel(DeflateParams P, uint64_t total_bytes_all, uint32_t chunk_bytes,
                                                                            uint32_t seg_begin, uint32_t seg_end) {
    __shared__ uint32_t s_seg;
    // positions of segments [seg_begin, seg_end) (seg_end == 0: every segment)
    const uint64_t g_begin = seg_end ? seg_base(P, seg_begin) : 0;
    const uint64_t total_bytes = seg_end ? seg_base(P, seg_end) : total_bytes_all;
    const uint64_t nchunks = (total_bytes - g_begin + chunk_bytes - 1) / chunk_bytes;
    for (uint64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const uint64_t g0 = g_begin + c * chunk_bytes;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t lo = 0, hi = P.nseg;  // last segment with base <= g0
            while (hi - lo > 1) {
                uint32_t mid = (lo + hi) >> 1;
                if (seg_base(P, mid) <= g0) lo = mid; else hi = mid;
            }
            s_seg = lo;
        }
        __syncthreads();
        uint32_t seg = s_seg;
        if (MODE == 2) {
            const DeflateTuning &t = P.tune;
            uint32_t off = threadIdx.x;
            bool have = false;
            MatchScan m;
            m.cur = nullptr; m.max_len = 0; m.cur4 = m.curw1 = m.curw2 = m.curw3 = 0; m.best_len = m.best_dist = m.cur_end = 0;
            uint32_t total = 0, d = 0, chain = 0, pos = 0;
            uint64_t g = 0;
            const uint16_t *pd = nullptr;
            for (;;) {
                if (!have) {
                    g = g0 + off;
                    if (off >= chunk_bytes || g >= total_bytes) break;
                    off += THREADS;
                    while (seg + 1 < P.nseg && seg_base(P, seg + 1) <= g) seg++;
                    const uint64_t base = seg_base(P, seg);
                    const uint32_t n = seg_len(P, seg);
                    pos = (uint32_t)(g - base);
                    uint32_t max_len = 0;
                    if (pos + CZK_MIN_MATCH <= n) { max_len = n - pos; if (max_len > CZK_MAX_MATCH) max_len = CZK_MAX_MATCH; }
                    if (max_len < 4) { P.match[g] = 0u; continue; }
                    pd = P.prevd + base;
                    const uint8_t *cur = P.in + P.seg_off[seg] + pos;
                    m.cur = cur; m.max_len = max_len;
                    m.cur4 = load32u(cur);
                    m.curw1 = max_len >= 8 ? load32u(cur + 4) : 0u;
                    m.curw2 = max_len >= 12 ? load32u(cur + 8) : 0u;
                    m.curw3 = max_len >= 16 ? load32u(cur + 12) : 0u;
                    m.best_len = 0; m.best_dist = 0; m.cur_end = 0;
                    total = 0; chain = t.max_chain; d = pd[pos];
                    have = true;
                }
                // one link of the chain
                bool fin = true;
                if (d && chain) {
                    chain--;
                    total += d;
                    if (!(total > CZK_WINDOW || total > pos)) {
                        const uint32_t dnext = pd[pos - total];
                        const uint32_t c4 = load32u(m.cur - total);
                        const uint32_t cb = m.best_len >= 4 ? *(m.cur - total + m.best_len) : 0u;
                        fin = match_passes(m, c4, cb) && match_extend(m, total, t, chain);
                        d = dnext;
                    }
                }
                if (fin) {
                    P.match[g] = m.best_len >= CZK_MIN_MATCH ? (m.best_len | (m.best_dist << 9)) : 0u;
                    have = false;
                }
            }
        } else {
            for (uint32_t off0 = 0; off0 < chunk_bytes; off0 += THREADS) {  // (uniform trip count: find_match_warp is warp-synchronous)
                const uint64_t g = g0 + off0 + threadIdx.x;
                const bool valid = off0 + threadIdx.x < chunk_bytes && g < total_bytes;
                if (valid) while (seg + 1 < P.nseg && seg_base(P, seg + 1) <= g) seg++;
                const uint64_t base = seg_base(P, seg);
                const uint32_t pos = valid ? (uint32_t)(g - base) : 0u;
                uint32_t r;
#ifdef CZ_EXPERIMENTS
                if (MODE == 1) r = find_match_warp(P.in + P.seg_off[seg], seg_len(P, seg), P.prevd + base, pos, P.tune, valid);
                else
#endif
                r = valid ? find_match(P.in + P.seg_off[seg], seg_len(P, seg), P.prevd + base, pos, P.tune, LINKS2 ? P.prevd2 + base : nullptr) : 0u;
                if (valid) P.match[g] = r;
            }
        }
    }
}

#ifdef CZ_EXPERIMENTS
// ------------------------------------------------------------------ K2 (candidate pairs; experiment, CZ_MATCH_V=2)
// (Measured on B200: 92 ms per GiB against 77 ms for the simple kernel — every candidate is extended, while find_match()
// skips the ones that cannot beat the best so far — so it is off by default.)
// The same search as find_match(), split where the divergence is (profiles/r1_deflate_match_v2_ncu.md: with one thread
// walking AND comparing, 12 of 32 lanes are active — chains differ in length, matches differ in length):
//   walk     lane = position: follow the chain for up to CZK_MP_BATCH links and only RECORD the candidate distances
//            (one dependent 2-byte load per link, nothing else);
//   compare  the (position, candidate) pairs of the whole warp are compacted into one list and every lane takes a pair:
//            first word, then word-wise extension — all 32 lanes busy whatever the chain lengths were;
//   reduce   lane = position again: scan its candidates in chain order with find_match()'s rules (strictly longer wins,
//            stop at nice_len / max_len), carry (best_len, best_dist) into the next batch of links.
// The first batch is 2 links, so a position inside a long run stops after one 258-byte compare instead of 16.
// Results are identical to find_match() (the quick rejects there never drop a candidate that would win).
#define CZK_MP_BATCH 16u
#define CZK_MP_STRIDE 17u  // u16 slots per lane (odd: lanes start in different banks)
__global__ void __launch_bounds__(256) deflate_match_pairs_kernel(DeflateParams P, uint64_t total_bytes) {
    __shared__ uint32_t s_seg;
    __shared__ uint16_t s_cand[8][32 * CZK_MP_STRIDE];
    __shared__ uint32_t s_pair[8][32 * CZK_MP_BATCH];
    __shared__ uint16_t s_ml[8][32];
    const uint64_t g0 = (uint64_t)blockIdx.x * 256;
    if (threadIdx.x == 0) {
        uint32_t lo = 0, hi = P.nseg;  // last segment with base <= g0
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (seg_base(P, mid) <= g0) lo = mid; else hi = mid;
        }
        s_seg = lo;
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t g = g0 + threadIdx.x;
    const bool in_range = g < total_bytes;
    uint32_t seg = s_seg;
    if (in_range) while (seg + 1 < P.nseg && seg_base(P, seg + 1) <= g) seg++;
    const uint64_t base = seg_base(P, seg);
    const uint32_t pos = in_range ? (uint32_t)(g - base) : 0u;
    const uint32_t n = in_range ? seg_len(P, seg) : 0u;
    const uint8_t *gin = P.in + P.seg_off[0];   // byte g of the launch (segments are contiguous)
    const uint16_t *pd = P.prevd + base;
    uint32_t max_len = 0;
    if (in_range && pos + CZK_MIN_MATCH <= n) { max_len = n - pos; if (max_len > CZK_MAX_MATCH) max_len = CZK_MAX_MATCH; }
    bool active = max_len >= 4;
    uint32_t total = 0, d = active ? pd[pos] : 0u, chain = P.tune.max_chain;
    uint32_t best_len = 0, best_dist = 0;
    const uint32_t nice = P.tune.nice_len;
    uint16_t *cand = s_cand[warp] + lane * CZK_MP_STRIDE;
    uint32_t *pair = s_pair[warp];
    s_ml[warp][lane] = (uint16_t)max_len;
    const uint64_t gw = g0 + warp * 32;  // byte index of lane 0's position
    uint32_t batch = 2;
    while (__any_sync(CZK_FULL, active)) {
        // ---- walk
        uint32_t cnt = 0;
        for (uint32_t k = 0; k < batch; k++) {
            if (active) {
                if (d && chain) {
                    chain--;
                    total += d;
                    if (total > CZK_WINDOW || total > pos) active = false;
                    else {
                        cand[cnt++] = (uint16_t)total;
                        d = pd[pos - total];
                    }
                } else active = false;
            }
        }
        // ---- compact the pairs of the warp
        uint32_t incl = cnt;
#pragma unroll
        for (int sft = 1; sft < 32; sft <<= 1) {
            const uint32_t v = __shfl_up_sync(CZK_FULL, incl, sft);
            if ((int)lane >= sft) incl += v;
        }
        const uint32_t npairs = __shfl_sync(CZK_FULL, incl, 31);
        const uint32_t off = incl - cnt;
        for (uint32_t k = 0; k < cnt; k++) pair[off + k] = (uint32_t)cand[k] | (lane << 25);
        __syncwarp();
        // ---- compare: one pair per lane
        for (uint32_t j = lane; j < npairs; j += 32) {
            const uint32_t e = pair[j];
            const uint32_t owner = e >> 25, tot = e & 0xffffu;
            const uint32_t ml = s_ml[warp][owner];
            const uint8_t *cur = gin + gw + owner;
            const uint8_t *cd = cur - tot;
            uint32_t l = 0;
            if (load32u(cd) == load32u(cur)) {
                l = 4;
                while (l + 4 <= ml) {
                    const uint32_t x = load32u(cd + l) ^ load32u(cur + l);
                    if (x) { l += ctz32(x) >> 3; break; }
                    l += 4;
                }
                if (l + 4 > ml)
                    while (l < ml && cd[l] == cur[l]) l++;
            }
            pair[j] = e | (l << 16);
        }
        __syncwarp();
        // ---- reduce in chain order
        for (uint32_t k = 0; k < cnt; k++) {
            const uint32_t e = pair[off + k];
            const uint32_t l = (e >> 16) & 0x1ffu;
            if (l > best_len) {
                best_len = l;
                best_dist = e & 0xffffu;
                if (l >= nice || l == max_len) { active = false; break; }
            }
        }
        __syncwarp();
        batch = CZK_MP_BATCH;
    }
    if (in_range) P.match[g] = best_len >= CZK_MIN_MATCH ? (best_len | (best_dist << 9)) : 0u;
}

// ------------------------------------------------------------------ K2 (tiled)
// The same search as find_match(), restructured for the machine (profiles/r1_deflate_match_v2_ncu.md: the simple
// thread-per-position kernel runs with 12 of 32 lanes active and waits on L1/L2 for every link of the chain):
// (Measured, profiles/r1_deflate_notes.md: NOT faster than the simple kernel — the per-iteration vote and state bookkeeping
// cost more than the idle lanes they save — so it is off by default; CZ_MATCH_TILED=1 selects it.)
//   * a CTA owns a tile of 4096 positions of ONE segment and stages the bytes and chain links it can reach
//     (32 KiB back, 258 bytes ahead) in shared memory once: every chain step is a shared-memory access;
//   * inside a warp the positions are handed out dynamically: each lane is a small state machine (IDLE -> CHAIN -> EXTEND)
//     that does ONE step per loop iteration — examine one candidate, or compare one more word — so lanes whose chains are
//     short pick up the next position instead of idling while their neighbours walk 16 links.
// Candidate order, reject rules and tie-breaking are those of find_match(); the tests compare the outputs byte for byte.
#define CZK_MT_POS 4096u                              // positions per tile
#define CZK_MT_THREADS 256u
#define CZK_MT_SPAN (CZK_WINDOW + CZK_MT_POS + 264u)  // bytes staged: window + tile + lookahead (258) + slack
#define CZK_MT_SB_BYTES ((CZK_MT_SPAN + 16u + 32u + 15u) & ~15u)  // + alignment shift (16-byte vector loads) + zero tail
__device__ __forceinline__ uint32_t mt_tile_first(const DeflateParams &P, uint32_t seg) { return (uint32_t)(seg_base(P, seg) >> 12) + seg; }
__host__ __device__ inline size_t deflate_match_tiled_smem() { return (size_t)CZK_MT_SB_BYTES + 2 * (size_t)(CZK_WINDOW + CZK_MT_POS + 16); }

__global__ void __launch_bounds__(CZK_MT_THREADS) deflate_match_tiled_kernel(DeflateParams P) {
    CZ_DYNAMIC_SMEM(smem_raw);
    __shared__ uint32_t s_seg, s_tile, s_ok;
    if (threadIdx.x == 0) {
        uint32_t lo = 0, hi = P.nseg;  // last segment with mt_tile_first <= blockIdx.x
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (mt_tile_first(P, mid) <= blockIdx.x) lo = mid; else hi = mid;
        }
        s_seg = lo;
        s_tile = blockIdx.x - mt_tile_first(P, lo);
        s_ok = (uint64_t)s_tile * CZK_MT_POS < seg_len(P, lo);
    }
    __syncthreads();
    if (!s_ok) return;
    const uint32_t seg = s_seg;
    const uint32_t n = seg_len(P, seg);
    const uint64_t base = seg_base(P, seg);
    const uint8_t *src = P.in + P.seg_off[seg];
    const uint16_t *pd = P.prevd + base;
    uint32_t *mt = P.match + base;
    const uint32_t t0 = s_tile * CZK_MT_POS;
    const uint32_t t1 = t0 + CZK_MT_POS < n ? t0 + CZK_MT_POS : n;
    const uint32_t w0 = t0 > CZK_WINDOW ? t0 - CZK_WINDOW : 0;
    const uint32_t wb1 = t1 + 261 < n ? t1 + 261 : n;  // bytes staged: [w0, wb1)
    // 16-byte vector staging: the shared copies start at the enclosing 16-byte boundary of the global data, so position p is
    // at sb[p - w0] with sb shifted by the misalignment (aligned vectors only touch words that hold at least one wanted byte)
    const uint32_t mis_b = (uint32_t)((uintptr_t)(src + w0) & 15);
    const uint32_t mis_p = (uint32_t)(((uintptr_t)(pd + w0) & 15) >> 1);
    uint8_t *sb_raw = smem_raw;
    uint16_t *sp_raw = (uint16_t *)(smem_raw + CZK_MT_SB_BYTES);
    {
        const uint4 *gb = (const uint4 *)(src + w0 - mis_b);
        const uint32_t nvb = (wb1 - w0 + mis_b + 15) >> 4;
        for (uint32_t i = threadIdx.x; i < nvb; i += CZK_MT_THREADS) ((uint4 *)sb_raw)[i] = gb[i];
        const uint4 *gp = (const uint4 *)(pd + w0 - mis_p);
        const uint32_t nvp = (t1 - w0 + mis_p + 7) >> 3;
        for (uint32_t i = threadIdx.x; i < nvp; i += CZK_MT_THREADS) ((uint4 *)sp_raw)[i] = gp[i];
        __syncthreads();
        // bytes past the end of the segment must compare as "no match" deterministically: zero them (words are read 7 bytes ahead)
        const uint32_t endb = wb1 - w0 + mis_b;
        if (threadIdx.x < 32) sb_raw[endb + threadIdx.x] = 0;
    }
    // byte p of the segment sits at sb_raw[p - w0 + mis_b]; words are read from the ALIGNED base with the shift folded in
    const uint8_t *sb = sb_raw;
    const uint32_t bo = mis_b - w0;  // (wraps; only ever added to a position >= w0)
    const uint16_t *sp = sp_raw + mis_p;
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // each warp owns a contiguous share of the tile and hands its positions out lane by lane
    const uint32_t share = (t1 - t0 + 7) / 8;
    uint32_t next = t0 + warp * share;
    const uint32_t wend = next + share < t1 ? next + share : t1;
    const uint32_t max_chain = P.tune.max_chain, nice_len = P.tune.nice_len;
    enum { IDLE = 0, CHAIN = 1, EXTEND = 2 };
    int state = IDLE;
    uint32_t pos = 0, cur4 = 0, max_len = 0, best_len = 0, best_dist = 0, total = 0, d = 0, chain = 0, l = 0, nd = 0;
    for (;;) {
        // ---- vote: run the phase most lanes are waiting for (every phase then executes with many lanes active)
        const uint32_t m_idle = __ballot_sync(CZK_FULL, state == IDLE);
        const uint32_t m_chain = __ballot_sync(CZK_FULL, state == CHAIN);
        const uint32_t n_ext = 32u - (uint32_t)__popc(m_idle) - (uint32_t)__popc(m_chain);
        const uint32_t avail = next < wend ? wend - next : 0;
        const uint32_t n_new = (uint32_t)__popc(m_idle) < avail ? (uint32_t)__popc(m_idle) : avail;
        const uint32_t n_chain = (uint32_t)__popc(m_chain);
        if (!n_chain && !n_ext && !n_new) break;
        if (n_new > n_chain && n_new >= n_ext) {
            // ---- hand out positions to idle lanes
            const uint32_t rank = __popc(m_idle & ((1u << lane) - 1u));
            if (state == IDLE && rank < avail) {
                pos = next + rank;
                max_len = n - pos < CZK_MAX_MATCH ? n - pos : CZK_MAX_MATCH;
                best_len = 0; best_dist = 0; total = 0;
                d = max_len >= 4 ? sp[pos - w0] : 0;  // the last 3 positions of a segment have no link
                chain = max_chain;
                cur4 = lds32u(sb, pos + bo);
                state = CHAIN;
            }
            next += n_new;
        } else if (n_chain >= n_ext) {
            // ---- examine one candidate per lane
            if (state == CHAIN) {
                bool finish = false;
                if (!d || !chain) finish = true;
                else {
                    chain--;
                    total += d;
                    if (total > CZK_WINDOW || total > pos) finish = true;
                    else {
                        const uint32_t cp = pos - total;  // candidate position
                        nd = sp[cp - w0];
                        bool take = lds32u(sb, cp + bo) == cur4;
                        if (take && best_len >= 4 && best_len < max_len)
                            take = lds32u(sb, cp + bo + best_len - 3) == lds32u(sb, pos + bo + best_len - 3);
                        if (take) { l = 4; state = EXTEND; }
                        else d = nd;
                    }
                }
                if (finish) {
                    mt[pos] = best_len >= CZK_MIN_MATCH ? (best_len | (best_dist << 9)) : 0u;
                    state = IDLE;
                }
            }
        } else if (state == EXTEND) {
            // ---- compare one more word per lane
            const uint32_t co = pos - total + bo, po = pos + bo;
            bool done = false;
            if (l + 4 <= max_len) {
                const uint32_t x = lds32u(sb, co + l) ^ lds32u(sb, po + l);
                if (x) { l += ((uint32_t)__ffs((int)x) - 1u) >> 3; done = true; }
                else l += 4;
            } else {
                // fewer than 4 bytes left: compare them in one masked word
                const uint32_t rem = max_len - l;
                uint32_t x = lds32u(sb, co + l) ^ lds32u(sb, po + l);
                x &= rem ? (0xffffffffu >> (32 - 8 * rem)) : 0u;
                l = x ? l + (((uint32_t)__ffs((int)x) - 1u) >> 3) : max_len;
                done = true;
            }
            if (done) {
                state = CHAIN;
                d = nd;
                if (l > best_len) {
                    best_len = l;
                    best_dist = total;
                    if (l >= nice_len || l == max_len) {  // good enough: stop searching
                        mt[pos] = best_len | (best_dist << 9);
                        state = IDLE;
                    }
                }
            }
        }
    }
}

#endif  // CZ_EXPERIMENTS

// ------------------------------------------------------------------ K3
// One warp per segment. The parse is a serial chain (the next position depends on the match taken at this one). The warp
// stages a tile of 2048 positions in shared memory with coalesced loads and all lanes precompute "advance if a token starts
// here" (match length, or 1 for a literal / a match deferred by the one-step lazy rule). The chain itself is then walked
// SPECULATIVELY: lane s walks sub-tile s (64 positions) from its first position, in parallel with the other lanes; parses
// that start at different positions fall into step within a few tokens, so the true walk — sub-tile after sub-tile, entering
// where the previous one left — only has to go until it meets a position the speculative walk also visited, and takes the
// rest of that sub-tile (and its exit) from the speculation. All lanes then emit the marked tokens with a popcount prefix.
// Same decisions as parse_segment() in deflate_core.cuh (the sequential statement the tests compare against).
// SUB = positions per speculative sub-tile (tile = 32 x SUB). SUB 64: 12.5 KB of shared memory per warp, 16 warps per SM;
// SUB 32: 6.3 KB, 28 warps per SM — all 4 096 segments of a 4 GiB launch are resident at once instead of two waves of
// 2 368 (43 -> 25 ms), at the price of twice as many tile set-ups per segment (slower when one wave suffices anyway).
template <uint32_t CZK_PARSE_SUB>
__global__ void __launch_bounds__(32) deflate_parse_kernel(DeflateParams P) {
    constexpr uint32_t CZK_PARSE_TILE = 32u * CZK_PARSE_SUB;
    auto parse_pad = [](uint32_t i) { return i + i / CZK_PARSE_SUB; };  // one pad entry per sub-tile: no bank conflicts
    __shared__ uint32_t raw_s[CZK_PARSE_TILE + 1];
    __shared__ uint16_t adv_s[CZK_PARSE_TILE + CZK_PARSE_TILE / CZK_PARSE_SUB];
    __shared__ uint32_t vis_s[CZK_PARSE_TILE / 32];
    const uint32_t lane = threadIdx.x;
    for (uint32_t seg = blockIdx.x; seg < P.nseg; seg += gridDim.x) {
        const uint64_t base = seg_base(P, seg);
        const uint8_t *src = P.in + P.seg_off[seg];
        const uint32_t n = seg_len(P, seg);
        uint32_t *mt = P.match + base;
        uint32_t *blk_end = P.blk_end + blk_first(P, seg);
        uint32_t nt = 0;      // tokens emitted so far
        uint32_t entry = 0;   // first token start inside the current tile (relative)
        for (uint32_t t0 = 0; t0 < n; t0 += CZK_PARSE_TILE) {
            const uint32_t T = n - t0 < CZK_PARSE_TILE ? n - t0 : CZK_PARSE_TILE;
            __syncwarp();
            for (uint32_t i = lane; i <= T; i += 32) raw_s[i] = t0 + i < n ? mt[t0 + i] : 0u;
            __syncwarp();
            for (uint32_t i = lane; i < T; i += 32) {
                const uint32_t len = raw_s[i] & 0x1ff;
                const uint32_t nlen = (P.tune.lazy && t0 + i + 1 < n) ? (raw_s[i + 1] & 0x1ff) : 0;
                adv_s[parse_pad(i)] = (uint16_t)((len >= CZK_MIN_MATCH && !(nlen > len)) ? len : 1u);
            }
            __syncwarp();
            // ---- speculative walk of sub-tile `lane` from its first position
            const uint32_t s0 = lane * CZK_PARSE_SUB;                          // first position of my sub-tile
            const uint32_t slen = s0 < T ? (T - s0 < CZK_PARSE_SUB ? T - s0 : CZK_PARSE_SUB) : 0;
            unsigned long long m0 = 0;
            uint32_t p = 0;
            while (p < slen) { m0 |= 1ull << p; p += adv_s[parse_pad(s0 + p)]; }
            const uint32_t exit0 = s0 + p;                                     // where the speculative walk leaves (tile-relative)
            // ---- the true walk, sub-tile after sub-tile
            unsigned long long fin = 0;
            uint32_t cur = entry;
            const uint32_t nsub = (T + CZK_PARSE_SUB - 1) / CZK_PARSE_SUB;
            for (uint32_t s = 0; s < nsub; s++) {
                uint32_t nxt = cur;
                if (lane == s && cur < s0 + slen) {  // (cur >= s0 always: the walk never goes backwards)
                    uint32_t q = cur - s0;
                    unsigned long long pre = 0;
                    while (q < slen && !((m0 >> q) & 1ull)) { pre |= 1ull << q; q += adv_s[parse_pad(s0 + q)]; }
                    if (q < slen) { fin = pre | (m0 & ~((1ull << q) - 1ull)); nxt = exit0; }  // met the speculation: rest is identical
                    else { fin = pre; nxt = s0 + q; }
                }
                cur = __shfl_sync(CZK_FULL, nxt, s);
            }
            entry = cur - T;
            if (CZK_PARSE_SUB == 64u) {
                vis_s[2 * lane] = (uint32_t)fin;
                vis_s[2 * lane + 1] = (uint32_t)(fin >> 32);
            } else vis_s[lane] = (uint32_t)fin;
            __syncwarp();
            for (uint32_t w = 0; w * 32 < T; w++) {
                const uint32_t bits = vis_s[w];
                if (bits >> lane & 1u) {
                    const uint32_t i = w * 32 + lane;
                    const uint32_t k = nt + (uint32_t)__popc(bits & ((1u << lane) - 1u));
                    const uint32_t a = adv_s[parse_pad(i)];
                    mt[k] = a == 1 ? (uint32_t)src[t0 + i] << 9 : raw_s[i];  // k <= t0 + i: never ahead of the tile being read
                    if ((k + 1) % CZK_BLOCK_TOKENS == 0) blk_end[(k + 1) / CZK_BLOCK_TOKENS - 1] = t0 + i + a;
                }
                nt += (uint32_t)__popc(bits);
            }
        }
        if (lane == 0) {
            uint32_t nb = nt / CZK_BLOCK_TOKENS;
            if (nt % CZK_BLOCK_TOKENS) blk_end[nb++] = n;  // last, partial block
            P.st[seg].ntok = nt;
            P.st[seg].nblk = nb;
        }
    }
}

// ------------------------------------------------------------------ K4a
__global__ void __launch_bounds__(128) deflate_hist_kernel(DeflateParams P) {
    __shared__ uint32_t hist[CZK_FREQ_STRIDE];
    __shared__ uint32_t s_seg, s_b, s_ok;
    if (threadIdx.x == 0) { uint32_t sg, b; s_ok = slot_to_block(P, blockIdx.x, sg, b); s_seg = sg; s_b = b; }
    for (uint32_t i = threadIdx.x; i < CZK_FREQ_STRIDE; i += 128) hist[i] = 0;
    __syncthreads();
    if (!s_ok) return;
    const uint32_t seg = s_seg, b = s_b;
    const uint32_t t0 = b * CZK_BLOCK_TOKENS;
    uint32_t t1 = t0 + CZK_BLOCK_TOKENS;
    if (t1 > P.st[seg].ntok) t1 = P.st[seg].ntok;
    const uint32_t *tok = P.match + seg_base(P, seg);
    for (uint32_t k = t0 + threadIdx.x; k < t1; k += 128) {
        uint32_t t = tok[k], len = t & 0x1ff;
        if (len < CZK_MIN_MATCH) atomicAdd(&hist[t >> 9], 1u);
        else { atomicAdd(&hist[len_code(len)], 1u); atomicAdd(&hist[288 + dist_code(t >> 9)], 1u); }
    }
    __syncthreads();
    uint32_t *f = P.freqs + (size_t)blockIdx.x * CZK_FREQ_STRIDE;
    for (uint32_t i = threadIdx.x; i < CZK_FREQ_STRIDE; i += 128) f[i] = i == 256 ? 1u : hist[i];
}

// ------------------------------------------------------------------ K4b
__global__ void __launch_bounds__(32) deflate_plan_kernel(DeflateParams P) {
    const uint32_t slot = blockIdx.x * 32 + threadIdx.x;
    if (slot >= P.n_slots) return;
    uint32_t seg, b;
    if (!slot_to_block(P, slot, seg, b)) return;
    BlockPlan &bp = P.plans[slot];
    bp.tok_begin = b * CZK_BLOCK_TOKENS;
    uint32_t t1 = bp.tok_begin + CZK_BLOCK_TOKENS;
    bp.tok_end = t1 < P.st[seg].ntok ? t1 : P.st[seg].ntok;
    bp.in_begin = b ? P.blk_end[slot - 1] : 0;
    bp.in_end = P.blk_end[slot];
    HuffScratch hs;
    const uint32_t *f = P.freqs + (size_t)slot * CZK_FREQ_STRIDE;
    plan_block(f, f + 288, bp, hs, P.tune);
}

// bits a block occupies when it starts at bit position `bits` (stored blocks pad to a byte boundary per piece)
__host__ __device__ inline uint64_t block_end_bit(const BlockPlan &bp, uint64_t bits) {
    if (bp.btype != 0) return bits + bp.hdr_bits + bp.body_bits;
    uint32_t left = bp.in_end - bp.in_begin;
    do {
        uint32_t k = left > 65535 ? 65535 : left;
        bits = ((bits + 3 + 7) & ~(uint64_t)7) + 32 + 8ull * k;
        left -= k;
    } while (left);
    return bits;
}

// ------------------------------------------------------------------ K5a
__global__ void __launch_bounds__(32) deflate_seg_layout_kernel(DeflateParams P) {
    const uint32_t seg = blockIdx.x * 32 + threadIdx.x;
    if (seg >= P.nseg) return;
    uint64_t bits = 0;  // a segment starts on a byte boundary
    const uint32_t nb = P.st[seg].nblk, first = blk_first(P, seg);
    for (uint32_t b = 0; b < nb; b++) {
        BlockPlan &bp = P.plans[first + b];
        bp.bit_off = bits;
        bits = block_end_bit(bp, bits);
    }
    P.st[seg].body_bits = bits;
    P.st[seg].out_bytes = ((bits + 3 + 7) >> 3) + 4;  // + empty stored block: 3 bits, pad, 00 00 ff ff
}

__device__ __forceinline__ void unit_segs(const DeflateParams &P, uint32_t u, uint32_t &s0, uint32_t &s1) {
    if (P.unit_seg) { s0 = P.unit_seg[u]; s1 = P.unit_seg[u + 1]; }
    else { s0 = u; s1 = u + 1; }
}
__device__ __forceinline__ uint32_t container_hdr(const DeflateParams &P) {
    const int wb = P.piece_mode ? -15 : P.window_bits;
    return wb == 15 ? 2 : wb > 15 ? 10 : 0;
}
__device__ __forceinline__ uint32_t container_trl(const DeflateParams &P) {
    const int wb = P.piece_mode ? -15 : P.window_bits;
    return wb == 15 ? 4 : wb > 15 ? 8 : 0;
}

// ------------------------------------------------------------------ K5c
__global__ void __launch_bounds__(32) deflate_unit_size_kernel(DeflateParams P) {
    const uint32_t u = blockIdx.x * 32 + threadIdx.x;
    if (u >= P.n_units) return;
    uint32_t s0, s1;
    unit_segs(P, u, s0, s1);
    const uint64_t cap = P.unit_out_off[u + 1] - P.unit_out_off[u];
    uint64_t total = container_hdr(P);
    for (uint32_t s = s0; s < s1; s++) total += P.st[s].out_bytes;
    total += (P.piece_mode ? 0 : 2) + container_trl(P);
    P.unit_out_len[u] = total;
    P.unit_status[u] = total > cap ? 1 /* CZ_ENCODE_NEED_OUTPUT: nothing of this unit is written */ : 2 /* FINISHED */;
    if (P.unit_checks) {
        // checksums of the unit's input: fold of the per-segment values
        uint32_t adler = 1, crc = 0;
        uint32_t pw_len = 0xffffffffu, pw = 0;  // cached x^(8 len) for the CRC combine (segments share one length)
        for (uint32_t s = s0; s < s1; s++) {
            const uint32_t len = seg_len(P, s);
            if (P.check_kind & 1) adler = adler32_combine_u(adler, P.st[s].adler, len);
            if (P.check_kind & 2) {
                if (len != pw_len) { pw = crc_xpow8n(len); pw_len = len; }
                crc = crc_mulmod(pw, crc) ^ P.st[s].crc;
            }
        }
        P.unit_checks[2 * u] = adler;
        P.unit_checks[2 * u + 1] = crc;
    }
}

// ------------------------------------------------------------------ K5s
// Exclusive scan of the sizes of the units that fit -> packed output positions. One CTA of 1024 threads.
__global__ void __launch_bounds__(1024) deflate_scan_kernel(DeflateParams P) {
    __shared__ uint64_t part[1024];
    const uint32_t t = threadIdx.x, n = P.n_units;
    const uint32_t per = (n + 1023) / 1024;
    const uint32_t b = t * per, e = b + per < n ? b + per : n;
    uint64_t sum = 0;
    for (uint32_t u = b; u < e; u++) sum += P.unit_status[u] == 2 ? P.unit_out_len[u] : 0;
    part[t] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {
        uint64_t v = t >= d ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint64_t run = part[t] - sum;
    for (uint32_t u = b; u < e; u++) {
        P.unit_out_pos[u] = run;
        run += P.unit_status[u] == 2 ? P.unit_out_len[u] : 0;
    }
    if (t == 1023 && P.total_out) *P.total_out = part[1023];
}

// ------------------------------------------------------------------ K5d
__global__ void __launch_bounds__(32) deflate_unit_frame_kernel(DeflateParams P) {
    const uint32_t u = blockIdx.x * 32 + threadIdx.x;
    if (u >= P.n_units) return;
    uint32_t s0, s1;
    unit_segs(P, u, s0, s1);
    if (P.unit_status[u] != 2) {
        for (uint32_t s = s0; s < s1; s++) P.st[s].out_off = ~0ull;
        return;
    }
    const uint64_t o0 = P.unit_out_pos ? P.unit_out_pos[u] : P.unit_out_off[u];
    const int wb = P.piece_mode ? -15 : P.window_bits;
    uint8_t *o = P.out + o0;
    uint64_t pos = 0;
    if (wb == 15) {
        // CMF = 0x78 (deflate, 32 KiB window); FLG: FLEVEL by level like zlib, FCHECK makes (CMF<<8|FLG) % 31 == 0
        uint32_t lvl = P.level < 0 ? 6 : P.level;
        uint32_t flevel = lvl < 2 ? 0 : lvl < 6 ? 1 : lvl == 6 ? 2 : 3;
        uint32_t h = (0x78u << 8) | (flevel << 6);
        h += 31 - (h % 31);
        o[0] = (uint8_t)(h >> 8);
        o[1] = (uint8_t)h;
        pos = 2;
    } else if (wb > 15) {
        uint32_t lvl = P.level < 0 ? 6 : P.level;
        const uint8_t g[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, (uint8_t)(lvl == 9 ? 2 : lvl == 1 ? 4 : 0), 3};
        for (int i = 0; i < 10; i++) o[i] = g[i];
        pos = 10;
    }
    for (uint32_t s = s0; s < s1; s++) {
        P.st[s].out_off = o0 + pos;
        if (P.seg_out_bytes) P.seg_out_bytes[s] = P.st[s].out_bytes;
        // flush marker of the segment: bytes wholly owned by it are written here, the shared byte is OR-ed by emit
        const uint64_t bits = P.st[s].body_bits;
        const uint64_t mark = (bits + 3 + 7) >> 3;  // first byte of 00 00 ff ff
        for (uint64_t k = (bits + 7) >> 3; k < mark; k++) o[pos + k] = 0;
        o[pos + mark] = 0; o[pos + mark + 1] = 0; o[pos + mark + 2] = 0xff; o[pos + mark + 3] = 0xff;
        pos += P.st[s].out_bytes;
    }
    if (!P.piece_mode) {
        o[pos++] = 0x03;  // final empty fixed block: BFINAL=1, BTYPE=01, end-of-block code 0000000
        o[pos++] = 0x00;
    }
    if (wb == 15) {
        const uint32_t adler = P.unit_checks[2 * u];
        o[pos++] = (uint8_t)(adler >> 24); o[pos++] = (uint8_t)(adler >> 16); o[pos++] = (uint8_t)(adler >> 8); o[pos++] = (uint8_t)adler;
    } else if (wb > 15) {
        const uint32_t crc = P.unit_checks[2 * u + 1];
        const uint64_t in_len = P.seg_off[s1] - P.seg_off[s0];
        for (int i = 0; i < 4; i++) o[pos++] = (uint8_t)(crc >> (8 * i));
        for (int i = 0; i < 4; i++) o[pos++] = (uint8_t)((uint32_t)in_len >> (8 * i));  // ISIZE = length mod 2^32
    }
}

// ------------------------------------------------------------------ K5b
__global__ void __launch_bounds__(128) deflate_zero_kernel(DeflateParams P) {
    const uint32_t slot = blockIdx.x * 128 + threadIdx.x;
    if (slot >= P.n_slots) return;
    uint32_t seg, b;
    if (!slot_to_block(P, slot, seg, b) || P.st[seg].out_off == ~0ull) return;
    const BlockPlan &bp = P.plans[slot];
    uint8_t *o = P.out + P.st[seg].out_off;
    const uint64_t e = block_end_bit(bp, bp.bit_off);
    o[bp.bit_off >> 3] = 0;          // first byte of the block (may hold the tail of the previous block)
    if (e & 7) o[e >> 3] = 0;        // last, partial byte of the block
}

// OR `v` into output byte `idx` (the byte is shared with a neighbouring block, so it was cleared by K5b)
__device__ __forceinline__ void or_byte(uint8_t *o, uint64_t idx, uint32_t v) {
    uintptr_t a = (uintptr_t)(o + idx);
    atomicOr((unsigned int *)(a & ~(uintptr_t)3), v << (8 * (a & 3)));
}

// ------------------------------------------------------------------ K6
#define CZK_EMIT_THREADS 256
#define CZK_EMIT_WORDS (CZK_EMIT_THREADS * 48 / 32 + 4)
__global__ void __launch_bounds__(CZK_EMIT_THREADS) deflate_emit_kernel(DeflateParams P) {
    __shared__ uint32_t buf[CZK_EMIT_WORDS];
    __shared__ uint32_t warp_sum[CZK_EMIT_THREADS / 32];
    __shared__ uint32_t carry_bits_s, carry_val_s;
    __shared__ uint32_t s_seg, s_b, s_ok;
    if (threadIdx.x == 0) {
        uint32_t sg, bb;
        s_ok = slot_to_block(P, blockIdx.x, sg, bb) && P.st[sg].out_off != ~0ull;
        s_seg = sg; s_b = bb;
    }
    __syncthreads();
    if (!s_ok) return;
    const uint32_t seg = s_seg;
    const BlockPlan &bp = P.plans[blockIdx.x];
    uint8_t *o = P.out + P.st[seg].out_off;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (bp.btype == 0) {
        // stored block(s): [3 zero bits + pad] [LEN NLEN] [bytes]; only the very first byte can be shared
        const uint8_t *src = P.in + P.seg_off[seg] + bp.in_begin;
        uint32_t left = bp.in_end - bp.in_begin;
        uint64_t bits = bp.bit_off;
        bool first = true;
        do {
            const uint32_t k = left > 65535 ? 65535 : left;
            const uint64_t hb = (bits + 3 + 7) >> 3;  // byte index of LEN
            if (tid == 0) {
                // header bits are zeros: the shared first byte was cleared by K5b; later piece headers own their byte
                if (!first) o[bits >> 3] = 0;
                if (((bits + 2) >> 3) != (bits >> 3)) o[(bits + 2) >> 3] = 0;  // 3 header bits spill into the next byte
                o[hb] = (uint8_t)k; o[hb + 1] = (uint8_t)(k >> 8); o[hb + 2] = (uint8_t)~k; o[hb + 3] = (uint8_t)(~k >> 8);
            }
            for (uint32_t i = tid; i < k; i += CZK_EMIT_THREADS) o[hb + 4 + i] = src[i];
            src += k;
            left -= k;
            bits = (hb + 4 + k) * 8;
            first = false;
        } while (left);
        return;
    }
    // ---- Huffman block: header words, then tokens (+ end-of-block), CZK_EMIT_THREADS items per round
    const uint32_t *tok = P.match + seg_base(P, seg);
    const uint32_t hdr_words = (bp.hdr_bits + 31) >> 5;
    const uint32_t ntok = bp.tok_end - bp.tok_begin;
    const uint32_t items = hdr_words + ntok + 1;  // + end-of-block
    uint64_t gbit = bp.bit_off;                   // next free bit of the segment's output
    if (tid == 0) { carry_bits_s = (uint32_t)(gbit & 7); carry_val_s = 0; }
    for (uint32_t base = 0; base < items; base += CZK_EMIT_THREADS) {
        for (uint32_t i = tid; i < CZK_EMIT_WORDS; i += CZK_EMIT_THREADS) buf[i] = 0;
        const uint32_t it = base + tid;
        uint64_t v = 0;
        uint32_t n = 0;
        if (it < hdr_words) {
            v = bp.hdr[it];
            n = it + 1 == hdr_words ? bp.hdr_bits - 32 * it : 32;
        } else if (it < hdr_words + ntok) {
            v = token_bits(tok[bp.tok_begin + it - hdr_words], bp, &n);
        } else if (it == hdr_words + ntok) {
            v = bp.lit_code[256];
            n = bp.lit_len[256];
        }
        // exclusive scan of n over the CTA
        uint32_t x = n;
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(CZK_FULL, x, d);
            if ((int)lane >= d) x += y;
        }
        if (lane == 31) warp_sum[wid] = x;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
        for (uint32_t w = 0; w < CZK_EMIT_THREADS / 32; w++) {
            uint32_t sx = warp_sum[w];
            if (w < wid) wbase += sx;
            total += sx;
        }
        const uint32_t cb = carry_bits_s;      // bits (< 8) already pending in the first byte
        uint32_t off = cb + wbase + x - n;     // bit offset inside buf
        if (tid == 0 && cb) atomicOr(&buf[0], carry_val_s);
        if (n) {
            uint32_t w = off >> 5, sh = off & 31;
            atomicOr(&buf[w], (uint32_t)(v << sh));
            uint64_t hi = sh ? (v >> (32 - sh)) : (v >> 32);
            if (sh + n > 32) atomicOr(&buf[w + 1], (uint32_t)hi);
            if (sh + n > 64) atomicOr(&buf[w + 2], (uint32_t)(hi >> 32));
        }
        __syncthreads();
        const uint32_t tbits = cb + total;               // valid bits in buf
        const uint32_t full = tbits >> 3;                // complete bytes
        const uint64_t byte0 = gbit >> 3;                // output byte of buf bit 0
        const bool last_round = base + CZK_EMIT_THREADS >= items;
        const bool first_shared = (base == 0) && (bp.bit_off & 7);  // byte 0 also holds bits of the previous block
        for (uint32_t i = tid; i < full; i += CZK_EMIT_THREADS) {
            uint32_t byte = (buf[i >> 2] >> (8 * (i & 3))) & 0xff;
            if (i == 0 && first_shared) or_byte(o, byte0, byte); else o[byte0 + i] = (uint8_t)byte;
        }
        if (tid == 0) {
            uint32_t rem = tbits & 7;
            uint32_t rv = rem ? ((buf[full >> 2] >> (8 * (full & 3))) & 0xff) : 0;
            if (last_round) {
                if (rem) or_byte(o, byte0 + full, rv);  // last partial byte: shared with the next block / flush marker
            } else {
                carry_bits_s = rem;
                carry_val_s = rv;
            }
        }
        gbit += total;
        __syncthreads();
    }
}

}  // namespace czk
