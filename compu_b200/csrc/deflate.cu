// deflate.cu — encoder side of the C ABI: the kernel chain launcher, the batched host engine (segmenting, batching,
// double-buffered H2D / kernels / D2H per device, host-side gather) and the streaming Encoder.
//
// Reference contract followed here (file:line under /root/reference):
//   cz_encoder_new       src/encoder/zlib_ng.rs:50-87  (deflateInit2_(level, Z_DEFLATED, windowBits, memLevel, strategy))
//   cz_encode            src/encoder/mod.rs:334-370    (internal_zlib_impl_encode!: op map + status map)
//   cz_encoder_reset     src/encoder/zlib_ng.rs:95-104 (deflateReset)
//   cz_encoder_free      src/encoder/zlib_ng.rs:107-111
// Multi-GPU: segments are independent, so batches of segments are dealt to the devices in contiguous
// ranges with no collective; the gather (offsets by prefix sum, checksum combine) is host-side scalar work.
#include <stdlib.h>

#include <algorithm>
#include <thread>

#include "deflate_kernels.cuh"
#include <chrono>
#include <map>
#include "host_common.h"

namespace czh {

using czk::BlockPlan;
using czk::DeflateParams;
using czk::SegState;

static const uint64_t kDefaultSegment = 1u << 20;
static const uint64_t kMinSegment = 4096;

static uint64_t clamp_segment(uint64_t s) {
    if (s == 0) s = kDefaultSegment;
    if (s < kMinSegment) s = kMinSegment;
    if (s > CZK_SEG_MAX) s = CZK_SEG_MAX;
    return s;
}

// worst case of one segment: every block stored (5 bytes per 65535-byte piece + a pad byte), plus the flush marker
static inline uint64_t segment_bound(uint64_t len) { return len + len / 2048 + 64; }

// ---- workspace carving (device) ---------------------------------------------------------------------------------
struct WsLayout {
    uint64_t st, prevd, prevd2, match, blk_end, freqs, plans, pos, total, n_slots, bytes;
};
// CZ_MATCH_LINKS=2: build the second-link array and fetch two candidates per step of the match search (measured 124.5 ms per
// GiB against 110.4 ms for the single-link walk: the speculative loads of the second candidate cost more than the round trip
// they save). Off by default; the array is only carved out of the workspace when it is on.
static bool match_two_links() {
#ifdef CZ_EXPERIMENTS
    static int v = -1;
    if (v < 0) { const char *e = getenv("CZ_MATCH_LINKS"); v = e ? atoi(e) == 2 : 0; }
    return v != 0;
#else
    return false;
#endif
}

static WsLayout ws_layout(uint64_t nseg, uint64_t n_units, uint64_t in_bytes) {
    WsLayout w;
    w.n_slots = (in_bytes >> 14) + nseg + 1;
    uint64_t o = 0;
    w.st = o; o = align_up(o + nseg * sizeof(SegState), 256);
    w.prevd = o; o = align_up(o + 2 * (in_bytes + 8), 256);
    w.prevd2 = o; if (match_two_links()) o = align_up(o + 2 * (in_bytes + 8), 256);
    w.match = o; o = align_up(o + 4 * (in_bytes + 8), 256);
    w.blk_end = o; o = align_up(o + 4 * w.n_slots, 256);
    w.freqs = o; o = align_up(o + 4ull * CZK_FREQ_STRIDE * w.n_slots, 256);
    w.plans = o; o = align_up(o + sizeof(BlockPlan) * w.n_slots, 256);
    w.pos = o; o = align_up(o + 8 * n_units, 256);
    w.total = o; o += 256;
    w.bytes = o;
    return w;
}

// A side stream (with its fork / join events) per caller stream, created on first use and kept.
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr, ev_a = nullptr, ev_b = nullptr;
};
static SideStream *side_stream_for(cudaStream_t st) {
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, SideStream *> pool;
    static const bool no_side = getenv("CZ_NO_SIDE_STREAM") != nullptr;
    if (no_side) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_pair(dev, st);
    auto it = pool.find(key);
    if (it != pool.end()) return it->second;
    SideStream *s = new (std::nothrow) SideStream();
    if (!s) return nullptr;
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_a, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_b, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        delete s;
        s = nullptr;
    }
    pool[key] = s;
    return s;
}

struct DeflateProf { cudaEvent_t e0, m0, m1, e1; };
static std::vector<DeflateProf> g_dprof;
static std::mutex g_dprof_mu;

struct DeflateLaunch {
    const uint8_t *d_in;          // base of the input; segment s starts at d_in + seg_off[s]
    uint8_t *d_out;
    const uint64_t *d_seg_off;    // nseg+1
    const uint32_t *d_unit_seg;   // n_units+1 or null (unit == segment)
    const uint64_t *d_unit_out_off;
    uint64_t *d_unit_out_len;
    int32_t *d_unit_status;
    uint32_t *d_unit_checks;
    uint64_t *d_seg_out_bytes;
    uint64_t nseg, n_units, in_bytes;
    int level, strategy, window_bits, piece_mode, check_kind;
    bool packed;                  // write the units back to back from d_out (positions by the scan kernel)
    uint64_t **d_unit_out_pos_ret, **d_total_ret;
};

static int launch_deflate(cudaStream_t st, DeviceCtx *ctx, const DeflateLaunch &L, void *d_ws, uint64_t ws_bytes) {
    if (L.nseg == 0 || L.n_units == 0) return 0;
    if (L.nseg > 0x7ffffff0ull || L.in_bytes > (1ull << 40)) { set_error("deflate launch too large"); return CZ_E_STREAM; }
    const WsLayout w = ws_layout(L.nseg, L.n_units, L.in_bytes);
    if (!d_ws || ws_bytes < w.bytes) { set_error("deflate workspace too small (%llu < %llu)", (unsigned long long)ws_bytes, (unsigned long long)w.bytes); return CZ_E_MEM; }
    uint8_t *ws = (uint8_t *)d_ws;
    DeflateParams P;
    memset(&P, 0, sizeof P);
    P.in = L.d_in; P.out = L.d_out; P.seg_off = L.d_seg_off;
    P.nseg = (uint32_t)L.nseg; P.n_units = (uint32_t)L.n_units; P.n_slots = (uint32_t)w.n_slots;
    P.unit_seg = L.d_unit_seg; P.unit_out_off = L.d_unit_out_off;
    P.unit_out_pos = L.packed ? (uint64_t *)(ws + w.pos) : nullptr;
    P.total_out = L.packed ? (uint64_t *)(ws + w.total) : nullptr;
    P.unit_out_len = L.d_unit_out_len; P.unit_status = L.d_unit_status; P.unit_checks = L.d_unit_checks;
    P.seg_out_bytes = L.d_seg_out_bytes;
    P.st = (SegState *)(ws + w.st); P.prevd = (uint16_t *)(ws + w.prevd); P.prevd2 = (uint16_t *)(ws + w.prevd2); P.match = (uint32_t *)(ws + w.match);
    P.blk_end = (uint32_t *)(ws + w.blk_end); P.freqs = (uint32_t *)(ws + w.freqs); P.plans = (BlockPlan *)(ws + w.plans);
    P.crc = ctx->d_crc;
    P.tune = czk::deflate_tuning(L.level, L.strategy);
    {   // experiment knob (tools/sweep_deflate_ratio.py picks, measured on the GPU with these): CZ_DEFLATE_CHAIN / CZ_DEFLATE_NICE
        static const int ov_chain = [] { const char *e = getenv("CZ_DEFLATE_CHAIN"); return e ? atoi(e) : 0; }();
        static const int ov_nice = [] { const char *e = getenv("CZ_DEFLATE_NICE"); return e ? atoi(e) : 0; }();
        if (ov_chain > 0 && P.tune.max_chain) P.tune.max_chain = (uint32_t)ov_chain;
        if (ov_nice > 0 && P.tune.max_chain) P.tune.nice_len = (uint32_t)ov_nice;
    }
    P.window_bits = L.window_bits; P.level = L.level; P.piece_mode = L.piece_mode; P.check_kind = L.check_kind;
    if (!L.piece_mode && L.window_bits != -15 && !L.d_unit_checks) { set_error("zlib/gzip framing needs the checks array"); return CZ_E_STREAM; }
    if (L.d_unit_out_pos_ret) *L.d_unit_out_pos_ret = P.unit_out_pos;
    if (L.d_total_ret) *L.d_total_ret = P.total_out;
    const unsigned nseg = P.nseg, nun = P.n_units, nsl = P.n_slots;
    // bench.py's per-kernel timing (cz_profile_enable): events around the whole chain and around the match search
    const bool prof = profiling_on();
    DeflateProf pr = {nullptr, nullptr, nullptr, nullptr};
    if (prof) {
        cudaEventCreate(&pr.e0); cudaEventCreate(&pr.m0); cudaEventCreate(&pr.m1); cudaEventCreate(&pr.e1);
        cudaEventRecord(pr.e0, st);
    }
    // The checksum pass only feeds the container framing (K5c), so it runs on a side stream next to the chain and match
    // kernels (its 128-thread CTAs fit beside the match search's one 1 024-thread CTA per SM) and is joined before K5c.
    SideStream *side = P.check_kind ? side_stream_for(st) : nullptr;
    unsigned split_at = 0;  // != 0: segments [split_at, nseg) get their chains on the side stream
    if (P.check_kind) {
        if (side && CZ_CUDA(cudaEventRecord(side->fork, st)) && CZ_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0))) {
            CZ_KL(czk::deflate_checksum_kernel<<<nseg < 65535u * 8u ? nseg : 65535u * 8u, 128, 0, side->stream>>>(P));
            if (!CZ_CUDA(cudaEventRecord(side->join, side->stream))) return CZ_E_MEM;
        } else {
            side = nullptr;
            CZ_KL(czk::deflate_checksum_kernel<<<nseg < 65535u * 8u ? nseg : 65535u * 8u, 128, 0, st>>>(P));
        }
    }
    const bool search = !(P.tune.level0 || P.tune.huffman_only || P.tune.rle_only);
#ifndef CZ_EXPERIMENTS
    // K1 chains (warp per segment), then K2 match search: persistent CTAs of 1 024 threads sweep contiguous chunks, so the
    // 96 KiB neighbourhood of the sweep (32 KiB of input, 64 KiB of links) stays in the SM's L1. Level 0 / HuffmanOnly / Rle
    // need no chains: the thread-per-position kernel writes their (empty / distance-1) matches.
    (void)split_at;
    P.prevd2 = nullptr;
    if (search) {
        static const int chain_per_sm = [] { const char *e = getenv("CZ_CHAIN_PER_SM"); return e ? atoi(e) : 12; }();
        const unsigned cgrid = nseg < (unsigned)ctx->sm_count * (unsigned)chain_per_sm ? nseg : (unsigned)ctx->sm_count * (unsigned)chain_per_sm;
        CZ_KL(czk::deflate_chain_kernel<<<cgrid, 32, 0, st>>>(P, 0u, nseg));
        if (prof) cudaEventRecord(pr.m0, st);
        const unsigned sms = (unsigned)ctx->sm_count;
        static const long chunk_kb = [] { const char *e = getenv("CZ_MATCH_CHUNK_KB"); return e ? atol(e) : 0l; }();
        uint64_t chunk = chunk_kb > 0 ? (uint64_t)chunk_kb << 10 : L.in_bytes / ((uint64_t)sms * 8);  // at least ~8 chunks per CTA
        chunk = (chunk + 65535) & ~65535ull;
        if (chunk < 65536) chunk = 65536;
        if (chunk > (1u << 20) && chunk_kb <= 0) chunk = 1u << 20;
        CZ_KL(czk::deflate_match_sweep_kernel<1024, 1, 0><<<sms, 1024, 0, st>>>(P, L.in_bytes, (uint32_t)chunk, 0u, 0u));
    } else {
        if (prof) cudaEventRecord(pr.m0, st);
        CZ_KL(czk::deflate_match_kernel<<<(unsigned)((L.in_bytes + 255) / 256 ? (L.in_bytes + 255) / 256 : 1), 256, 0, st>>>(P, L.in_bytes));
    }
#else
    // experiments build: every chain / match-search variant stays selectable (CZ_MATCH_V, CZ_MATCH_TILED, CZ_MATCH_LINKS,
    // CZ_CHAIN_SPLIT); measurements in profiles/r1_notes.md
    static const int match_v = [] { const char *e = getenv("CZ_MATCH_V"); return e ? atoi(e) : 3; }();
    static const bool match_tiled = getenv("CZ_MATCH_TILED") != nullptr, chain_split = getenv("CZ_CHAIN_SPLIT") != nullptr;
    (void)search;
    if (!P.tune.level0 && !P.tune.huffman_only && !P.tune.rle_only) {
        static int chain_per_sm = -1;
        if (chain_per_sm < 0) { const char *e = getenv("CZ_CHAIN_PER_SM"); chain_per_sm = e ? atoi(e) : 12; }
        unsigned grid = nseg < (unsigned)ctx->sm_count * (unsigned)chain_per_sm ? nseg : (unsigned)ctx->sm_count * (unsigned)chain_per_sm;
        // Chains in two parts when the launch is more than two waves of resident warps: the first part is exactly one wave, the
        // second runs on the side stream BESIDE the match search of the first part (its 32-thread CTAs fit next to the match
        // kernel's one CTA per SM), so only one wave of the chain pass is exposed.
        // MEASURED: 377 ms against 294 ms per 4 GiB — the chain warps' 17 KB of shared memory each move the SM's L1 / shared
        // memory split, and the match search loses the L1-resident neighbourhood the sweep geometry is built on. Off unless
        // CZ_CHAIN_SPLIT is set.
        const unsigned wave = (unsigned)ctx->sm_count * (unsigned)chain_per_sm;
        split_at = 0;
        if (side && match_v == 3 && !match_tiled && !match_two_links() && chain_split && nseg >= 2 * wave) split_at = wave;
        if (split_at) {
            CZ_KL(czk::deflate_chain_kernel<<<wave, 32, 0, st>>>(P, 0u, split_at));
            bool ok = CZ_CUDA(cudaEventRecord(side->ev_a, st)) && CZ_CUDA(cudaStreamWaitEvent(side->stream, side->ev_a, 0));
            if (!ok) return CZ_E_MEM;
            unsigned g2 = nseg - split_at < wave ? nseg - split_at : wave;
            CZ_KL(czk::deflate_chain_kernel<<<g2, 32, 0, side->stream>>>(P, split_at, nseg));
            if (!CZ_CUDA(cudaEventRecord(side->ev_b, side->stream))) return CZ_E_MEM;
        } else
        CZ_KL(czk::deflate_chain_kernel<<<grid, 32, 0, st>>>(P, 0u, nseg));
        const bool two_links = match_two_links();
        if (two_links) CZ_KL(czk::deflate_chain2_kernel<<<(unsigned)((L.in_bytes + 255) / 256 ? (L.in_bytes + 255) / 256 : 1), 256, 0, st>>>(P, L.in_bytes));
        else P.prevd2 = nullptr;
    } else P.prevd2 = nullptr;
    if (prof) cudaEventRecord(pr.m0, st);
    // Match search geometry. Default (3): persistent CTAs of 1024 threads sweep contiguous chunks, so the 96 KiB neighbourhood
    // of the sweep stays in L1 (94.7 ms per GiB of Markov text at level 6; 4 / 5: 2 x 512 / 2 x 768 threads per SM: 99.3 / 96.5).
    // 7: the sweep with a flattened walk (one chain step per loop iteration, lanes on different positions: 159 ms);
    // 6: the sweep with the warp-synchronous walk/extend alternation of find_match_warp (139 ms: lanes that found a candidate
    // wait for the slowest walker). CZ_MATCH_V=1: 256-position CTAs dealt round robin (110.4 ms); 2: the candidate-pairs experiment (132.6 ms: it gives up
    // find_match's pruning of candidates that cannot beat the best so far); CZ_MATCH_TILED=1: the tiled experiment.
    if (!(P.tune.level0 || P.tune.huffman_only || P.tune.rle_only) && match_v >= 3 && match_v <= 7 && !match_tiled) {
        const unsigned sms = (unsigned)ctx->sm_count;
        // chunk swept by one CTA: large (the neighbourhood is fetched once per chunk), but at least ~8 chunks per CTA
        static long chunk_kb = -1;
        if (chunk_kb < 0) { const char *e = getenv("CZ_MATCH_CHUNK_KB"); chunk_kb = e ? atol(e) : 0; }
        uint64_t chunk = chunk_kb > 0 ? (uint64_t)chunk_kb << 10 : L.in_bytes / ((uint64_t)sms * 8);
        chunk = (chunk + 65535) & ~65535ull;
        if (chunk < 65536) chunk = 65536;
        if (chunk > (1u << 20) && chunk_kb <= 0) chunk = 1u << 20;
        if (match_v == 3 && split_at) {
            // first part now, second part once its chains are there
            CZ_KL(czk::deflate_match_sweep_kernel<1024, 1, 0><<<sms, 1024, 0, st>>>(P, L.in_bytes, (uint32_t)chunk, 0u, split_at));
            if (!CZ_CUDA(cudaStreamWaitEvent(st, side->ev_b, 0))) return CZ_E_MEM;
            CZ_KL(czk::deflate_match_sweep_kernel<1024, 1, 0><<<sms, 1024, 0, st>>>(P, L.in_bytes, (uint32_t)chunk, split_at, nseg));
        } else if (match_v == 3 && P.prevd2) CZ_KL(czk::deflate_match_sweep_kernel<1024, 1, 0, true><<<sms, 1024, 0, st>>>(P, L.in_bytes, (uint32_t)chunk, 0u, 0u));
        else if (match_v == 3) CZ_KL(czk::deflate_match_sweep_kernel<1024, 1, 0><<<sms, 1024, 0, st>>>(P, L.in_bytes, (uint32_t)chunk, 0u, 0u));
        else if (match_v == 7) CZ_KL(czk::deflate_match_sweep_kernel<1024, 1, 2><<<sms, 1024, 0, st>>>(P, L.in_bytes, (uint32_t)chunk, 0u, 0u));
        else if (match_v == 6) CZ_KL(czk::deflate_match_sweep_kernel<1024, 1, 1><<<sms, 1024, 0, st>>>(P, L.in_bytes, (uint32_t)chunk, 0u, 0u));
        else if (match_v == 4) CZ_KL(czk::deflate_match_sweep_kernel<512, 2, 0><<<sms * 2, 512, 0, st>>>(P, L.in_bytes, (uint32_t)chunk, 0u, 0u));
        else CZ_KL(czk::deflate_match_sweep_kernel<768, 2, 0><<<sms * 2, 768, 0, st>>>(P, L.in_bytes, (uint32_t)chunk, 0u, 0u));
    } else if (!(P.tune.level0 || P.tune.huffman_only || P.tune.rle_only) && match_v == 2 && !match_tiled) {
        CZ_KL(czk::deflate_match_pairs_kernel<<<(unsigned)((L.in_bytes + 255) / 256 ? (L.in_bytes + 255) / 256 : 1), 256, 0, st>>>(P, L.in_bytes));
    } else if (P.tune.level0 || P.tune.huffman_only || P.tune.rle_only || !match_tiled) {
        CZ_KL(czk::deflate_match_kernel<<<(unsigned)((L.in_bytes + 255) / 256 ? (L.in_bytes + 255) / 256 : 1), 256, 0, st>>>(P, L.in_bytes));
    } else {
        static bool configured[64] = {};
        const size_t smem = czk::deflate_match_tiled_smem();
        if (!configured[ctx->dev & 63]) {
            if (!CZ_CUDA(cudaFuncSetAttribute(czk::deflate_match_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return CZ_E_MEM;
            configured[ctx->dev & 63] = true;
        }
        const unsigned tiles = (unsigned)((L.in_bytes >> 12) + L.nseg + 1);
        CZ_KL(czk::deflate_match_tiled_kernel<<<tiles, CZK_MT_THREADS, smem, st>>>(P));
    }
#endif  // CZ_EXPERIMENTS
    {
    if (prof) cudaEventRecord(pr.m1, st);
        // one wave of 64-position sub-tiles when the segments fit 16 warps per SM, else 32-position sub-tiles at 28 per SM
        static int parse_per_sm = -1;
        if (parse_per_sm < 0) { const char *e = getenv("CZ_PARSE_PER_SM"); parse_per_sm = e ? atoi(e) : 0; }
        const unsigned sms = (unsigned)ctx->sm_count;
        const bool small_tiles = parse_per_sm ? parse_per_sm > 16 : nseg > sms * 16u;
        const unsigned per_sm = parse_per_sm ? (unsigned)parse_per_sm : (small_tiles ? 28u : 16u);
        unsigned grid = nseg < sms * per_sm ? nseg : sms * per_sm;
        if (small_tiles) CZ_KL(czk::deflate_parse_kernel<32><<<grid, 32, 0, st>>>(P));
        else CZ_KL(czk::deflate_parse_kernel<64><<<grid, 32, 0, st>>>(P));
    }
    CZ_KL(czk::deflate_hist_kernel<<<nsl, 128, 0, st>>>(P));
    CZ_KL(czk::deflate_plan_kernel<<<(nsl + 31) / 32, 32, 0, st>>>(P));
    CZ_KL(czk::deflate_seg_layout_kernel<<<(nseg + 31) / 32, 32, 0, st>>>(P));
    CZ_KL(czk::deflate_unit_size_kernel<<<(nun + 31) / 32, 32, 0, st>>>(P));
    if (L.packed) CZ_KL(czk::deflate_scan_kernel<<<1, 1024, 0, st>>>(P));
    if (side && !CZ_CUDA(cudaStreamWaitEvent(st, side->join, 0))) return CZ_E_MEM;
    CZ_KL(czk::deflate_unit_frame_kernel<<<(nun + 31) / 32, 32, 0, st>>>(P));
    CZ_KL(czk::deflate_zero_kernel<<<(nsl + 127) / 128, 128, 0, st>>>(P));
    CZ_KL(czk::deflate_emit_kernel<<<nsl, CZK_EMIT_THREADS, 0, st>>>(P));
    if (prof) {
        cudaEventRecord(pr.e1, st);
        std::lock_guard<std::mutex> lk(g_dprof_mu);
        g_dprof.push_back(pr);
    }
    return CZ_CUDA(cudaGetLastError()) ? 0 : CZ_E_MEM;
}

// ------------------------------------------------------------------------------------------------------------------
// Host engine. Compresses n units (host memory) into raw-deflate payloads (segments + flush markers, no container):
// the callers add header / final block / trailer, which is O(1) per unit.
struct UnitResult {
    uint64_t payload_len = 0;
    uint32_t adler = 1, crc = 0;
    int32_t status = CZ_ENCODE_FINISHED;  // FINISHED or NEED_OUTPUT (payload did not fit dst_cap)
};

struct EngineJob {
    const uint8_t *in = nullptr;
    const uint64_t *unit_off = nullptr;  // n+1 offsets into `in`
    size_t n = 0;
    std::vector<uint8_t *> dst;          // payload destination per unit
    std::vector<uint64_t> dst_cap;
    uint64_t seg_bytes = kDefaultSegment;
    int level = 6, strategy = 0;
    int check_kind = 3;
    std::vector<UnitResult> res;
    std::vector<uint64_t> *seg_sizes = nullptr;  // optional: compressed size of every segment, in order
    // plan
    std::vector<uint64_t> seg_off;   // global segment offsets (nseg+1)
    std::vector<uint32_t> seg_unit;  // unit of each segment
    std::vector<uint64_t> unit_written;  // payload bytes already placed per unit
    std::vector<size_t> batch_seg;   // batch b = segments [batch_seg[b], batch_seg[b+1])
};

static uint64_t batch_bytes_limit() {
    static uint64_t v = 0;
    if (!v) {
        // Large enough that the warp-per-segment passes (checksum, chains, parse) fill the machine, small enough that the copies
        // of one batch overlap the kernels of the other. cfg3 end to end on a B200 (kernel chains of consecutive batches
        // serialised, see slot_front): 1 GiB / 2 GiB / 4 GiB batches: 383 / 370 / 418 ms. A 2 GiB batch needs ~17 GiB per
        // pipeline slot (input, output bound, workspace): the default on devices with >= 64 GiB, 1 GiB elsewhere.
        size_t mem_free = 0, mem_total = 0;
        v = 1024ull << 20;
        if (cudaMemGetInfo(&mem_free, &mem_total) == cudaSuccess && mem_total >= (64ull << 30) && mem_free >= (48ull << 30)) v = 2048ull << 20;
        if (const char *e = getenv("CZ_BATCH_MB")) { long m = atol(e); if (m >= 1 && m <= 8192) v = (uint64_t)m << 20; }
    }
    return v;
}

static void plan_job(EngineJob &J) {
    J.seg_off.clear(); J.seg_unit.clear(); J.batch_seg.clear();
    const uint64_t S = J.seg_bytes;
    for (size_t u = 0; u < J.n; u++) {
        uint64_t a = J.unit_off[u], b = J.unit_off[u + 1];
        do {  // an empty unit still gets one (empty) segment so that its payload is a valid flush point
            J.seg_off.push_back(a);
            J.seg_unit.push_back((uint32_t)u);
            a += std::min<uint64_t>(S, b - a);
        } while (a < b);
    }
    J.seg_off.push_back(J.n ? J.unit_off[J.n] : 0);
    const size_t nseg = J.seg_unit.size();
    const uint64_t lim = batch_bytes_limit();
    size_t s = 0;
    J.batch_seg.push_back(0);
    while (s < nseg) {
        uint64_t bytes = 0;
        size_t e = s;
        while (e < nseg && (e == s || bytes + (J.seg_off[e + 1] - J.seg_off[e]) <= lim) && e - s < (1u << 20)) {
            bytes += J.seg_off[e + 1] - J.seg_off[e];
            e++;
        }
        J.batch_seg.push_back(e);
        s = e;
    }
    J.res.assign(J.n, UnitResult());
    J.unit_written.assign(J.n, 0);
    if (J.seg_sizes) J.seg_sizes->assign(nseg, 0);
}

// One pipeline slot of one device: its own stream, device buffers and pinned metadata staging.
struct DeflateSlot {
    DevBuf in, out, ws, meta;
    PinBuf hmeta, hres, stage;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    cudaEvent_t kdone = nullptr;  // the batch's kernels have finished (recorded before the result copies)
    bool kdone_set = false;
    int dev = -1;
    // batch in flight
    bool busy = false;
    size_t b = 0, s0 = 0, s1 = 0, np = 0;
    std::vector<uint32_t> piece_unit;
    std::vector<uint32_t> piece_seg;
    uint64_t res_off_len = 0, res_off_pos = 0, res_off_stat = 0, res_off_chk = 0, res_off_total = 0, res_off_segsz = 0;
    bool init(int d) {
        if (dev == d && stream) return true;
        dev = d;
        return CZ_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking)) &&
               CZ_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming)) &&
               CZ_CUDA(cudaEventCreateWithFlags(&kdone, cudaEventDisableTiming));
    }
    ~DeflateSlot() {
        if (done) cudaEventDestroy(done);
        if (kdone) cudaEventDestroy(kdone);
        if (stream) cudaStreamDestroy(stream);
    }
};

static bool engine_trace() {
    static int v = -1;
    if (v < 0) v = getenv("CZ_TRACE") ? 1 : 0;
    return v != 0;
}
static double trace_ms() {
    static const auto t0 = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
#define CZ_TRACE_PT(what, b) do { if (engine_trace()) fprintf(stderr, "[cz] %9.2f ms  %s batch %zu\n", trace_ms(), what, (size_t)(b)); } while (0)

static int slot_front(DeflateSlot &w, DeviceCtx *ctx, EngineJob &J, size_t b, DeflateSlot *prev) {
    CZ_TRACE_PT("front begin", b);
    const size_t s0 = J.batch_seg[b], s1 = J.batch_seg[b + 1], nseg = s1 - s0;
    w.b = b; w.s0 = s0; w.s1 = s1;
    w.piece_unit.clear(); w.piece_seg.clear();
    for (size_t s = s0; s < s1; s++)
        if (s == s0 || J.seg_unit[s] != J.seg_unit[s - 1]) { w.piece_unit.push_back(J.seg_unit[s]); w.piece_seg.push_back((uint32_t)(s - s0)); }
    w.piece_seg.push_back((uint32_t)nseg);
    const size_t np = w.piece_unit.size();
    w.np = np;
    const uint64_t ib = J.seg_off[s0], ie = J.seg_off[s1], in_bytes = ie - ib;
    // device meta: seg_off[nseg+1] | piece_seg[np+1] | cap_off[np+1] | len[np] | status[np] | checks[2np] | segsz[nseg]
    const uint64_t m_seg = 0, m_pseg = align_up(m_seg + 8 * (nseg + 1), 8), m_cap = align_up(m_pseg + 4 * (np + 1), 8),
                   m_len = m_cap + 8 * (np + 1), m_stat = m_len + 8 * np, m_chk = align_up(m_stat + 4 * np, 8),
                   m_segsz = align_up(m_chk + 8 * np, 8), m_total = m_segsz + 8 * nseg;
    const uint64_t up_bytes = m_len;  // the part uploaded
    if (!w.hmeta.reserve(up_bytes) || !w.meta.reserve(m_total)) return CZ_E_MEM;
    uint8_t *hm = w.hmeta.as<uint8_t>();
    uint64_t *h_seg = (uint64_t *)(hm + m_seg);
    for (size_t i = 0; i <= nseg; i++) h_seg[i] = J.seg_off[s0 + i] - ib;
    memcpy(hm + m_pseg, w.piece_seg.data(), 4 * (np + 1));
    uint64_t *h_cap = (uint64_t *)(hm + m_cap);
    uint64_t out_bound = 0;
    for (size_t p = 0; p < np; p++) {
        h_cap[p] = out_bound;
        for (uint32_t s = w.piece_seg[p]; s < w.piece_seg[p + 1]; s++) out_bound += segment_bound(h_seg[s + 1] - h_seg[s]);
    }
    h_cap[np] = out_bound;
    const WsLayout wl = ws_layout(nseg, np, in_bytes);
    if (!w.in.reserve(in_bytes + 64) || !w.out.reserve(out_bound + 64) || !w.ws.reserve(wl.bytes)) return CZ_E_MEM;
    uint8_t *dm = w.meta.as<uint8_t>();
    cudaStream_t st = w.stream;
    if (!CZ_CUDA(cudaMemcpyAsync(dm, hm, up_bytes, cudaMemcpyHostToDevice, st))) return CZ_E_MEM;
    if (in_bytes && !CZ_CUDA(cudaMemcpyAsync(w.in.p, J.in + ib, in_bytes, cudaMemcpyHostToDevice, st))) return CZ_E_MEM;
    DeflateLaunch L;
    memset(&L, 0, sizeof L);
    uint64_t *d_pos = nullptr, *d_total = nullptr;
    L.d_in = w.in.as<uint8_t>(); L.d_out = w.out.as<uint8_t>();
    L.d_seg_off = (const uint64_t *)(dm + m_seg); L.d_unit_seg = (const uint32_t *)(dm + m_pseg);
    L.d_unit_out_off = (const uint64_t *)(dm + m_cap); L.d_unit_out_len = (uint64_t *)(dm + m_len);
    L.d_unit_status = (int32_t *)(dm + m_stat); L.d_unit_checks = (uint32_t *)(dm + m_chk);
    L.d_seg_out_bytes = J.seg_sizes ? (uint64_t *)(dm + m_segsz) : nullptr;
    L.nseg = nseg; L.n_units = np; L.in_bytes = in_bytes;
    L.level = J.level; L.strategy = J.strategy; L.window_bits = -15; L.piece_mode = 1; L.check_kind = J.check_kind;
    L.packed = true; L.d_unit_out_pos_ret = &d_pos; L.d_total_ret = &d_total;
    // The kernels of this batch start when those of the previous batch are done: the copies still overlap the other batch's
    // kernels, but two kernel chains in flight delay each other — the earlier batch's parse / emit passes queue behind the
    // later batch's match search, and its device-to-host copy with them (traced: the first of two 2 GiB batches finished
    // after 334 ms instead of ~205 ms).
    if (prev && prev->kdone_set && !CZ_CUDA(cudaStreamWaitEvent(st, prev->kdone, 0))) return CZ_E_MEM;
    int rc = launch_deflate(st, ctx, L, w.ws.p, w.ws.cap);
    if (rc) return rc;
    if (!CZ_CUDA(cudaEventRecord(w.kdone, st))) return CZ_E_MEM;
    w.kdone_set = true;
    // results back: len[np] pos[np] status[np] checks[2np] total segsz[nseg]
    w.res_off_len = 0; w.res_off_pos = 8 * np; w.res_off_stat = 16 * np; w.res_off_chk = align_up(20 * np, 8);
    w.res_off_total = w.res_off_chk + 8 * np; w.res_off_segsz = w.res_off_total + 8;
    if (!w.hres.reserve(w.res_off_segsz + 8 * nseg)) return CZ_E_MEM;
    uint8_t *hr = w.hres.as<uint8_t>();
    bool ok = CZ_CUDA(cudaMemcpyAsync(hr + w.res_off_len, dm + m_len, 8 * np, cudaMemcpyDeviceToHost, st)) &&
              CZ_CUDA(cudaMemcpyAsync(hr + w.res_off_pos, d_pos, 8 * np, cudaMemcpyDeviceToHost, st)) &&
              CZ_CUDA(cudaMemcpyAsync(hr + w.res_off_stat, dm + m_stat, 4 * np, cudaMemcpyDeviceToHost, st)) &&
              CZ_CUDA(cudaMemcpyAsync(hr + w.res_off_chk, dm + m_chk, 8 * np, cudaMemcpyDeviceToHost, st)) &&
              CZ_CUDA(cudaMemcpyAsync(hr + w.res_off_total, d_total, 8, cudaMemcpyDeviceToHost, st));
    if (ok && J.seg_sizes) ok = CZ_CUDA(cudaMemcpyAsync(hr + w.res_off_segsz, dm + m_segsz, 8 * nseg, cudaMemcpyDeviceToHost, st));
    if (!ok || !CZ_CUDA(cudaEventRecord(w.done, st))) return CZ_E_MEM;
    w.busy = true;
    CZ_TRACE_PT("front enqueued", b);
    return 0;
}

// Waits for the slot's kernels, then moves the packed payloads to their destinations. Batches of one unit complete in
// order (deflate_engine_one finishes batches in increasing order).
static int slot_back(DeflateSlot &w, EngineJob &J) {
    if (!w.busy) return 0;
    w.busy = false;
    CZ_TRACE_PT("back wait", w.b);
    if (!CZ_CUDA(cudaEventSynchronize(w.done))) return CZ_E_MEM;
    CZ_TRACE_PT("back kernels done", w.b);
    const size_t np = w.np;
    const uint8_t *hr = w.hres.as<uint8_t>();
    const uint64_t *lens = (const uint64_t *)(hr + w.res_off_len), *pos = (const uint64_t *)(hr + w.res_off_pos);
    const int32_t *stat = (const int32_t *)(hr + w.res_off_stat);
    const uint32_t *chk = (const uint32_t *)(hr + w.res_off_chk);
    const uint64_t total = *(const uint64_t *)(hr + w.res_off_total);
    if (J.seg_sizes) memcpy(J.seg_sizes->data() + w.s0, hr + w.res_off_segsz, 8 * (w.s1 - w.s0));
    cudaStream_t st = w.stream;
    // destinations (sequential: pieces of one unit arrive in order)
    std::vector<uint8_t *> dst(np, nullptr);
    for (size_t p = 0; p < np; p++) {
        const uint32_t u = w.piece_unit[p];
        UnitResult &R = J.res[u];
        if (stat[p] != CZ_ENCODE_FINISHED) { set_error("internal: deflate piece did not fit its bound"); return CZ_E_MEM; }
        const uint64_t ulen = (uint64_t)(J.seg_off[w.s0 + w.piece_seg[p + 1]] - J.seg_off[w.s0 + w.piece_seg[p]]);
        R.adler = czk::adler32_combine_u(R.adler, chk[2 * p], ulen);
        R.crc = czk::crc32_combine_u(R.crc, chk[2 * p + 1], ulen);
        if (R.status == CZ_ENCODE_FINISHED && J.unit_written[u] + lens[p] <= J.dst_cap[u]) dst[p] = J.dst[u] + J.unit_written[u];
        else R.status = CZ_ENCODE_NEED_OUTPUT;
        J.unit_written[u] += lens[p];
        R.payload_len = J.unit_written[u];
    }
    // Payloads leave the device packed. A large piece is copied straight into the caller's buffer by the copy engine (no staging
    // of ours when that buffer is pinned); consecutive small pieces — one copy each would cost more in launches than in bytes —
    // leave in one copy to pinned staging and are placed by the host.
    const uint64_t kDirect = 128u << 10;
    const bool few = np <= 16;
    std::vector<uint32_t> staged;  // pieces that leave through the staging buffer
    uint64_t staged_bytes = 0;
    if (!few) {
        if (!w.stage.reserve(total + 64)) return CZ_E_MEM;
    }
    for (size_t p = 0; p < np;) {
        if (few || lens[p] >= kDirect) {
            if (dst[p] && lens[p] && !CZ_CUDA(cudaMemcpyAsync(dst[p], w.out.as<uint8_t>() + pos[p], lens[p], cudaMemcpyDeviceToHost, st))) return CZ_E_MEM;
            p++;
            continue;
        }
        size_t e = p;
        while (e < np && lens[e] < kDirect && (e == p || pos[e] == pos[e - 1] + lens[e - 1])) e++;
        const uint64_t a = pos[p], b = pos[e - 1] + lens[e - 1];
        if (b > a && !CZ_CUDA(cudaMemcpyAsync(w.stage.as<uint8_t>() + a, w.out.as<uint8_t>() + a, b - a, cudaMemcpyDeviceToHost, st))) return CZ_E_MEM;
        for (size_t q = p; q < e; q++) staged.push_back((uint32_t)q);
        staged_bytes += b - a;
        p = e;
    }
    if (!CZ_CUDA(cudaStreamSynchronize(st))) return CZ_E_MEM;
    CZ_TRACE_PT("back payload copied", w.b);
    if (!staged.empty()) {
        const uint8_t *sp = w.stage.as<uint8_t>();
#pragma omp parallel for schedule(static) if (staged_bytes > (1u << 20))
        for (long k = 0; k < (long)staged.size(); k++) {
            const size_t p = staged[k];
            if (dst[p]) memcpy(dst[p], sp + pos[p], lens[p]);
        }
    }
    return 0;
}

struct DeviceSlots {
    std::mutex mu;        // one host thread at a time drives a device's pipeline slots
    DeflateSlot slot[2];
};
// multi-device path: payloads of units that straddle a device boundary are gathered through a pinned buffer that belongs
// to the CALL (borrowed from a bounded per-device pool), so concurrent callers never share it
struct PartBuf { PinBuf buf; };
static DevicePool<PartBuf, 1> g_part_pool;
static std::mutex g_dslots_mu;
static DeviceSlots *g_dslots[64] = {};  // created on first use, kept for the life of the process (buffers are reused)
static DeviceSlots *device_slots(int dev) {
    std::lock_guard<std::mutex> lk(g_dslots_mu);
    if (!g_dslots[dev]) g_dslots[dev] = new (std::nothrow) DeviceSlots();
    return g_dslots[dev];
}

static void parallel_memcpy(uint8_t *dst, const uint8_t *src, uint64_t n) {
    const uint64_t chunk = 4u << 20;
    if (n <= chunk) { memcpy(dst, src, n); return; }
    const long k = (long)((n + chunk - 1) / chunk);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < k; i++) memcpy(dst + (uint64_t)i * chunk, src + (uint64_t)i * chunk, (size_t)std::min<uint64_t>(chunk, n - (uint64_t)i * chunk));
}

// All batches of the job on ONE device: double-buffered so that H2D + kernels of batch b overlap the D2H of b-1.
static int deflate_engine_one(EngineJob &J, int dev) {
    DeviceCtx *ctx = device_ctx(dev);
    DeviceSlots *S = device_slots(dev);
    if (!ctx || !S) return CZ_E_NO_DEVICE;
    std::lock_guard<std::mutex> lk(S->mu);
    if (!CZ_CUDA(cudaSetDevice(dev))) return CZ_E_MEM;
    const size_t nb = J.batch_seg.size() - 1;
    int rc = 0;
    for (int k = 0; k < 2 && !rc; k++) rc = S->slot[k].init(dev) ? 0 : CZ_E_MEM;
    for (size_t b = 0; b < nb && !rc; b++) {
        rc = slot_front(S->slot[b & 1], ctx, J, b, b ? &S->slot[(b - 1) & 1] : nullptr);                  // its previous batch (b-2) was finished last turn
        if (!rc && b >= 1) rc = slot_back(S->slot[(b - 1) & 1], J);  // finish b-1 while b runs
    }
    if (!rc && nb) rc = slot_back(S->slot[(nb - 1) & 1], J);
    for (int k = 0; k < 2; k++) {  // on failure make sure nothing is left in flight
        if (S->slot[k].stream) cudaStreamSynchronize(S->slot[k].stream);
        S->slot[k].busy = false;
    }
    return rc;
}

// Runs the job on the devices of `mask`. Segments are dealt to the devices in contiguous ranges balanced by bytes, with no
// collective. Units wholly inside one device's range are written straight to their destination; a unit
// that straddles a boundary (e.g. one long stream over 8 GPUs) is concatenated on the host from the devices' pieces.
static int deflate_engine(EngineJob &J, uint32_t mask) {
    if (cz_device_count() == 0) { if (!*cz_last_error()) set_error("no usable sm_100 CUDA device"); return CZ_E_NO_DEVICE; }
    if (!mask) mask = 1;
    std::vector<int> devs;
    for (int d = 0; d < 32; d++)
        if (mask >> d & 1) devs.push_back(d);
    for (int d : devs)
        if (!device_ctx(d)) return CZ_E_NO_DEVICE;
    plan_job(J);
    const size_t nseg = J.seg_unit.size();
    if (nseg == 0) return 0;
    int prev = 0;
    cudaGetDevice(&prev);
    const uint64_t total = J.seg_off.back() - J.seg_off.front();
    size_t nd = devs.size();
    if (total < nd * (8ull << 20)) nd = std::max<size_t>(1, total / (8ull << 20));  // tiny jobs: fewer devices
    int rc_all = 0;
    if (nd == 1) {
        rc_all = deflate_engine_one(J, devs[0]);
        cudaSetDevice(prev);
        return rc_all;
    }
    // segment ranges per device
    std::vector<size_t> cut(nd + 1, nseg);
    cut[0] = 0;
    {
        size_t s = 0;
        for (size_t k = 1; k < nd; k++) {
            const uint64_t target = J.seg_off.front() + total * k / nd;
            while (s < nseg && J.seg_off[s] < target) s++;
            cut[k] = std::max(s, cut[k - 1]);
        }
    }
    struct Part {
        EngineJob J;
        std::vector<uint64_t> off;
        std::vector<uint32_t> unit_ids;
        std::vector<uint8_t> whole;
        std::vector<uint64_t> seg_sizes;
        PartBuf *part = nullptr;
        int dev = 0;
        int rc = 0;
        ~Part() { g_part_pool.release(dev, part); }
    };
    std::vector<Part> parts(nd);
    for (size_t k = 0; k < nd; k++) {
        Part &Pk = parts[k];
        const size_t s0 = cut[k], s1 = cut[k + 1];
        if (s0 == s1) continue;
        for (size_t s = s0; s < s1; s++)
            if (s == s0 || J.seg_unit[s] != J.seg_unit[s - 1]) { Pk.off.push_back(J.seg_off[s]); Pk.unit_ids.push_back(J.seg_unit[s]); }
        Pk.off.push_back(J.seg_off[s1]);
        Pk.J.in = J.in; Pk.J.unit_off = Pk.off.data(); Pk.J.n = Pk.unit_ids.size();
        Pk.J.seg_bytes = J.seg_bytes; Pk.J.level = J.level; Pk.J.strategy = J.strategy; Pk.J.check_kind = J.check_kind;
        if (J.seg_sizes) Pk.J.seg_sizes = &Pk.seg_sizes;
    }
    std::vector<std::thread> th;
    for (size_t k = 0; k < nd; k++) {
        th.emplace_back([&, k]() {
            Part &Pk = parts[k];
            if (!Pk.J.n) return;
            const int dev = devs[k];
            DeviceSlots *S = device_slots(dev);
            if (!S) { Pk.rc = CZ_E_MEM; return; }
            cudaSetDevice(dev);
            const size_t n = Pk.J.n;
            Pk.whole.assign(n, 0);
            std::vector<uint64_t> priv_off(n + 1, 0);
            for (size_t i = 0; i < n; i++) {
                const uint32_t u = Pk.unit_ids[i];
                Pk.whole[i] = Pk.off[i] == J.unit_off[u] && Pk.off[i + 1] == J.unit_off[u + 1];
                uint64_t b = 0;
                if (!Pk.whole[i]) {
                    const uint64_t len = Pk.off[i + 1] - Pk.off[i];
                    b = segment_bound(len) + 64 * (len / J.seg_bytes + 1);
                }
                priv_off[i + 1] = priv_off[i] + b;
            }
            Pk.dev = dev;
            if (priv_off[n] && (!(Pk.part = g_part_pool.acquire(dev)) || !Pk.part->buf.reserve(priv_off[n] + 64))) { Pk.rc = CZ_E_MEM; return; }
            Pk.J.dst.resize(n); Pk.J.dst_cap.resize(n);
            for (size_t i = 0; i < n; i++) {
                const uint32_t u = Pk.unit_ids[i];
                if (Pk.whole[i]) { Pk.J.dst[i] = J.dst[u]; Pk.J.dst_cap[i] = J.dst_cap[u]; }
                else { Pk.J.dst[i] = Pk.part->buf.as<uint8_t>() + priv_off[i]; Pk.J.dst_cap[i] = priv_off[i + 1] - priv_off[i]; }
            }
            plan_job(Pk.J);
            Pk.rc = deflate_engine_one(Pk.J, dev);
        });
    }
    for (auto &t : th) t.join();
    size_t seg_cursor = 0;
    for (size_t k = 0; k < nd; k++) {
        Part &Pk = parts[k];
        if (Pk.rc && !rc_all) rc_all = Pk.rc;
        if (rc_all) continue;
        for (size_t i = 0; i < Pk.J.n; i++) {
            const uint32_t u = Pk.unit_ids[i];
            UnitResult &R = J.res[u];
            const UnitResult &r = Pk.J.res[i];
            if (Pk.whole[i]) { R = r; J.unit_written[u] = r.payload_len; continue; }
            const uint64_t ulen = Pk.off[i + 1] - Pk.off[i];
            R.adler = czk::adler32_combine_u(R.adler, r.adler, ulen);
            R.crc = czk::crc32_combine_u(R.crc, r.crc, ulen);
            if (R.status == CZ_ENCODE_FINISHED && r.status == CZ_ENCODE_FINISHED && J.unit_written[u] + r.payload_len <= J.dst_cap[u])
                parallel_memcpy(J.dst[u] + J.unit_written[u], Pk.J.dst[i], r.payload_len);
            else R.status = CZ_ENCODE_NEED_OUTPUT;
            J.unit_written[u] += r.payload_len;
            R.payload_len = J.unit_written[u];
        }
        if (J.seg_sizes)
            for (uint64_t v : Pk.seg_sizes) if (seg_cursor < J.seg_sizes->size()) (*J.seg_sizes)[seg_cursor++] = v;
    }
    cudaSetDevice(prev);
    return rc_all;
}

static inline uint32_t container_hdr_bytes(int wb) { return wb == 15 ? 2 : wb > 15 ? 10 : 0; }
static inline uint32_t container_trl_bytes(int wb) { return wb == 15 ? 4 : wb > 15 ? 8 : 0; }

static size_t write_header(uint8_t *o, int wb, int level) {
    const uint32_t lvl = level < 0 ? 6 : (uint32_t)level;
    if (wb == 15) {
        // CMF = 0x78; FLEVEL by level like zlib; FCHECK makes (CMF<<8 | FLG) % 31 == 0  (RFC 1950)
        const uint32_t flevel = lvl < 2 ? 0 : lvl < 6 ? 1 : lvl == 6 ? 2 : 3;
        uint32_t h = (0x78u << 8) | (flevel << 6);
        h += 31 - (h % 31);
        o[0] = (uint8_t)(h >> 8); o[1] = (uint8_t)h;
        return 2;
    }
    if (wb > 15) {  // RFC 1952: no optional fields, mtime 0, XFL by level like zlib, OS = 3 (Unix)
        const uint8_t g[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, (uint8_t)(lvl == 9 ? 2 : lvl == 1 ? 4 : 0), 3};
        memcpy(o, g, 10);
        return 10;
    }
    return 0;
}
static size_t write_trailer(uint8_t *o, int wb, uint32_t adler, uint32_t crc, uint64_t in_len) {
    if (wb == 15) { o[0] = (uint8_t)(adler >> 24); o[1] = (uint8_t)(adler >> 16); o[2] = (uint8_t)(adler >> 8); o[3] = (uint8_t)adler; return 4; }
    if (wb > 15) {
        for (int i = 0; i < 4; i++) o[i] = (uint8_t)(crc >> (8 * i));
        for (int i = 0; i < 4; i++) o[4 + i] = (uint8_t)((uint32_t)in_len >> (8 * i));  // ISIZE = length mod 2^32
        return 8;
    }
    return 0;
}

static bool valid_wbits_enc(int wb) { return wb == -15 || wb == 15 || wb == 31; }

}  // namespace czh

using namespace czh;

extern "C" int cz_profile_read_deflate(double *ms_match, double *ms_chain) {
    std::lock_guard<std::mutex> lk(g_dprof_mu);
    double a = 0, b = 0;
    int k = 0;
    for (DeflateProf &r : g_dprof) {
        float x = 0, y = 0;
        if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&x, r.m0, r.m1) == cudaSuccess &&
            cudaEventElapsedTime(&y, r.e0, r.e1) == cudaSuccess) { a += x; b += y; k++; }
        cudaEventDestroy(r.e0); cudaEventDestroy(r.m0); cudaEventDestroy(r.m1); cudaEventDestroy(r.e1);
    }
    g_dprof.clear();
    if (ms_match) *ms_match = a;
    if (ms_chain) *ms_chain = b;
    return k;
}

extern "C" uint64_t cz_deflate_max_segment(void) { return CZK_SEG_MAX; }
extern "C" uint64_t cz_deflate_segment_bound(uint64_t n) { return segment_bound(n); }
extern "C" uint64_t cz_deflate_bound(uint64_t len, int window_bits, uint64_t segment_bytes) {
    const uint64_t S = clamp_segment(segment_bytes);
    const uint64_t nseg = len ? (len + S - 1) / S : 1;
    return len + len / 2048 + 64 * nseg + 2 + container_hdr_bytes(window_bits) + container_trl_bytes(window_bits);
}
extern "C" uint64_t cz_deflate_workspace_bytes(size_t n_segments, uint64_t total_in_bytes) {
    return ws_layout(n_segments, n_segments, total_in_bytes).bytes;
}

extern "C" int cz_deflate_segments_device(void *cuda_stream, size_t n, const uint8_t *d_in, const uint64_t *d_in_offsets,
                                          uint64_t total_in_bytes, uint8_t *d_out, const uint64_t *d_out_offsets,
                                          uint64_t *d_out_lens, int32_t *d_statuses, uint32_t *d_checks, int level, int strategy,
                                          void *d_workspace, uint64_t workspace_bytes) {
    int dev = 0;
    if (!CZ_CUDA(cudaGetDevice(&dev))) return CZ_E_NO_DEVICE;
    DeviceCtx *ctx = device_ctx(dev);
    if (!ctx) return CZ_E_NO_DEVICE;
    DeflateLaunch L;
    memset(&L, 0, sizeof L);
    L.d_in = d_in; L.d_out = d_out; L.d_seg_off = d_in_offsets; L.d_unit_seg = nullptr; L.d_unit_out_off = d_out_offsets;
    L.d_unit_out_len = d_out_lens; L.d_unit_status = d_statuses; L.d_unit_checks = d_checks; L.d_seg_out_bytes = nullptr;
    L.nseg = n; L.n_units = n; L.in_bytes = total_in_bytes; L.level = level; L.strategy = strategy; L.window_bits = -15;
    L.piece_mode = 1; L.check_kind = d_checks ? 3 : 0; L.packed = false;
    return launch_deflate((cudaStream_t)cuda_stream, ctx, L, d_workspace, workspace_bytes);
}

// ------------------------------------------------------------------------------------------------------------------
static int deflate_units_host(size_t n, const uint8_t *in, const uint64_t *in_off, uint8_t *out, const uint64_t *out_off,
                              uint64_t *out_lens, int32_t *statuses, int level, int wb, int strategy, uint64_t segment_bytes,
                              uint32_t mask, std::vector<uint64_t> *seg_sizes) {
    if (!valid_wbits_enc(wb)) { set_error("unsupported window_bits %d", wb); return CZ_E_STREAM; }
    if (level < -1 || level > 9 || strategy < 0 || strategy > 4) { set_error("bad level/strategy"); return CZ_E_STREAM; }
    EngineJob J;
    J.in = in; J.unit_off = in_off; J.n = n; J.seg_bytes = clamp_segment(segment_bytes); J.level = level; J.strategy = strategy;
    J.check_kind = wb == 15 ? 1 : wb > 15 ? 2 : 0;
    J.seg_sizes = seg_sizes;
    const uint32_t hb = container_hdr_bytes(wb), tb = container_trl_bytes(wb);
    J.dst.resize(n); J.dst_cap.resize(n);
    for (size_t u = 0; u < n; u++) {
        const uint64_t cap = out_off[u + 1] - out_off[u];
        J.dst[u] = out + out_off[u] + hb;
        J.dst_cap[u] = cap >= hb + 2 + tb ? cap - hb - 2 - tb : 0;
    }
    int rc = deflate_engine(J, mask);
    if (rc) return rc;
    for (size_t u = 0; u < n; u++) {
        const UnitResult &R = J.res[u];
        const uint64_t cap = out_off[u + 1] - out_off[u];
        const uint64_t total = hb + R.payload_len + 2 + tb;
        out_lens[u] = total;
        if (R.status != CZ_ENCODE_FINISHED || total > cap) { statuses[u] = CZ_ENCODE_NEED_OUTPUT; continue; }
        uint8_t *o = out + out_off[u];
        write_header(o, wb, level);
        uint8_t *t = o + hb + R.payload_len;
        t[0] = 0x03; t[1] = 0x00;  // final empty fixed block: BFINAL=1, BTYPE=01, end-of-block
        write_trailer(t + 2, wb, R.adler, R.crc, in_off[u + 1] - in_off[u]);
        statuses[u] = CZ_ENCODE_FINISHED;
    }
    return 0;
}

extern "C" int cz_deflate_batch(size_t n, const uint8_t *in, const uint64_t *in_offsets, uint8_t *out, const uint64_t *out_offsets,
                                uint64_t *out_lens, int32_t *statuses, int level, int window_bits, int strategy,
                                uint64_t segment_bytes, uint32_t devices_mask) {
    if (!n) return 0;
    return deflate_units_host(n, in, in_offsets, out, out_offsets, out_lens, statuses, level, window_bits, strategy, segment_bytes,
                              devices_mask, nullptr);
}

extern "C" int cz_deflate_segmented(const uint8_t *in, uint64_t len, uint8_t *out, uint64_t cap, uint64_t *out_len, int level,
                                    int window_bits, int strategy, uint64_t segment_bytes, uint32_t devices_mask,
                                    uint64_t *seg_index, uint64_t seg_index_cap, uint64_t *n_segments) {
    uint64_t in_off[2] = {0, len}, out_off[2] = {0, cap}, olen = 0;
    int32_t status = 0;
    std::vector<uint64_t> seg_sizes;
    int rc = deflate_units_host(1, in, in_off, out, out_off, &olen, &status, level, window_bits, strategy, segment_bytes, devices_mask,
                                &seg_sizes);
    if (rc) return rc;
    if (out_len) *out_len = olen;
    if (n_segments) *n_segments = seg_sizes.size();
    if (status != CZ_ENCODE_FINISHED) { set_error("output buffer too small: need %llu bytes", (unsigned long long)olen); return CZ_E_BUF; }
    if (seg_index) {
        if (seg_index_cap < seg_sizes.size() + 1) { set_error("segment index too small"); return CZ_E_BUF; }
        uint64_t o = container_hdr_bytes(window_bits);
        for (size_t i = 0; i < seg_sizes.size(); i++) { seg_index[i] = o; o += seg_sizes[i]; }
        seg_index[seg_sizes.size()] = o;  // offset of the final 03 00 block
    }
    return 0;
}

extern "C" int cz_inflate_segmented(const uint8_t *in, uint64_t len, uint8_t *out, uint64_t cap, uint64_t *out_len, int window_bits,
                                    uint64_t segment_bytes, const uint64_t *seg_index, uint64_t n_segments, uint32_t devices_mask) {
    if (!valid_wbits_enc(window_bits)) { set_error("unsupported window_bits %d", window_bits); return CZ_E_STREAM; }
    const uint64_t S = clamp_segment(segment_bytes);
    const uint32_t hb = container_hdr_bytes(window_bits), tb = container_trl_bytes(window_bits);
    if (!seg_index || n_segments == 0 || seg_index[0] != hb || seg_index[n_segments] + 2 + tb != len) {
        set_error("segment index does not match the stream");
        return CZ_E_DATA;
    }
    // container header as written by cz_deflate_segmented
    if (window_bits == 15 && (in[0] != 0x78 || ((in[0] << 8 | in[1]) % 31) != 0 || (in[1] & 0x20))) return CZ_E_DATA;
    if (window_bits > 15 && (in[0] != 0x1f || in[1] != 0x8b || in[2] != 8 || in[3] != 0)) return CZ_E_DATA;
    // On ONE device the block-parallel path decodes the stream faster than a warp per 1 MiB segment does (runs of ~32 KiB of
    // input instead of whole segments: B200, 1 GiB gzip: 114 ms against 133 ms), so it goes first; with several devices the
    // index shards the segments over them, and it remains the fallback.
    if (len >= runs_min_unit_bytes() && __builtin_popcount(devices_mask ? devices_mask : 1u) == 1) {
        const uint64_t ioff[2] = {0, len}, ooff[2] = {0, cap};
        uint64_t got = 0, used = 0;
        int32_t st1 = 0;
        uint8_t done1 = 0;
        int prev = 0;
        cudaGetDevice(&prev);
        const int rc1 = inflate_long_units(devices_mask ? __builtin_ctz(devices_mask) : 0, std::vector<size_t>(1, 0), in, ioff, out, ooff, &got, &st1,
                                           &used, window_bits, &done1);
        cudaSetDevice(prev);
        if (rc1 == 0 && done1 && used == len) {
            if (out_len) *out_len = got;
            return 0;
        }
    }
    std::vector<uint64_t> in_off(seg_index, seg_index + n_segments + 1), out_off(n_segments + 1), lens(n_segments);
    std::vector<int32_t> st(n_segments);
    std::vector<uint32_t> chk(2 * n_segments);
    for (uint64_t i = 0; i <= n_segments; i++) out_off[i] = std::min<uint64_t>(i * S, cap);
    for (uint64_t i = 0; i < n_segments; i++)
        if (in_off[i + 1] < in_off[i] || in_off[i + 1] > len) return CZ_E_DATA;
    int rc = inflate_batch_host(n_segments, in, in_off.data(), out, out_off.data(), lens.data(), st.data(), nullptr, -15, 1, chk.data(),
                                devices_mask);
    if (rc) return rc;
    uint64_t total = 0;
    uint32_t adler = 1, crc = 0;
    for (uint64_t i = 0; i < n_segments; i++) {
        if (st[i] == CZ_DECODE_NEED_OUTPUT) { set_error("output buffer too small"); return CZ_E_BUF; }
        if (st[i] != CZ_DECODE_FINISHED) { set_error("segment %llu: status %d", (unsigned long long)i, st[i]); return st[i] < 0 ? st[i] : CZ_E_DATA; }
        if (i + 1 < n_segments && lens[i] != S) { set_error("segment %llu is not %llu bytes", (unsigned long long)i, (unsigned long long)S); return CZ_E_DATA; }
        adler = czk::adler32_combine_u(adler, chk[2 * i], lens[i]);
        crc = czk::crc32_combine_u(crc, chk[2 * i + 1], lens[i]);
        total += lens[i];
    }
    const uint8_t *t = in + seg_index[n_segments];
    if (t[0] != 0x03 || t[1] != 0x00) return CZ_E_DATA;
    t += 2;
    if (window_bits == 15) {
        const uint32_t want = (uint32_t)t[0] << 24 | (uint32_t)t[1] << 16 | (uint32_t)t[2] << 8 | t[3];
        if (want != adler) { set_error("incorrect data check"); return CZ_E_DATA; }
    } else if (window_bits > 15) {
        const uint32_t want = (uint32_t)t[0] | (uint32_t)t[1] << 8 | (uint32_t)t[2] << 16 | (uint32_t)t[3] << 24;
        const uint32_t isz = (uint32_t)t[4] | (uint32_t)t[5] << 8 | (uint32_t)t[6] << 16 | (uint32_t)t[7] << 24;
        if (want != crc) { set_error("incorrect data check"); return CZ_E_DATA; }
        if (isz != (uint32_t)total) { set_error("incorrect length check"); return CZ_E_DATA; }
    }
    if (out_len) *out_len = total;
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Streaming encoder. The kernels are one-shot per segment, the contract is chunked (Process / Flush / Finish with
// arbitrary input and output sizes). The backend stages: Process copies the caller's slice into a pinned staging buffer
// (the slice is only borrowed for the call); Flush and Finish compress everything staged since the last flush point as
// independent segments (each ends byte-aligned with 00 00 ff ff, which is what a sync flush requires and more) and queue
// the bytes; every call drains the queue into the caller's output. Compressed bytes therefore depend only on the data
// between flush points, never on how the caller chunked it (tests/encoder.rs:56-57, 65-66).
struct EncoderState {
    int level, window_bits, mem_level, strategy;
    PinBuf in;             // staged, not yet compressed input
    size_t in_len = 0;
    PinBuf pend;           // compressed bytes not yet handed to the caller
    size_t pend_len = 0, pend_pos = 0;
    bool header_done = false, finished = false, error = false;
    uint32_t adler = 1, crc = 0;
    uint64_t total_in = 0;
    int last_rank = -4;    // zlib's RANK(last_flush); deflateReset leaves last_flush = -2
    int dev = 0;           // the device this encoder compresses on (cz_set_stream_device at construction)
};

extern "C" void *cz_encoder_new(int level, int window_bits, int mem_level, int strategy) {
    if (!valid_wbits_enc(window_bits) || level < -1 || level > 9 || strategy < 0 || strategy > 4 || mem_level < 1 || mem_level > 9) {
        set_error("unsupported encoder options (level %d, window_bits %d, mem_level %d, strategy %d)", level, window_bits, mem_level, strategy);
        return nullptr;  // deflateInit2_ would return Z_STREAM_ERROR => Interface::zlib_cuda returns None
    }
    if (cz_device_count() == 0) {
        if (!*cz_last_error()) set_error("no usable sm_100 CUDA device");
        return nullptr;  // there is no CPU path
    }
    EncoderState *s = new (std::nothrow) EncoderState();
    if (!s) return nullptr;
    s->level = level; s->window_bits = window_bits; s->mem_level = mem_level; s->strategy = strategy;
    s->dev = stream_device();
    return s;
}

extern "C" void *cz_encoder_reset(void *state) {
    EncoderState *s = (EncoderState *)state;
    if (!s) return nullptr;
    s->in_len = s->pend_len = s->pend_pos = 0;  // keeps the pinned allocations
    s->header_done = s->finished = s->error = false;
    s->adler = 1; s->crc = 0; s->total_in = 0;
    s->last_rank = -4;
    return s;
}

extern "C" void cz_encoder_free(void *state) { delete (EncoderState *)state; }

// Compress the first `nbytes` staged bytes (segments + flush markers) and append them to the pending queue; with nothing to
// compress, `empty_marker` appends the bare marker (what zlib emits for an empty sync flush); `finish` closes the stream.
static int encoder_emit(EncoderState *s, size_t nbytes, bool empty_marker, bool finish) {
    const uint32_t hb = container_hdr_bytes(s->window_bits), tb = container_trl_bytes(s->window_bits);
    const uint64_t S = clamp_segment(0);
    const uint64_t nseg = nbytes ? (nbytes + S - 1) / S : 1;
    const uint64_t need = hb + segment_bound(nbytes) + 64 * nseg + 2 + tb + 16;
    // compact the queue first
    if (s->pend_pos) {
        memmove(s->pend.p, s->pend.as<uint8_t>() + s->pend_pos, s->pend_len - s->pend_pos);
        s->pend_len -= s->pend_pos;
        s->pend_pos = 0;
    }
    if (!s->pend.reserve(s->pend_len + need, true, s->pend_len)) return CZ_E_MEM;
    uint8_t *o = s->pend.as<uint8_t>() + s->pend_len;
    size_t k = 0;
    if (!s->header_done) { k += write_header(o, s->window_bits, s->level); s->header_done = true; }
    if (nbytes) {
        EngineJob J;
        uint64_t off[2] = {0, nbytes};
        J.in = s->in.as<uint8_t>(); J.unit_off = off; J.n = 1; J.seg_bytes = S; J.level = s->level; J.strategy = s->strategy;
        J.check_kind = s->window_bits == 15 ? 1 : s->window_bits > 15 ? 2 : 0;
        J.dst.assign(1, o + k); J.dst_cap.assign(1, need - k);
        int rc = deflate_engine(J, 1u << s->dev);
        if (rc) return rc;
        if (J.res[0].status != CZ_ENCODE_FINISHED) { set_error("internal: staged output bound too small"); return CZ_E_MEM; }
        k += J.res[0].payload_len;
        s->adler = czk::adler32_combine_u(s->adler, J.res[0].adler, nbytes);
        s->crc = czk::crc32_combine_u(s->crc, J.res[0].crc, nbytes);
        s->total_in += nbytes;
        if (nbytes < s->in_len) memmove(s->in.p, s->in.as<uint8_t>() + nbytes, s->in_len - nbytes);
        s->in_len -= nbytes;
    } else if (empty_marker) {
        const uint8_t m[5] = {0x00, 0x00, 0x00, 0xff, 0xff};
        memcpy(o + k, m, 5);
        k += 5;
    }
    if (finish) {
        o[k++] = 0x03; o[k++] = 0x00;
        k += write_trailer(o + k, s->window_bits, s->adler, s->crc, s->total_in);
        s->finished = true;
    }
    s->pend_len += k;
    return 0;
}

// Staged input is compressed in slices of this many bytes as soon as they are complete, so a long Process() sequence
// neither holds the whole input in pinned memory nor leaves all the work to Finish. The slice is a multiple of the segment
// size and slices start at multiples of it, so the segments — and therefore the compressed bytes — are exactly those of a
// one-shot call (tests/encoder.rs:56-57, 65-66: output must not depend on the caller's chunking).
static const size_t kEncoderSlice = 64u << 20;

extern "C" cz_result cz_encode(void *state, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len, int op) {
    EncoderState *s = (EncoderState *)state;
    cz_result r;
    r.input_remain = in_len;
    r.output_remain = out_len;
    r.status = CZ_ENCODE_ERROR;
    if (!s || s->error || op < CZ_OP_PROCESS || op > CZ_OP_FINISH) return r;
    // The calls zlib's deflate() refuses before it touches the stream, in its order (deflate.c: deflate()), mapped like
    // src/encoder/mod.rs:356-369 maps them: Z_STREAM_ERROR -> Error, Z_BUF_ERROR -> NeedOutput, nothing consumed.
    if (s->finished && op != CZ_OP_FINISH) return r;                     // FINISH_STATE && flush != Z_FINISH
    if (out_len == 0) { r.status = CZ_ENCODE_NEED_OUTPUT; return r; }    // avail_out == 0
    {
        const int rank = op == CZ_OP_PROCESS ? 0 : op == CZ_OP_FLUSH ? 4 : 8;  // RANK(Z_NO_FLUSH / Z_SYNC_FLUSH / Z_FINISH)
        const int old_rank = s->last_rank;
        s->last_rank = rank;
        // no pending output, no new input and a flush no stronger than the last one: zlib has nothing to do
        if (s->pend_pos == s->pend_len && in_len == 0 && rank <= old_rank && op != CZ_OP_FINISH) { r.status = CZ_ENCODE_NEED_OUTPUT; return r; }
    }
    if (s->finished && in_len) { r.status = CZ_ENCODE_NEED_OUTPUT; return r; }  // FINISH_STATE && avail_in != 0: Z_BUF_ERROR
    if (in_len) {
        if (!s->in.reserve(s->in_len + in_len + 16, true, s->in_len)) { s->error = true; return r; }
        memcpy(s->in.as<uint8_t>() + s->in_len, in, in_len);
        s->in_len += in_len;
    }
    r.input_remain = 0;  // like zlib, all input is taken into the window/staging (tests/encoder.rs:17-18)
    while (!s->finished && s->in_len >= kEncoderSlice + (op == CZ_OP_PROCESS ? 0 : 1)) {
        if (encoder_emit(s, kEncoderSlice, false, false)) { s->error = true; return r; }
    }
    // (a Flush that finds undelivered output and no new input only drains: the compressed bytes must not depend on the
    //  sizes of the caller's output buffers, tests/encoder.rs:56-57, 65-66)
    if (!s->finished && ((op == CZ_OP_FLUSH && (s->in_len || s->pend_pos == s->pend_len)) || op == CZ_OP_FINISH)) {
        int rc = encoder_emit(s, s->in_len, op == CZ_OP_FLUSH, op == CZ_OP_FINISH);
        if (rc) { s->error = true; return r; }
    }
    size_t avail = s->pend_len - s->pend_pos;
    size_t k = avail < out_len ? avail : out_len;
    if (k) memcpy(out, s->pend.as<uint8_t>() + s->pend_pos, k);
    s->pend_pos += k;
    r.output_remain = out_len - k;
    if (op == CZ_OP_FINISH) r.status = s->pend_pos == s->pend_len ? CZ_ENCODE_FINISHED : CZ_ENCODE_NEED_OUTPUT;
    else r.status = CZ_ENCODE_CONTINUE;
    return r;
}
