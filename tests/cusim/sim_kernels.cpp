// sim_kernels.cpp — runs the product's kernel sources on the CPU emulator (tests only; see cusim.h).
#include "cusim.h"
#include "../../compu_b200/csrc/inflate_kernel.cuh"
#include "../../compu_b200/csrc/inflate_lane_kernel.cuh"
#include "../../compu_b200/csrc/inflate_lc_kernel.cuh"
#include "../../compu_b200/csrc/inflate_two_phase.cuh"
#include "../../compu_b200/csrc/deflate_kernels.cuh"
#include "../../compu_b200/csrc/inflate_runs_host.h"

using namespace czk;

template <int D, int WARPS>
static void run_inflate(InflateParams P, unsigned grid) {
    cusim::launch(grid, WARPS * 32, inflate_smem_bytes<D, WARPS>(), inflate_kernel<D, WARPS>, P);
}

extern "C" int sim_inflate(size_t n, const uint8_t *in, const uint64_t *in_off, uint8_t *out, const uint64_t *out_off,
                           uint64_t *out_lens, int32_t *statuses, uint64_t *in_consumed, uint32_t *checks, int window_bits,
                           int segment_mode, int check_kind, int D, int grid, uint64_t seed) {
    static CrcTables crc;
    static bool crc_init = false;
    if (!crc_init) { init_crc_tables(&crc); crc_init = true; }
    unsigned long long counter = 0;
    InflateParams P;
    memset(&P, 0, sizeof P);
    P.in = in; P.in_off = in_off; P.out = out; P.out_off = out_off; P.out_lens = out_lens; P.statuses = statuses;
    P.in_consumed = in_consumed; P.checks = checks; P.counter = &counter; P.crc = &crc; P.n = (uint32_t)n; P.ids = nullptr; P.count_only = 0; P.serial_only = (seed % 4) == 3;
    P.window_bits = window_bits; P.segment_mode = segment_mode; P.check_kind = check_kind;
    cusim::set_seed(seed);
    switch (D) {
        case 1: run_inflate<1, 2>(P, grid); break;
        case 2: run_inflate<2, 2>(P, grid); break;
        case 4: run_inflate<4, 2>(P, grid); break;
        case 8: run_inflate<8, 1>(P, grid); break;
        case 32: run_inflate<32, 1>(P, grid); break;
        case -1: cusim::launch(grid, 2 * 32, inflate_lc_smem_bytes<2>(), inflate_lc_kernel<2>, P); break;
        case -2:
        case -3:
        case -4:
        case -5:
        case -6: {
            const uint64_t total_out = out_off[n] - out_off[0];
            std::vector<uint32_t> tok(total_out + 8 * n + 64, 0xDEADBEEFu);
            std::vector<TokMeta> meta(n);
            unsigned long long counter_b = 0, counter_c = 0;
            TwoPhaseParams Q;
            memset(&Q, 0, sizeof Q);
            Q.base = P; Q.tok = tok.data(); Q.meta = meta.data(); Q.counter_b = &counter_b; Q.count_only = 0;
            Q.counter_c = &counter_c; Q.cta_tile = D == -4 ? CZK_LZ_TILE : 0; Q.spin_ns = 0;
            Q.lane_step = seed % 3 == 0 ? 4 : seed % 3 == 1 ? 32 : 0;  // (the sparse-lane launches of small batches, too)
            cusim::launch(grid, 2 * 32, inflate_tok_smem_bytes<2>(), inflate_tok_kernel<2>, Q);
            if (D == -4) {
                cusim::launch(grid, 8 * 32, inflate_lz_cta_smem_bytes<8>(), inflate_lz_cta_kernel<8, 4>, Q);
                cusim::launch(grid, 2 * 32, 0, inflate_lz_kernel<2, 0>, Q);
            } else if (D == -5) cusim::launch(grid, 2 * 32, 0, inflate_lzw_kernel<2, 2, 12, 1>, Q);
            else if (D == -6) cusim::launch(grid, 2 * 32, 0, inflate_lzw_kernel<2, 4, 8, 1>, Q);
            else if (D == -3) cusim::launch(grid, 2 * 32, 0, inflate_lz_kernel<2, 0>, Q);
            else if (seed % 3 == 0) cusim::launch(grid, 2 * 32, 0, inflate_lz_kernel<2, 4>, Q);
            else if (seed % 3 == 1) cusim::launch(grid, 2 * 32, 0, inflate_lz_kernel<2, 2>, Q);
            else cusim::launch(grid, 2 * 32, 0, inflate_lz_kernel<2, 0>, Q);
            for (size_t i = total_out + 8 * n; i < tok.size(); i++) if (tok[i] != 0xDEADBEEFu) return -2;  // token area overrun
            break;
        }
        case -9: cusim::launch(grid, 2 * 32, inflate_lane_smem_bytes<9, 8, 2>(), inflate_lane_kernel<9, 8, 2>, P); break;
        case -8: cusim::launch(grid, 1 * 32, inflate_lane_smem_bytes<8, 7, 1>(), inflate_lane_kernel<8, 7, 1>, P); break;
        default: return -1;
    }
    return 0;
}

// One launch of the warp-per-stream kernel in resumable mode (what cz_decode does per call): `in` = staged bytes of the stream
// from the byte that holds state->bit_pos' origin, `slot` = output slot with the history bytes right in front of it.
extern "C" int sim_inflate_resume(const uint8_t *in, uint64_t in_len, uint8_t *slot, uint64_t cap, void *state, int window_bits,
                                  uint64_t *out_len, int32_t *status, uint64_t seed) {
    static CrcTables crc;
    static bool crc_init = false;
    if (!crc_init) { init_crc_tables(&crc); crc_init = true; }
    unsigned long long counter = 0;
    const uint64_t in_off[2] = {0, in_len}, out_off[2] = {0, cap};
    InflateParams P;
    memset(&P, 0, sizeof P);
    P.in = in; P.in_off = in_off; P.out = slot; P.out_off = out_off; P.out_lens = out_len; P.statuses = status;
    P.counter = &counter; P.crc = &crc; P.n = 1; P.serial_only = 1; P.window_bits = window_bits;
    P.resume = (ResumeState *)state;
    cusim::set_seed(seed);
    run_inflate<1, 2>(P, 1);
    return 0;
}
extern "C" uint64_t sim_resume_state_bytes() { return sizeof(ResumeState); }

// The block-parallel path for long streams (inflate_runs.cuh) with the product's own host orchestration
// (inflate_runs_host.h) on an emulator backend: "device" memory is host memory, launches are cusim launches.
struct SimRunsBackend {
    uint8_t *in, *out;
    std::vector<uint8_t> arena;
    size_t used = 0;
    CrcTables crc_tab;
    unsigned grid;
    uint64_t seed;
    SimRunsBackend() { init_crc_tables(&crc_tab); arena.resize(64u << 20); }
    uint8_t *d_in() { return in; }
    uint8_t *d_out() { return out; }
    void scratch_reset() { used = 0; }
    bool scratch_need(size_t total) { if (total > arena.size()) arena.resize(total); return true; }
    bool out_need(size_t bytes) { out_store.assign(bytes, 0xEE); out = out_store.data(); return true; }
    std::vector<uint8_t> out_store, tok_store;
    void *tok_buffer(size_t bytes) { tok_store.assign(bytes + 64, 0xDB); return tok_store.data(); }
    void *scratch(size_t bytes) {
        const size_t a = (used + 255) & ~(size_t)255;
        if (a + bytes + 256 > arena.size()) return nullptr;
        used = a + bytes;
        memset(arena.data() + a, 0xCD, bytes);  // scratch is NOT zeroed on the device either
        return arena.data() + a;
    }
    bool h2d(void *d, const void *h, size_t n) { memcpy(d, h, n); return true; }
    bool d2h(void *h, const void *d, size_t n) { memcpy(h, d, n); return true; }
    bool zero(void *d, size_t n) { memset(d, 0, n); return true; }
    const CrcTables *crc() { return &crc_tab; }
    void mark(const char *) {}
    bool candidates(const CandChunk *c, uint32_t n, uint64_t *cand) {
        cusim::set_seed(seed++);
        cusim::launch((n + 3) / 4, 128, 0, inflate_candidates_kernel, (const uint8_t *)in, c, n, cand);
        return true;
    }
    bool tok(const TwoPhaseParams &Q) {
        cusim::set_seed(seed++);
        const char *force = getenv("CUSIM_TOKW");  // "1": always the warp-per-run kernel, "0": never (tests pin both)
        if (!Q.count_only && (force ? force[0] == '1' : seed % 4 < 2)) {  // (default: half of the launches)
            cusim::launch(grid, 2 * 32, 0, inflate_tokw_kernel<2>, Q);
            return true;
        }
        TwoPhaseParams QL = Q;
        QL.lane_step = seed % 3 == 0 ? 8 : seed % 3 == 1 ? 32 : 0;
        if (Q.count_only) cusim::launch(grid, 2 * 32, inflate_tok_smem_bytes<2>(), inflate_tok_kernel<2, false>, QL);
        else cusim::launch(grid, 2 * 32, inflate_tok_smem_bytes<2>(), inflate_tok_kernel<2, true>, QL);
        return true;
    }
    bool lz16(const TwoPhaseParams &Q, uint16_t *sym) {
        cusim::set_seed(seed++);
        cusim::launch(grid, 2 * 32, 0, inflate_lz16_kernel<2>, Q, sym);
        return true;
    }
    bool tail_markers(const uint64_t *run_off, uint32_t n, const uint16_t *sym, uint8_t *flags) {
        cusim::set_seed(seed++);
        cusim::launch((n + 1) / 2, 2 * 32, 0, inflate_tail_markers_kernel<2>, run_off, n, sym, flags);
        return true;
    }
    bool window(const RunStream *st, uint32_t ns, const uint64_t *run_off, const uint16_t *sym, uint8_t *win, uint32_t *bad) {
        if (!ns) return true;
        cusim::set_seed(seed++);
        cusim::launch(ns, 1024, CZK_WINDOW_SMEM, inflate_window_kernel, st, run_off, sym, win, bad);
        return true;
    }
    bool resolve(const RunSlice *sl, uint32_t nsl, const uint64_t *run_off, const uint64_t *final_off, const uint8_t *first,
                 const uint16_t *sym, const uint8_t *win, uint8_t *o, uint32_t *bad) {
        if (!nsl) return true;
        cusim::set_seed(seed++);
        cusim::launch(nsl, 256, 0, inflate_resolve_kernel, sl, nsl, run_off, final_off, first, sym, win, o, bad);
        return true;
    }
    bool check(const uint64_t *run_off, const uint64_t *final_off, uint32_t n, const uint8_t *o, int kind, uint32_t *checks) {
        cusim::set_seed(seed++);
        cusim::launch((n + 1) / 2, 2 * 32, 0, inflate_run_check_kernel<2>, run_off, final_off, n, o, (const CrcTables *)&crc_tab, kind, checks);
        return true;
    }
};

// n long streams, packed like sim_inflate. ok[i] = 1: decoded by the parallel path (out_lens / consumed valid); 0: left to the
// serial path. n_runs[i] = runs the stream was cut into.
extern "C" int sim_inflate_runs(size_t n, uint8_t *in, const uint64_t *in_off, uint8_t *out, const uint64_t *out_off, uint64_t *out_lens,
                                uint64_t *consumed, uint8_t *ok, uint32_t *n_runs, int window_bits, uint64_t chunk_bytes, int grid,
                                uint64_t seed) {
    static SimRunsBackend bk;
    bk.in = in; bk.out = nullptr; bk.grid = (unsigned)grid; bk.seed = seed;
    std::vector<czh::BigUnit> units(n);
    for (size_t i = 0; i < n; i++) {
        units[i].h_in = in + in_off[i]; units[i].in_len = in_off[i + 1] - in_off[i]; units[i].d_in_lo = in_off[i];
        units[i].out_cap = out_off[i + 1] - out_off[i]; units[i].window_bits = window_bits;
        units[i].d_out_off = out_off[i] - out_off[0] + 32 * i;  // the caller's layout (plus a guard gap the emulator can check)
    }
    int rc = czh::inflate_runs_batch(bk, units, chunk_bytes);
    for (size_t i = 0; i < n; i++) {
        ok[i] = units[i].ok; out_lens[i] = units[i].out_len; consumed[i] = units[i].in_consumed; n_runs[i] = units[i].n_runs;
        if (units[i].ok) {
            memcpy(out + out_off[i], bk.out + units[i].d_out_off, units[i].out_len);
            // nothing may be written between the units of the device output buffer
            for (uint64_t k = units[i].out_len; k < units[i].out_len + 16; k++) if (bk.out[units[i].d_out_off + k] != 0xEE) return -3;
        }
    }
    return rc;
}

extern "C" uint32_t sim_crc32_combine(uint32_t a, uint32_t b, uint64_t len2) { return crc32_combine_u(a, b, len2); }
extern "C" uint32_t sim_adler32_combine(uint32_t a, uint32_t b, uint64_t len2) { return adler32_combine_u(a, b, len2); }

// The encoder's kernel chain in the order compu_b200/csrc/deflate.cu launches it, over host memory.
// units: n_units with unit_seg[n_units+1] (null => unit == segment). window_bits/piece_mode as in DeflateParams.
extern "C" int sim_deflate(size_t nseg, size_t n_units, const uint8_t *in, const uint64_t *seg_off, const uint32_t *unit_seg,
                           uint8_t *out, const uint64_t *unit_out_off, uint64_t *unit_out_len, int32_t *unit_status,
                           uint32_t *unit_checks, uint64_t *seg_out_bytes, int level, int strategy, int window_bits,
                           int piece_mode, int packed, uint64_t *unit_out_pos, uint64_t *total_out, uint64_t seed) {
    static CrcTables crc;
    static bool crc_init = false;
    if (!crc_init) { init_crc_tables(&crc); crc_init = true; }
    const uint64_t in_bytes = seg_off[nseg] - seg_off[0];
    const uint64_t n_slots = (in_bytes >> 14) + nseg + 1;
    std::vector<SegState> st(nseg);
    std::vector<uint16_t> prevd(in_bytes + 8), prevd2(in_bytes + 8);
    std::vector<uint32_t> match(in_bytes + 8), blk_end(n_slots), freqs(n_slots * CZK_FREQ_STRIDE);
    std::vector<BlockPlan> plans(n_slots);
    DeflateParams P;
    memset(&P, 0, sizeof P);
    P.in = in; P.out = out; P.seg_off = seg_off; P.nseg = (uint32_t)nseg; P.n_units = (uint32_t)n_units; P.n_slots = (uint32_t)n_slots;
    P.unit_seg = unit_seg; P.unit_out_off = unit_out_off; P.unit_out_pos = packed ? unit_out_pos : nullptr;
    P.total_out = packed ? total_out : nullptr; P.unit_out_len = unit_out_len; P.unit_status = unit_status;
    P.unit_checks = unit_checks; P.seg_out_bytes = seg_out_bytes; P.st = st.data(); P.prevd = prevd.data(); P.prevd2 = (seed & 4) ? nullptr : prevd2.data(); P.match = match.data();
    P.blk_end = blk_end.data(); P.freqs = freqs.data(); P.plans = plans.data(); P.crc = &crc;
    P.tune = deflate_tuning(level, strategy); P.window_bits = window_bits; P.level = level; P.piece_mode = piece_mode;
    P.check_kind = unit_checks ? 3 : 0;
    cusim::set_seed(seed);
    const unsigned ns = P.nseg, nu = P.n_units, nsl = P.n_slots;
    if (P.check_kind) cusim::launch(ns < 3 ? ns : 3, 128, 0, deflate_checksum_kernel, P);
    if (!P.tune.level0 && !P.tune.huffman_only && !P.tune.rle_only) {
        cusim::launch(ns < 3 ? ns : 3, 32, 0, deflate_chain_kernel, P, 0u, ns);
        if (P.prevd2) cusim::launch((unsigned)((in_bytes + 255) / 256 ? (in_bytes + 255) / 256 : 1), 256, 0, deflate_chain2_kernel, P, in_bytes);
    }
    if (!(P.tune.level0 || P.tune.huffman_only || P.tune.rle_only) && (seed % 5) == 0)
        {
        if (seed % 3 == 2) cusim::launch(3, 128, 0, deflate_match_sweep_kernel<128, 1, 2>, P, in_bytes, 65536u, 0u, 0u);
        else if (seed % 2) cusim::launch(3, 128, 0, deflate_match_sweep_kernel<128, 1, 1>, P, in_bytes, 65536u, 0u, 0u);
        else if (P.prevd2) cusim::launch(3, 128, 0, deflate_match_sweep_kernel<128, 1, 0, true>, P, in_bytes, 65536u, 0u, 0u);
        else cusim::launch(3, 128, 0, deflate_match_sweep_kernel<128, 1, 0, false>, P, in_bytes, 65536u, 0u, 0u);
    }
    else if (!(P.tune.level0 || P.tune.huffman_only || P.tune.rle_only) && (seed % 3) == 0)
        cusim::launch((unsigned)((in_bytes + 255) / 256 ? (in_bytes + 255) / 256 : 1), 256, 0, deflate_match_pairs_kernel, P, in_bytes);
    else if (P.tune.level0 || P.tune.huffman_only || P.tune.rle_only || (seed & 2))
        cusim::launch((unsigned)((in_bytes + 255) / 256 ? (in_bytes + 255) / 256 : 1), 256, 0, deflate_match_kernel, P, in_bytes);
    else
        cusim::launch((unsigned)((in_bytes >> 12) + nseg + 1), CZK_MT_THREADS, deflate_match_tiled_smem(), deflate_match_tiled_kernel, P);
    if (seed & 1) cusim::launch(ns < 5 ? ns : 5, 32, 0, deflate_parse_kernel<32>, P);
    else cusim::launch(ns < 5 ? ns : 5, 32, 0, deflate_parse_kernel<64>, P);
    cusim::launch(nsl, 128, 0, deflate_hist_kernel, P);
    cusim::launch((nsl + 31) / 32, 32, 0, deflate_plan_kernel, P);
    cusim::launch((ns + 31) / 32, 32, 0, deflate_seg_layout_kernel, P);
    cusim::launch((nu + 31) / 32, 32, 0, deflate_unit_size_kernel, P);
    if (packed) cusim::launch(1, 1024, 0, deflate_scan_kernel, P);
    cusim::launch((nu + 31) / 32, 32, 0, deflate_unit_frame_kernel, P);
    cusim::launch((nsl + 127) / 128, 128, 0, deflate_zero_kernel, P);
    cusim::launch(nsl, CZK_EMIT_THREADS, 0, deflate_emit_kernel, P);
    return 0;
}
