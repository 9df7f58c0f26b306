// inflate.cu — instantiates the inflate kernel for sm_100a and exposes the device-memory C ABI
// (cz_inflate_batch_device / cz_inflate_segments_device, include/compu_b200.h).
#include <stdlib.h>

#include <algorithm>
#include <thread>
#include <chrono>

#include "host_common.h"
// The default build holds the kernels the product runs: inflate_tok_kernel + inflate_lz_kernel (two-phase path),
// inflate_kernel<1, 8> (warp per stream: big units the block-parallel path declines, the resumable streaming decoder) and the
// run kernels of inflate_runs.cuh. The variants that measured slower (profiles/r1_notes.md) are only compiled with
// -DCZ_EXPERIMENTS (make EXPERIMENTS=1); cz_has_experiments() says which build is loaded.
#include "inflate_kernel.cuh"
#ifdef CZ_EXPERIMENTS
#include "inflate_lane_kernel.cuh"
#endif
#include "inflate_lc_kernel.cuh"
#include "inflate_two_phase.cuh"
#include "inflate_runs_host.h"

namespace czh {

struct InflateCfg {
    int D, W;
};

template <int D, int W>
static int launch_cfg(cudaStream_t st, DeviceCtx *ctx, const czk::InflateParams &P) {
    auto kern = czk::inflate_kernel<D, W>;
    const size_t smem = czk::inflate_smem_bytes<D, W>();
    static bool configured[64] = {};
    if (!configured[ctx->dev & 63]) {
        if (!CZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return CZ_E_MEM;
        configured[ctx->dev & 63] = true;
    }
    int per_sm = 0;
    if (!CZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, W * 32, smem))) return CZ_E_MEM;
    if (per_sm < 1) { set_error("inflate kernel <%d,%d> does not fit on an SM", D, W); return CZ_E_MEM; }
    // persistent warps: one resident wave; slots pull units from the global counter
    uint64_t slots_needed = (P.n + D - 1) / D;               // warps
    uint64_t ctas_needed = (slots_needed + W - 1) / W;
    uint64_t grid = (uint64_t)ctx->sm_count * per_sm;
    if (ctas_needed < grid) grid = ctas_needed;
    if (grid < 1) grid = 1;
    CZ_KL(kern<<<(unsigned)grid, W * 32, smem, st>>>(P));
    return CZ_CUDA(cudaGetLastError()) ? 0 : CZ_E_MEM;
}

#ifdef CZ_EXPERIMENTS
// "one lane = one stream" variant: LB/DB = log2 of the primary litlen / distance table sizes, W warps per CTA (1 CTA per SM)
template <int LB, int DB, int W>
static int launch_lane(cudaStream_t st, DeviceCtx *ctx, const czk::InflateParams &P) {
    auto kern = czk::inflate_lane_kernel<LB, DB, W>;
    const size_t smem = czk::inflate_lane_smem_bytes<LB, DB, W>();
    static bool configured[64] = {};
    if (!configured[ctx->dev & 63]) {
        if (!CZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return CZ_E_MEM;
        configured[ctx->dev & 63] = true;
    }
    int per_sm = 0;
    if (!CZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, W * 32, smem))) return CZ_E_MEM;
    if (per_sm < 1) { set_error("inflate lane kernel <%d,%d,%d> does not fit on an SM", LB, DB, W); return CZ_E_MEM; }
    uint64_t ctas_needed = (P.n + 32 * W - 1) / (32 * W);
    uint64_t grid = (uint64_t)ctx->sm_count * per_sm;
    if (ctas_needed < grid) grid = ctas_needed;
    if (grid < 1) grid = 1;
    CZ_KL(kern<<<(unsigned)grid, W * 32, smem, st>>>(P));
    return CZ_CUDA(cudaGetLastError()) ? 0 : CZ_E_MEM;
}

// "one lane = one stream, canonical decode" variant: W warps per CTA, one CTA per SM
template <int W>
static int launch_lc(cudaStream_t st, DeviceCtx *ctx, const czk::InflateParams &P) {
    auto kern = czk::inflate_lc_kernel<W>;
    const size_t smem = czk::inflate_lc_smem_bytes<W>();
    static bool configured[64] = {};
    if (!configured[ctx->dev & 63]) {
        if (!CZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return CZ_E_MEM;
        configured[ctx->dev & 63] = true;
    }
    int per_sm = 0;
    if (!CZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, W * 32, smem))) return CZ_E_MEM;
    if (per_sm < 1) { set_error("inflate lc kernel <%d> does not fit on an SM", W); return CZ_E_MEM; }
    uint64_t ctas_needed = (P.n + 32 * W - 1) / (32 * W);
    uint64_t grid = (uint64_t)ctx->sm_count * per_sm;
    if (ctas_needed < grid) grid = ctas_needed;
    if (grid < 1) grid = 1;
    CZ_KL(kern<<<(unsigned)grid, W * 32, smem, st>>>(P));
    return CZ_CUDA(cudaGetLastError()) ? 0 : CZ_E_MEM;
}

#endif  // CZ_EXPERIMENTS

// Two-phase path (default): phase A = lane-per-stream Huffman decode into tokens, phase B = warp-per-stream LZ77 resolution.
// Workspace: [counter A 128 B | counter B 128 B] [TokMeta n] [token words total_out + 8 n]
static inline uint64_t two_phase_meta_off() { return 256; }
static inline uint64_t two_phase_tok_off(size_t n) { return align_up(256 + sizeof(czk::TokMeta) * (uint64_t)n, 256); }
uint64_t inflate_workspace_bytes(size_t n, uint64_t total_out_bytes) {
    return two_phase_tok_off(n) + 4 * (total_out_bytes + 8 * (uint64_t)n) + 256;
}

// Optional per-kernel timing of the two-phase path (bench.py: live CUDA-event durations of each kernel inside the timed region)
static std::atomic<bool> g_prof_on{false};
bool profiling_on() { return g_prof_on.load(std::memory_order_relaxed); }
struct ProfRec { cudaEvent_t e0, e1, e2; };
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;

static int g_lz_cta = -1, g_lz_spin = -1;

// Launches of phase A with few units (sub-batches of a handful of streams, the runs of one long stream): only every lane_step-th
// lane of a warp takes a unit. A warp issues every path that any of its lanes takes — literal, length, distance, both refills —
// so a lane that shares its warp decodes at the pace of all of them together; alone in its warp it is 2-3 x faster. The step is
// the largest power of two that still fits all units into the lanes resident at 4 warps per CTA, 3 CTAs per SM.
static uint32_t tok_lane_step(const DeviceCtx *ctx, uint64_t n) {
    const uint64_t lanes = (uint64_t)ctx->sm_count * 3 * 4 * 32;
    uint32_t step = 32;
    while (step > 1 && n * step > lanes) step >>= 1;
    return step;
}

template <int WA, int WB, int H>
static int launch_two_phase(cudaStream_t st, DeviceCtx *ctx, const czk::InflateParams &P, void *d_ws, uint64_t ws_bytes,
                            uint64_t total_out_bytes, size_t n_span, bool alone = false) {
    if (ws_bytes < inflate_workspace_bytes(n_span, total_out_bytes)) {
        set_error("inflate workspace too small: %llu < %llu (see cz_inflate_workspace_bytes)", (unsigned long long)ws_bytes,
                  (unsigned long long)inflate_workspace_bytes(n_span, total_out_bytes));
        return CZ_E_MEM;
    }
    czk::TwoPhaseParams Q;
    memset(&Q, 0, sizeof Q);
    Q.base = P;
    Q.count_only = 0;
    Q.counter_b = (unsigned long long *)((uint8_t *)d_ws + 128);
    Q.meta = (czk::TokMeta *)((uint8_t *)d_ws + two_phase_meta_off());
    Q.tok = (uint32_t *)((uint8_t *)d_ws + two_phase_tok_off(n_span));
    Q.counter_c = (unsigned long long *)((uint8_t *)d_ws + 192);
#ifdef CZ_EXPERIMENTS
    // phase B of units whose output slot fits the shared-memory tile runs one CTA per unit (CZ_LZ_CTA=0: off)
    if (g_lz_cta < 0) { const char *e = getenv("CZ_LZ_CTA"); g_lz_cta = e ? atoi(e) : 0; }
    if (g_lz_spin < 0) { const char *e = getenv("CZ_LZ_SPIN_NS"); g_lz_spin = e ? atoi(e) : 100; }
    Q.cta_tile = (g_lz_cta == 1 || g_lz_cta == 2) ? CZK_LZ_TILE : 0;
    Q.spin_ns = (uint32_t)g_lz_spin;
    constexpr int WC = 8;
    auto kc = czk::inflate_lz_cta_kernel<WC, 32 / WC>;
    auto kc4 = czk::inflate_lz_cta_kernel<4, 8>;
    const size_t smem_c = czk::inflate_lz_cta_smem_bytes<WC>();
#endif
    auto ka = czk::inflate_tok_kernel<WA>;
    auto kb = czk::inflate_lz_kernel<WB, H>;
    const size_t smem = czk::inflate_tok_smem_bytes<WA>();
    static bool configured[64] = {};
    static int per_sm_a[64], per_sm_b[64];
    const int d = ctx->dev & 63;
#ifdef CZ_EXPERIMENTS
    static int per_sm_c[64];
#endif
    if (!configured[d]) {
#ifdef CZ_EXPERIMENTS
        if (!CZ_CUDA(cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c))) return CZ_E_MEM;
        if (!CZ_CUDA(cudaFuncSetAttribute(kc4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c))) return CZ_E_MEM;
        if (!CZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_c[d], kc, WC * 32, smem_c))) return CZ_E_MEM;
        if (per_sm_c[d] < 1) { set_error("inflate_lz_cta_kernel does not fit on an SM"); return CZ_E_MEM; }
#endif
        if (!CZ_CUDA(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return CZ_E_MEM;
        if (!CZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_a[d], ka, WA * 32, smem))) return CZ_E_MEM;
        if (!CZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_b[d], kb, WB * 32, 0))) return CZ_E_MEM;
        if (per_sm_a[d] < 1 || per_sm_b[d] < 1) { set_error("two-phase inflate kernels do not fit on an SM"); return CZ_E_MEM; }
        configured[d] = true;
    }
    uint64_t ga = (P.n + 32 * WA - 1) / (32 * WA), gmax = (uint64_t)ctx->sm_count * per_sm_a[d];
    if (ga > gmax) ga = gmax;
    ProfRec pr = {nullptr, nullptr, nullptr};
    if (g_prof_on) {
        cudaEventCreate(&pr.e0); cudaEventCreate(&pr.e1); cudaEventCreate(&pr.e2);
        cudaEventRecord(pr.e0, st);
    }
    // A launch that cannot fill the machine with WA warps per SM (a sub-batch of the pipelined host path) is spread over the
    // SMs with 4 warps per CTA instead: a lane decodes a stream ~2x sooner when its warp shares the schedulers with 3 others
    // instead of 13, and the latency of phase A is what delays the first device-to-host copy of the pipeline.
    constexpr int WL = 4;
    if (WA > WL && P.n <= (uint64_t)ctx->sm_count * 32 * WL * 2) {
        auto kl = czk::inflate_tok_kernel<WL>;
        const size_t smem_l = czk::inflate_tok_smem_bytes<WL>();
        static bool conf_l[64] = {};
        if (!conf_l[d]) {
            if (!CZ_CUDA(cudaFuncSetAttribute(kl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l))) return CZ_E_MEM;
            conf_l[d] = true;
        }
        // Dense lanes for the sub-batches of the pipelined host path: several of them are in flight at once, and one unit per
        // warp made cfg2 end to end 136 ms instead of 98 ms. A launch that is the caller's whole (small) batch has the machine to
        // itself: there only every lane_step-th lane takes a unit, like the runs of a long stream.
        Q.lane_step = alone ? tok_lane_step(ctx, P.n) : 0;
        const uint64_t lanes_needed = (uint64_t)P.n * (Q.lane_step ? Q.lane_step : 1);
        CZ_KL(kl<<<(unsigned)((lanes_needed + 32 * WL - 1) / (32 * WL)), WL * 32, smem_l, st>>>(Q));
    } else
    CZ_KL(ka<<<(unsigned)ga, WA * 32, smem, st>>>(Q));
    if (g_prof_on) cudaEventRecord(pr.e1, st);
#ifdef CZ_EXPERIMENTS
    if (Q.cta_tile) {
        uint64_t gc = P.n, gcmax = (uint64_t)ctx->sm_count * per_sm_c[d];
        if (gc > gcmax) gc = gcmax;
        if (g_lz_cta == 2) CZ_KL(kc4<<<(unsigned)gc, 4 * 32, smem_c, st>>>(Q));
        else CZ_KL(kc<<<(unsigned)gc, WC * 32, smem_c, st>>>(Q));
    }
#endif
    uint64_t gb = (P.n + WB - 1) / WB;
    static int lz_cap = -1;  // experiment knob: CTAs of phase B per SM (fewer streams in flight => their windows fit L2)
    if (lz_cap < 0) { const char *e = getenv("CZ_LZ_CTAS_PER_SM"); lz_cap = e ? atoi(e) : 0; }
    gmax = (uint64_t)ctx->sm_count * (lz_cap > 0 && lz_cap < per_sm_b[d] ? lz_cap : per_sm_b[d]);
    if (gb > gmax) gb = gmax;
#ifdef CZ_EXPERIMENTS
    // phase B variants with several tokens per lane (cz_tune_inflate_lz / CZ_LZ_CTA = 3..7)
#define CZ_LZW(mode, TPL, SHORT, MINB)                                                                              \
    if (g_lz_cta == mode) {                                                                                         \
        auto kw = czk::inflate_lzw_kernel<8, TPL, SHORT, MINB>;                                                     \
        static int per_sm_w[64] = {};                                                                               \
        if (!per_sm_w[d] && !CZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_w[d], kw, 256, 0))) return CZ_E_MEM; \
        uint64_t gw = (P.n + 7) / 8, gwmax = (uint64_t)ctx->sm_count * (per_sm_w[d] > MINB ? MINB : per_sm_w[d]);  \
        if (gw > gwmax) gw = gwmax;                                                                                 \
        CZ_KL(kw<<<(unsigned)gw, 256, 0, st>>>(Q));                                                                        \
    } else
    CZ_LZW(3, 2, 12, 3) CZ_LZW(4, 2, 12, 2) CZ_LZW(5, 4, 8, 2) CZ_LZW(6, 2, 8, 4) CZ_LZW(7, 4, 8, 3)
#undef CZ_LZW
#endif
    CZ_KL(kb<<<(unsigned)gb, WB * 32, 0, st>>>(Q));
    if (g_prof_on) {
        cudaEventRecord(pr.e2, st);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof.push_back(pr);
    }
    return CZ_CUDA(cudaGetLastError()) ? 0 : CZ_E_MEM;
}


static int par_decode_off() {
    static const int v = getenv("CZ_PAR_DECODE") ? 0 : 1;
    return v;
}

// ------------------------------------------------------------------------------------------------------------------
// Block-parallel inflate of long streams: the CUDA backend of inflate_runs_host.h and the per-device driver.
struct CudaRunsBackend {
    DeviceCtx *ctx = nullptr;
    cudaStream_t st = nullptr;
    DevBuf in, out, arena, tokbuf;
    size_t used = 0;
    int per_sm_tok = 0, per_sm_lz16 = 0;
    bool configured = false;
    ~CudaRunsBackend() { if (st) cudaStreamDestroy(st); }
    bool init(DeviceCtx *c) {
        ctx = c;
        if (!st && !CZ_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking))) return false;
        if (!configured) {
            auto ka = czk::inflate_tok_kernel<14, true>;
            auto kl = czk::inflate_tok_kernel<4, true>;
            auto kac = czk::inflate_tok_kernel<14, false>;
            auto klc = czk::inflate_tok_kernel<4, false>;
            if (!CZ_CUDA(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)czk::inflate_tok_smem_bytes<14>())) ||
                !CZ_CUDA(cudaFuncSetAttribute(kl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)czk::inflate_tok_smem_bytes<4>())) ||
                !CZ_CUDA(cudaFuncSetAttribute(kac, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)czk::inflate_tok_smem_bytes<14>())) ||
                !CZ_CUDA(cudaFuncSetAttribute(klc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)czk::inflate_tok_smem_bytes<4>())) ||
                !CZ_CUDA(cudaFuncSetAttribute(czk::inflate_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CZK_WINDOW_SMEM)) ||
                !CZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_tok, ka, 14 * 32, czk::inflate_tok_smem_bytes<14>())) ||
                !CZ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_lz16, czk::inflate_lz16_kernel<8>, 256, 0))) return false;
            if (per_sm_tok < 1 || per_sm_lz16 < 1) { set_error("run-mode inflate kernels do not fit on an SM"); return false; }
            configured = true;
        }
        return true;
    }
    uint8_t *d_in() { return in.as<uint8_t>(); }
    uint8_t *d_out() { return out.as<uint8_t>(); }
    void scratch_reset() { used = 0; }
    bool scratch_need(size_t total) { return arena.reserve(total + 4096); }
    bool out_need(size_t bytes) { return out.reserve(bytes); }
    void *tok_buffer(size_t bytes) { return tokbuf.reserve(bytes + 256) ? tokbuf.p : nullptr; }
    void *scratch(size_t bytes) {
        const size_t a = (used + 255) & ~(size_t)255;
        if (a + bytes > arena.cap) { set_error("internal: run scratch arena too small"); return nullptr; }
        used = a + bytes;
        return arena.as<uint8_t>() + a;
    }
    bool h2d(void *d, const void *h, size_t n) { return !n || (CZ_CUDA(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, st)) && CZ_CUDA(cudaStreamSynchronize(st))); }
    bool d2h(void *h, const void *d, size_t n) { return !n || (CZ_CUDA(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, st)) && CZ_CUDA(cudaStreamSynchronize(st))); }
    bool zero(void *d, size_t n) { return CZ_CUDA(cudaMemsetAsync(d, 0, n, st)); }
    const czk::CrcTables *crc() { return ctx->d_crc; }
    bool ok() { return CZ_CUDA(cudaGetLastError()); }
    // CZ_TRACE=1: host wall-clock timeline of the phases (each mark waits for the stream: tracing serialises, timing does not)
    std::chrono::steady_clock::time_point t_last;
    void mark(const char *phase) {
        static const bool tracing = getenv("CZ_TRACE") != nullptr;
        if (!tracing) return;
        cudaStreamSynchronize(st);
        const auto now = std::chrono::steady_clock::now();
        if (strcmp(phase, "start")) fprintf(stderr, "[cz] runs: %-14s %8.2f ms\n", phase, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    }
    bool candidates(const czk::CandChunk *c, uint32_t n, uint64_t *cand) {
        CZ_KL(czk::inflate_candidates_kernel<<<(n + 3) / 4, 128, 0, st>>>(d_in(), c, n, cand));
        return ok();
    }
    bool tok(const czk::TwoPhaseParams &Q) {
        const uint32_t n = Q.base.n;
        // few runs (one or a few long streams): a warp per run with one decoding lane and look-up tables — up to two waves of
        // 16 warps per SM; beyond that the lane-per-run kernel's throughput wins
        static const bool no_warp_runs = getenv("CZ_NO_WARP_RUNS") != nullptr;
        if (Q.runs && !Q.count_only && !no_warp_runs && n <= (uint32_t)ctx->sm_count * 16 * 2) {
            unsigned g = (n + 7) / 8, gmax = (unsigned)ctx->sm_count * 2;
            if (g > gmax) g = gmax;
            CZ_KL(czk::inflate_tokw_kernel<8><<<g, 256, 0, st>>>(Q));
            return ok();
        }
        // few runs: 4 warps per CTA spread over the SMs (a lane decodes sooner when its warp shares the schedulers with 3 others)
        if (n <= (uint32_t)ctx->sm_count * 32 * 4 * 2) {
            czk::TwoPhaseParams QL = Q;
            QL.lane_step = tok_lane_step(ctx, n);
            const unsigned g = (unsigned)(((uint64_t)n * QL.lane_step + 127) / 128);
            if (Q.count_only) CZ_KL(czk::inflate_tok_kernel<4, false><<<g, 128, czk::inflate_tok_smem_bytes<4>(), st>>>(QL));
            else CZ_KL(czk::inflate_tok_kernel<4, true><<<g, 128, czk::inflate_tok_smem_bytes<4>(), st>>>(QL));
        } else {
            uint64_t g = (n + 32 * 14 - 1) / (32 * 14), gmax = (uint64_t)ctx->sm_count * per_sm_tok;
            if (g > gmax) g = gmax;
            if (Q.count_only) CZ_KL(czk::inflate_tok_kernel<14, false><<<(unsigned)g, 14 * 32, czk::inflate_tok_smem_bytes<14>(), st>>>(Q));
            else CZ_KL(czk::inflate_tok_kernel<14, true><<<(unsigned)g, 14 * 32, czk::inflate_tok_smem_bytes<14>(), st>>>(Q));
        }
        return ok();
    }
    bool lz16(const czk::TwoPhaseParams &Q, uint16_t *sym) {
        uint64_t g = (Q.base.n + 7) / 8, gmax = (uint64_t)ctx->sm_count * per_sm_lz16;
        if (g > gmax) g = gmax;
        CZ_KL(czk::inflate_lz16_kernel<8><<<(unsigned)g, 256, 0, st>>>(Q, sym));
        return ok();
    }
    bool tail_markers(const uint64_t *run_off, uint32_t n, const uint16_t *sym, uint8_t *flags) {
        CZ_KL(czk::inflate_tail_markers_kernel<8><<<(n + 7) / 8, 256, 0, st>>>(run_off, n, sym, flags));
        return ok();
    }
    bool window(const czk::RunStream *s, uint32_t ns, const uint64_t *run_off, const uint16_t *sym, uint8_t *win, uint32_t *bad) {
        if (!ns) return true;
        CZ_KL(czk::inflate_window_kernel<<<ns, 1024, CZK_WINDOW_SMEM, st>>>(s, run_off, sym, win, bad));
        return ok();
    }
    bool resolve(const czk::RunSlice *sl, uint32_t nsl, const uint64_t *run_off, const uint64_t *final_off, const uint8_t *first,
                 const uint16_t *sym, const uint8_t *win, uint8_t *o, uint32_t *bad) {
        if (!nsl) return true;
        CZ_KL(czk::inflate_resolve_kernel<<<nsl, 256, 0, st>>>(sl, nsl, run_off, final_off, first, sym, win, o, bad));
        return ok();
    }
    bool check(const uint64_t *run_off, const uint64_t *final_off, uint32_t n, const uint8_t *o, int kind, uint32_t *checks) {
        CZ_KL(czk::inflate_run_check_kernel<8><<<(n + 7) / 8, 256, 0, st>>>(run_off, final_off, n, o, ctx->d_crc, kind, checks));
        return ok();
    }
};

static uint64_t runs_chunk_bytes() {
    static const uint64_t v = [] { const char *e = getenv("CZ_RUN_CHUNK_KB"); long k = e ? atol(e) : 0; return k >= 4 ? (uint64_t)k << 10 : (uint64_t)(16u << 10); }();
    return v;
}
uint64_t runs_min_unit_bytes() { return std::max<uint64_t>(64u << 10, 4 * runs_chunk_bytes()); }

// Decodes the long streams `ids` of a packed host batch on device `dev` with the block-parallel path, in batches bounded by
// device memory. done[i] = 1 for the units it decoded (out, out_lens, statuses = Finished, in_consumed written); the others
// are left untouched for the serial path.
static DevicePool<CudaRunsBackend, 4> g_runs_pool;

// One batch [a, b) of `ids` on backend `bk`: host -> device, the run pipeline, device -> host, results.
static int inflate_long_batch(CudaRunsBackend *bk, const std::vector<size_t> &ids, size_t a, size_t b, const uint8_t *in,
                              const uint64_t *in_off, uint8_t *out, const uint64_t *out_off, uint64_t *out_lens, int32_t *statuses,
                              uint64_t *in_consumed, int window_bits, uint8_t *done) {
    std::vector<BigUnit> units(b - a);
    uint64_t in_bytes = 0;
    for (size_t k = a; k < b; k++) in_bytes += runs_align(in_off[ids[k] + 1] - in_off[ids[k]] + 64);
    if (!bk->in.reserve(in_bytes + 256)) { cudaGetLastError(); return 0; }  // no memory for this batch: the serial path takes these units
    uint64_t o = 0;
    bool copied = true;
    for (size_t k = a; k < b; k++) {
        BigUnit &U = units[k - a];
        const size_t i = ids[k];
        U.h_in = in + in_off[i]; U.in_len = in_off[i + 1] - in_off[i]; U.d_in_lo = o;
        U.out_cap = out_off[i + 1] - out_off[i]; U.window_bits = window_bits;
        U.d_out_off = out_off[i] - out_off[ids[a]];  // the caller's layout, so that neighbouring units leave in one copy
        copied = copied && CZ_CUDA(cudaMemcpyAsync(bk->d_in() + o, U.h_in, U.in_len, cudaMemcpyHostToDevice, bk->st));
        o += runs_align(U.in_len + 64);
    }
    bk->mark("start");
    if (!copied || !CZ_CUDA(cudaStreamSynchronize(bk->st))) return CZ_E_MEM;
    bk->mark("h2d");
    const int rc = inflate_runs_batch(*bk, units, runs_chunk_bytes());
    if (rc != 0) { cudaGetLastError(); return 0; }  // (out of scratch memory etc.: the serial path takes the batch)
    bool ok = true;
    // device -> host: consecutive decoded units whose slots are exactly full and adjacent leave in one copy
    size_t k = a;
    while (k < b) {
        if (!units[k - a].ok || !units[k - a].out_len) { k++; continue; }
        size_t e = k;
        uint64_t bytes = units[k - a].out_len;
        while (e + 1 < b && units[e + 1 - a].ok && units[e - a].out_len == units[e - a].out_cap && ids[e + 1] == ids[e] + 1 &&
               units[e + 1 - a].d_out_off == units[e - a].d_out_off + units[e - a].out_cap) {
            e++;
            bytes = units[e - a].d_out_off + units[e - a].out_len - units[k - a].d_out_off;
        }
        ok = ok && CZ_CUDA(cudaMemcpyAsync(out + out_off[ids[k]], bk->d_out() + units[k - a].d_out_off, bytes, cudaMemcpyDeviceToHost, bk->st));
        k = e + 1;
    }
    if (!ok || !CZ_CUDA(cudaStreamSynchronize(bk->st))) return CZ_E_MEM;
    bk->mark("d2h");
    for (size_t q = a; q < b; q++) {
        const BigUnit &U = units[q - a];
        if (!U.ok) continue;
        const size_t i = ids[q];
        out_lens[i] = U.out_len; statuses[i] = CZ_DECODE_FINISHED;
        if (in_consumed) in_consumed[i] = U.in_consumed;
        done[i] = 1;
    }
    return 0;
}

// Decodes the long streams `ids` of a packed host batch on device `dev` with the block-parallel path, in batches bounded by
// device memory. done[i] = 1 for the units it decoded (out, out_lens, statuses = Finished, in_consumed written); the others
// are left untouched for the serial path. With more than one batch, two host threads drive alternate batches on two
// backends (own stream and buffers), so the copies of one batch overlap the kernels of the other.
int inflate_long_units(int dev, const std::vector<size_t> &ids, const uint8_t *in, const uint64_t *in_off, uint8_t *out,
                       const uint64_t *out_off, uint64_t *out_lens, int32_t *statuses, uint64_t *in_consumed, int window_bits,
                       uint8_t *done) {
    DeviceCtx *ctx = device_ctx(dev);
    if (!ctx) return CZ_E_NO_DEVICE;
    if (ids.empty()) return 0;
    // batches: bounded by input bytes and by (the callers' slots as a proxy for) output bytes
    uint64_t total_in = 0;
    for (size_t i : ids) total_in += in_off[i + 1] - in_off[i];
    const bool small_job = total_in <= (768ull << 20);
    static const uint64_t batch_mb = [] { const char *e = getenv("CZ_RUNS_BATCH_MB"); long v = e ? atol(e) : 0; return v >= 16 && v <= 4096 ? (uint64_t)v : 384ull; }();
    const uint64_t batch_in = small_job ? (1024ull << 20) : (batch_mb << 20), batch_out = small_job ? (2048ull << 20) : ((batch_mb * 8 / 3) << 20);
    std::vector<size_t> cut(1, 0);
    {
        size_t a = 0;
        while (a < ids.size()) {
            size_t b = a;
            uint64_t in_bytes = 0, cap_bytes = 0;
            while (b < ids.size()) {
                const uint64_t il = in_off[ids[b] + 1] - in_off[ids[b]], ol = out_off[ids[b] + 1] - out_off[ids[b]];
                if (b > a && (in_bytes + il > batch_in || cap_bytes + ol > batch_out ||
                              out_off[ids[b] + 1] - out_off[ids[a]] > 2 * batch_out)) break;  // (the device buffer mirrors the callers' span)
                in_bytes += il; cap_bytes += ol;
                b++;
            }
            cut.push_back(b);
            a = b;
        }
    }
    const size_t nb = cut.size() - 1;
    // every backend is sized for the largest batch of the call up front: which backend gets which batch depends on timing, and a
    // buffer that has to grow in the middle of the call costs a cudaFree + cudaMalloc (a device-wide synchronisation)
    uint64_t max_in = 0, max_span = 0;
    for (size_t k = 0; k < nb; k++) {
        uint64_t ib = 0;
        for (size_t q = cut[k]; q < cut[k + 1]; q++) ib += runs_align(in_off[ids[q] + 1] - in_off[ids[q]] + 64);
        max_in = std::max(max_in, ib);
        max_span = std::max(max_span, out_off[ids[cut[k + 1] - 1] + 1] - out_off[ids[cut[k]]]);
    }
    // (2 / 3 / 4 backends measured on the cfg5 shape, 3 batches on one GPU: see profiles/r2_notes.md)
    static const int max_workers = [] { const char *e = getenv("CZ_RUNS_WORKERS"); int v = e ? atoi(e) : 0; return v >= 1 && v <= 4 ? v : 3; }();
    const int nworkers = (int)std::min<size_t>(nb, (size_t)max_workers);
    std::vector<int> rcs(nworkers, 0);
    auto worker = [&](int w) {
        if (!CZ_CUDA(cudaSetDevice(dev))) { rcs[w] = CZ_E_MEM; return; }
        CudaRunsBackend *bk = g_runs_pool.acquire(dev);
        if (!bk) { rcs[w] = CZ_E_MEM; return; }
        if (!bk->init(ctx)) rcs[w] = CZ_E_MEM;
        if (nb > 1 && !(bk->in.reserve(max_in + 256) && bk->out.reserve(max_span + 256) && bk->tokbuf.reserve(10 * max_in + (8u << 20))))
            cudaGetLastError();  // (not fatal: the batches reserve what they need themselves)
        for (size_t k = (size_t)w; k < nb && !rcs[w]; k += (size_t)nworkers)
            rcs[w] = inflate_long_batch(bk, ids, cut[k], cut[k + 1], in, in_off, out, out_off, out_lens, statuses, in_consumed, window_bits, done);
        g_runs_pool.release(dev, bk);
    };
    {
        std::vector<std::thread> th;
        for (int w = 1; w < nworkers; w++) th.emplace_back(worker, w);
        worker(0);
        for (auto &t : th) t.join();
    }
    for (int r : rcs) if (r) return r;
    return 0;
}

// One resumable unit (the streaming Decoder): the warp-per-stream kernel with its ResumeState. d_ws: 256 bytes.
int launch_inflate_resume(cudaStream_t st, DeviceCtx *ctx, const uint8_t *d_in, const uint64_t *d_in_off, uint8_t *d_out,
                          const uint64_t *d_out_off, uint64_t *d_out_lens, int32_t *d_statuses, int window_bits,
                          czk::ResumeState *d_resume, void *d_ws) {
    czk::InflateParams P;
    memset(&P, 0, sizeof P);
    P.in = d_in; P.in_off = d_in_off; P.out = d_out; P.out_off = d_out_off; P.out_lens = d_out_lens; P.statuses = d_statuses;
    P.counter = (unsigned long long *)d_ws; P.crc = ctx->d_crc; P.n = 1; P.window_bits = window_bits; P.serial_only = 1;
    P.resume = d_resume;
    if (!CZ_CUDA(cudaMemsetAsync(d_ws, 0, 256, st))) return CZ_E_MEM;
    return launch_cfg<1, 8>(st, ctx, P);
}

static InflateCfg g_cfg = {0, 0};


static InflateCfg pick_cfg() {
    if (g_cfg.D == 0) {
        InflateCfg c{-2, 14};  // two-phase: lane-per-stream token decode (14 warps per SM), then warp-per-stream LZ77
        if (const char *e = getenv("CZ_INFLATE_CFG")) {
            int d = 0, w = 0;
            if (sscanf(e, "%d,%d", &d, &w) == 2) { c.D = d; c.W = w; }
        }
        g_cfg = c;
    }
    return g_cfg;
}

int launch_inflate(cudaStream_t st, DeviceCtx *ctx, size_t n, const uint8_t *d_in, const uint64_t *d_in_off, uint8_t *d_out,
                   const uint64_t *d_out_off, uint64_t *d_out_lens, int32_t *d_statuses, uint64_t *d_in_consumed,
                   uint32_t *d_checks, int window_bits, int segment_mode, int check_kind, void *d_ws, uint64_t ws_bytes,
                   uint64_t total_out_bytes, const uint32_t *d_ids, size_t n_ids, int big) {
    if (n == 0 || (d_ids && n_ids == 0)) return 0;
    const bool alone = (big & 2) != 0;  // bit 1: this launch is the caller's whole batch (nothing else of the call is in flight)
    big &= 1;
    if (n > 0xfffffff0u) { set_error("too many units in one launch"); return CZ_E_STREAM; }
    if (ws_bytes < 256 || !d_ws) { set_error("inflate workspace too small"); return CZ_E_MEM; }
    if (!(window_bits == -15 || window_bits == 15 || window_bits == 31 || window_bits == 47) && !segment_mode) {
        set_error("unsupported window_bits %d", window_bits);
        return CZ_E_STREAM;
    }
    czk::InflateParams P;
    P.in = d_in; P.in_off = d_in_off; P.out = d_out; P.out_off = d_out_off; P.out_lens = d_out_lens; P.statuses = d_statuses;
    P.in_consumed = d_in_consumed; P.checks = d_checks; P.counter = (unsigned long long *)d_ws; P.crc = ctx->d_crc;
    P.n = (uint32_t)(d_ids ? n_ids : n); P.ids = d_ids;
    P.window_bits = window_bits; P.segment_mode = segment_mode; P.check_kind = check_kind; P.count_only = 0;
    P.serial_only = par_decode_off();
    P.resume = nullptr;
    if (!CZ_CUDA(cudaMemsetAsync(d_ws, 0, 256, st))) return CZ_E_MEM;
    InflateCfg c = pick_cfg();
    // big units (one stream is megabytes): one WARP per stream, a single decoder lane feeding warp-cooperative LZ77 rounds —
    // about 6x the single-stream speed of the lane-per-stream kernels, which only pay off across thousands of streams
    if (big) return launch_cfg<1, 8>(st, ctx, P);
    // default configurations
    if (c.D == 1 && c.W == 8) return launch_cfg<1, 8>(st, ctx, P);
    if (c.D == -2 && (c.W == 14 || c.W == 0)) return launch_two_phase<14, 8, 0>(st, ctx, P, d_ws, ws_bytes, total_out_bytes, n, alone);
#ifdef CZ_EXPERIMENTS
#define CZ_CFG(d, w) if (c.D == d && c.W == w) return launch_cfg<d, w>(st, ctx, P)
    CZ_CFG(2, 8); CZ_CFG(4, 7); CZ_CFG(4, 4); CZ_CFG(8, 7); CZ_CFG(8, 4); CZ_CFG(8, 2); CZ_CFG(16, 3);
    CZ_CFG(16, 1); CZ_CFG(32, 1);
#undef CZ_CFG
    // D = -2: two-phase, W = rounds per iteration of phase B (loads in flight per lane)
    if (c.D == -2 && c.W == 2) return launch_two_phase<14, 8, 2>(st, ctx, P, d_ws, ws_bytes, total_out_bytes, n);
    if (c.D == -2 && c.W == -8) return launch_two_phase<14, 8, -8>(st, ctx, P, d_ws, ws_bytes, total_out_bytes, n);
    if (c.D == -2 && c.W == -12) return launch_two_phase<14, 8, -12>(st, ctx, P, d_ws, ws_bytes, total_out_bytes, n);
    if (c.D == -2 && c.W == -24) return launch_two_phase<14, 8, -24>(st, ctx, P, d_ws, ws_bytes, total_out_bytes, n);
    if (c.D == -2 && c.W == -32) return launch_two_phase<14, 8, -32>(st, ctx, P, d_ws, ws_bytes, total_out_bytes, n);
    if (c.D == -3 && c.W == 15) return launch_two_phase<15, 8, 0>(st, ctx, P, d_ws, ws_bytes, total_out_bytes, n);
    if (c.D == -3 && c.W == 13) return launch_two_phase<13, 8, 0>(st, ctx, P, d_ws, ws_bytes, total_out_bytes, n);
    if (c.D == -3 && c.W == 4) return launch_two_phase<14, 4, 0>(st, ctx, P, d_ws, ws_bytes, total_out_bytes, n);
    if (c.D == -3 && c.W == 16) return launch_two_phase<14, 16, 0>(st, ctx, P, d_ws, ws_bytes, total_out_bytes, n);
    if (c.D == -2 && c.W == 1) return launch_two_phase<14, 8, 1>(st, ctx, P, d_ws, ws_bytes, total_out_bytes, n);
    if (c.D == -2 && c.W == 4) return launch_two_phase<14, 8, 4>(st, ctx, P, d_ws, ws_bytes, total_out_bytes, n);
    if (c.D == -2 && c.W == 8) return launch_two_phase<14, 8, 8>(st, ctx, P, d_ws, ws_bytes, total_out_bytes, n);
    // D = -1: lane-per-stream canonical-decode kernel with W warps per CTA
    if (c.D == -1 && c.W == 14) return launch_lc<14>(st, ctx, P);
    if (c.D == -1 && c.W == 12) return launch_lc<12>(st, ctx, P);
    if (c.D == -1 && c.W == 10) return launch_lc<10>(st, ctx, P);
    if (c.D == -1 && c.W == 8) return launch_lc<8>(st, ctx, P);
    if (c.D == -1 && c.W == 7) return launch_lc<7>(st, ctx, P);
    if (c.D == -1 && c.W == 6) return launch_lc<6>(st, ctx, P);
    if (c.D == -1 && c.W == 5) return launch_lc<5>(st, ctx, P);
    if (c.D == -1 && c.W == 4) return launch_lc<4>(st, ctx, P);
    // other negative D select the table-based lane-per-stream kernel: D = -LB, W = DB
    if (c.D == -9 && c.W == 8) return launch_lane<9, 8, 3>(st, ctx, P);
    if (c.D == -8 && c.W == 7) return launch_lane<8, 7, 4>(st, ctx, P);
    if (c.D == -10 && c.W == 8) return launch_lane<10, 8, 2>(st, ctx, P);
    if (c.D == -9 && c.W == 7) return launch_lane<9, 7, 3>(st, ctx, P);
#endif
    set_error("inflate configuration %d,%d is not in this build (experiments need make EXPERIMENTS=1)", c.D, c.W);
    return CZ_E_STREAM;
}

}  // namespace czh

using namespace czh;

extern "C" int cz_has_experiments(void) {
#ifdef CZ_EXPERIMENTS
    return 1;
#else
    return 0;
#endif
}

extern "C" int cz_tune_inflate_lz(int cta_mode, int spin_ns) {
#ifndef CZ_EXPERIMENTS
    if (cta_mode != 0) { set_error("phase B variants need a build with -DCZ_EXPERIMENTS"); return CZ_E_STREAM; }
#endif
    g_lz_cta = cta_mode;
    g_lz_spin = spin_ns;
    return 0;
}

extern "C" uint64_t cz_inflate_workspace_bytes(size_t n, uint64_t total_out_bytes) { return inflate_workspace_bytes(n, total_out_bytes); }

extern "C" int cz_tune_inflate(int slots_per_warp, int warps_per_cta) {
#ifndef CZ_EXPERIMENTS
    if (!((slots_per_warp == -2 && (warps_per_cta == 14 || warps_per_cta == 0)) || (slots_per_warp == 1 && warps_per_cta == 8))) {
        set_error("inflate configuration %d,%d needs a build with -DCZ_EXPERIMENTS", slots_per_warp, warps_per_cta);
        return CZ_E_STREAM;
    }
#endif
    g_cfg.D = slots_per_warp;
    g_cfg.W = warps_per_cta;
    return 0;
}

extern "C" int cz_inflate_batch_device(void *cuda_stream, size_t n, const uint8_t *d_in, const uint64_t *d_in_offsets,
                                       uint8_t *d_out, const uint64_t *d_out_offsets, uint64_t total_out_bytes, uint64_t *d_out_lens,
                                       int32_t *d_statuses, uint64_t *d_in_consumed, int window_bits, void *d_workspace,
                                       uint64_t workspace_bytes) {
    int dev = 0;
    if (!CZ_CUDA(cudaGetDevice(&dev))) return CZ_E_NO_DEVICE;
    DeviceCtx *ctx = device_ctx(dev);
    if (!ctx) return CZ_E_NO_DEVICE;
    return launch_inflate((cudaStream_t)cuda_stream, ctx, n, d_in, d_in_offsets, d_out, d_out_offsets, d_out_lens, d_statuses,
                          d_in_consumed, nullptr, window_bits, 0, 0, d_workspace, workspace_bytes, total_out_bytes, nullptr, 0, 0);
}

extern "C" int cz_inflate_segments_device(void *cuda_stream, size_t n, const uint8_t *d_in, const uint64_t *d_in_offsets,
                                          uint8_t *d_out, const uint64_t *d_out_offsets, uint64_t total_out_bytes, uint64_t *d_out_lens,
                                          int32_t *d_statuses, uint32_t *d_checks, void *d_workspace, uint64_t workspace_bytes) {
    int dev = 0;
    if (!CZ_CUDA(cudaGetDevice(&dev))) return CZ_E_NO_DEVICE;
    DeviceCtx *ctx = device_ctx(dev);
    if (!ctx) return CZ_E_NO_DEVICE;
    return launch_inflate((cudaStream_t)cuda_stream, ctx, n, d_in, d_in_offsets, d_out, d_out_offsets, d_out_lens, d_statuses,
                          nullptr, d_checks, -15, 1, d_checks ? 3 : 0, d_workspace, workspace_bytes, total_out_bytes, nullptr, 0, 0);
}

extern "C" void cz_profile_enable(int on) { g_prof_on.store(on != 0); }

extern "C" int cz_profile_read(double *ms_decode, double *ms_resolve) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    double a = 0, b = 0;
    int k = 0;
    for (ProfRec &r : g_prof) {
        float x = 0, y = 0;
        if (cudaEventSynchronize(r.e2) == cudaSuccess && cudaEventElapsedTime(&x, r.e0, r.e1) == cudaSuccess &&
            cudaEventElapsedTime(&y, r.e1, r.e2) == cudaSuccess) { a += x; b += y; k++; }
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); cudaEventDestroy(r.e2);
    }
    g_prof.clear();
    if (ms_decode) *ms_decode = a;
    if (ms_resolve) *ms_resolve = b;
    return k;
}
