// host.cu — host side of the C ABI: device contexts, pinned/device buffer management, the batched host-memory entry
// points and the streaming Decoder that maps compu's chunked contract onto one-shot GPU kernels.
//
// Reference contract followed here (file:line under /root/reference):
//   cz_decode status map / remainders  src/decoder/mod.rs:459-486 (internal_zlib_impl_decode!)
//   reset returns the instance to use  src/decoder/mod.rs:433-441, src/decoder/zlib_ng.rs:99-108
//   describe_error static text         src/decoder/zlib_ng.rs:118-123 (zError), src/utils.rs:4-13
//   compu_malloc/compu_free            src/mem.rs:27-49 (here: page-locked variants for DMA)
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <thread>

#include "host_common.h"
#include "inflate_kernel.cuh"

namespace czh {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

bool cuda_ok(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return false;
}

static std::atomic<uint64_t> g_launches{0};
void count_launches(unsigned k) { g_launches.fetch_add(k, std::memory_order_relaxed); }

static std::mutex g_ctx_mu;
static DeviceCtx g_ctx[64];
static int g_ndev = -1;

static int probe_devices() {
    if (g_ndev >= 0) return g_ndev;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
    if (n > 64) n = 64;
    g_ndev = n;
    return n;
}

DeviceCtx *device_ctx(int dev) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    int n = probe_devices();
    if (dev < 0 || dev >= n) { set_error("no CUDA device %d (found %d)", dev, n); return nullptr; }
    DeviceCtx &c = g_ctx[dev];
    if (c.dev == dev) return c.ok ? &c : nullptr;
    c.dev = dev;
    cudaDeviceProp prop;
    if (!CZ_CUDA(cudaGetDeviceProperties(&prop, dev))) return nullptr;
    if (prop.major != 10) {  // the only code in this library is sm_100a SASS
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
        return nullptr;
    }
    c.sm_count = prop.multiProcessorCount;
    int prev = 0;
    cudaGetDevice(&prev);
    if (!CZ_CUDA(cudaSetDevice(dev))) return nullptr;
    if (const char *g = getenv("CZ_L2_FETCH")) {  // experiment knob: L2 fetch granularity hint (32/64/128 bytes)
        int v = atoi(g);
        if (v == 32 || v == 64 || v == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)v);
    }
    czk::CrcTables t;
    czk::init_crc_tables(&t);
    bool ok = CZ_CUDA(cudaMalloc(&c.d_crc, sizeof t)) && CZ_CUDA(cudaMemcpy(c.d_crc, &t, sizeof t, cudaMemcpyHostToDevice));
    cudaSetDevice(prev);
    c.ok = ok;
    return ok ? &c : nullptr;
}

int usable_device_count() {
    int n;
    {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        n = probe_devices();
    }
    int k = 0;
    for (int d = 0; d < n; d++)
        if (device_ctx(d)) k++;
    return k;
}

bool DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return true;
    release();
    size_t want = align_up(bytes + bytes / 8 + 256, 256);
    if (!CZ_CUDA(cudaMalloc(&p, want))) { p = nullptr; return false; }
    cap = want;
    return true;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}
bool PinBuf::reserve(size_t bytes, bool keep, size_t keep_bytes) {
    if (bytes <= cap) return true;
    size_t want = align_up(bytes + bytes / 2 + 4096, 4096);
    void *np = nullptr;
    if (!CZ_CUDA(cudaHostAlloc(&np, want, cudaHostAllocDefault))) return false;
    if (keep && p && keep_bytes) memcpy(np, p, keep_bytes);
    if (p) cudaFreeHost(p);
    p = np;
    cap = want;
    return true;
}
void PinBuf::release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
}

// ------------------------------------------------------------------------------------------------------------------
// One device's share of a batched inflate. The shard is cut into sub-batches (by output bytes) that are enqueued round
// robin on a few streams: H2D of sub-batch k+1 overlaps the kernels of k and the D2H of k-1 (PCIe is full duplex), and the
// kernels of different sub-batches overlap each other. Device buffers span the whole shard, so sub-batches never share
// memory and need no ordering between them. Buffers are cached per thread+device.
#define CZ_INFLATE_STREAMS 8
struct InflateWork {
    DevBuf in, out, meta, ws;
    PinBuf hres;  // per-unit results land here (pinned, so the D2H copies are truly asynchronous), then go to the caller's arrays
    size_t res_n = 0, res_u0 = 0;
    uint64_t *res_lens = nullptr, *res_cons = nullptr;
    int32_t *res_stat = nullptr;
    uint32_t *res_chk = nullptr;
    cudaStream_t streams[CZ_INFLATE_STREAMS] = {};
    int dev = -1;
    bool init(int d) {
        if (dev == d && streams[0]) return true;
        dev = d;
        // sub-batches are dealt round robin, so stream i holds earlier work than stream i+1 within every round: earlier
        // sub-batches get the higher priority, so that they FINISH first and their device-to-host copies start while the later
        // ones still compute (the copy engine, not the SMs, bounds the end-to-end path). CZ_INFLATE_PRIO=0: equal priorities.
        static const bool prio = [] { const char *e = getenv("CZ_INFLATE_PRIO"); return !e || atoi(e) != 0; }();
        int lo = 0, hi = 0;  // lo = least (numerically greatest), hi = greatest priority
        if (!prio || cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) { lo = hi = 0; cudaGetLastError(); }
        for (int i = 0; i < CZ_INFLATE_STREAMS; i++) {
            const int p = hi + (lo - hi) * i / (CZ_INFLATE_STREAMS - 1);
            if (!CZ_CUDA(cudaStreamCreateWithPriority(&streams[i], cudaStreamNonBlocking, p))) return false;
        }
        return true;
    }
    ~InflateWork() {
        for (int i = 0; i < CZ_INFLATE_STREAMS; i++)
            if (streams[i]) cudaStreamDestroy(streams[i]);
    }
    // CZ_TRACE=1: device timeline of the sub-batches (events: start of the shard, H2D done, kernels done, D2H done)
    struct TraceRec { cudaEvent_t h2d, k, d2h; uint64_t in_bytes, out_bytes; int stream; };
    cudaEvent_t trace_t0 = nullptr;
    std::vector<TraceRec> trace;
    bool sync_all() {
        bool ok = true;
        for (int i = 0; i < CZ_INFLATE_STREAMS; i++)
            if (streams[i] && !CZ_CUDA(cudaStreamSynchronize(streams[i]))) ok = false;
        if (trace_t0) {
            for (size_t k = 0; k < trace.size(); k++) {
                float a = 0, b = 0, c = 0;
                cudaEventElapsedTime(&a, trace_t0, trace[k].h2d);
                cudaEventElapsedTime(&b, trace_t0, trace[k].k);
                cudaEventElapsedTime(&c, trace_t0, trace[k].d2h);
                fprintf(stderr, "[cz] sub-batch %2zu stream %d  in %7.1f MB out %7.1f MB  h2d done %7.2f  kernels done %7.2f  d2h done %7.2f ms\n",
                        k, trace[k].stream, trace[k].in_bytes / 1e6, trace[k].out_bytes / 1e6, a, b, c);
                cudaEventDestroy(trace[k].h2d); cudaEventDestroy(trace[k].k); cudaEventDestroy(trace[k].d2h);
            }
            cudaEventDestroy(trace_t0);
            trace_t0 = nullptr;
            trace.clear();
        }
        return ok;
    }
};

static uint64_t inflate_sub_batch_bytes() {
    static uint64_t v = 0;
    if (!v) {
        v = 256ull << 20;  // (128 / 256 / 512 / 1024 MiB measured on cfg2: 108 / 100.7 / 102.5 / 105 ms end to end)
        if (const char *e = getenv("CZ_INFLATE_SUB_MB")) { long m = atol(e); if (m >= 1 && m <= 65536) v = (uint64_t)m << 20; }
    }
    return v;
}

// Inflate units [u0,u1) of a packed host batch on device `dev`. Offsets are rebased so that the device buffers only
// hold this shard. Everything is enqueued; the caller waits with InflateWork::sync_all().
static int inflate_shard(InflateWork &w, int dev, size_t u0, size_t u1, const uint8_t *in, const uint64_t *in_off, uint8_t *out,
                         const uint64_t *out_off, uint64_t *out_lens, int32_t *statuses, uint64_t *in_consumed,
                         int window_bits, int segment_mode, uint32_t *checks, const uint8_t *skip) {
    DeviceCtx *ctx = device_ctx(dev);
    if (!ctx) return CZ_E_NO_DEVICE;
    if (!CZ_CUDA(cudaSetDevice(dev)) || !w.init(dev)) return CZ_E_MEM;
    const size_t n = u1 - u0;
    if (!n) return 0;
    const uint64_t ib = in_off[u0], ie = in_off[u1], ob = out_off[u0], oe = out_off[u1];
    // sub-batches by output bytes
    std::vector<size_t> cut;
    cut.push_back(0);
    {
        // graded: the first sub-batches are small so that the first device-to-host copy starts early, the rest are large
        const uint64_t full = inflate_sub_batch_bytes();
        size_t a = 0;
        while (a < n) {
            const size_t k = cut.size() - 1;
            const uint64_t lim = k < 2 ? full / 4 : k < 4 ? full / 2 : full;
            size_t e = a + 1;
            while (e < n && out_off[u0 + e + 1] - out_off[u0 + a] <= lim) e++;
            cut.push_back(e);
            a = e;
        }
    }
    const size_t nsub = cut.size() - 1;
    // big units (megabytes in one stream) go to the warp-per-stream kernel, the rest to the two-phase path
    // (a lane decodes ~4 MB/s, a warp ~26 MB/s: beyond ~256 KiB of output the lane-per-stream path becomes the tail of the batch)
    // A shard of few units leaves most lanes of the two-phase path idle anyway: there the warp-per-stream kernel takes over earlier.
    const bool few = n < 16384;
    const uint64_t big_in = few ? 40u << 10 : 96u << 10, big_out = few ? 96u << 10 : 256u << 10;
    static const long fast_mb = [] { const char *e = getenv("CZ_INFLATE_FAST_MB"); return e ? atol(e) : 0l; }();
    const uint64_t fast_head = oe - ob >= (2048ull << 20) ? (uint64_t)fast_mb << 20 : 0;
    std::vector<uint32_t> ids;  // per sub-batch: [small ids..., big ids...], relative to the sub-batch's first unit
    std::vector<size_t> n_small(nsub, 0), n_big(nsub, 0), ids_at(nsub + 1, 0);
    bool any_big = false;
    for (size_t k = 0; k < nsub; k++) {
        ids_at[k] = ids.size();
        for (int pass = 0; pass < 2; pass++)
            for (size_t i = cut[k]; i < cut[k + 1]; i++) {
                // the head of a large shard also takes the warp-per-stream kernel: it hands a 64 KiB stream back after ~2.5 ms
                // (a lane of the two-phase path needs ~11 ms), so the first device-to-host copies start that much sooner
                const bool head = fast_head && out_off[u0 + cut[k]] - ob < fast_head;
                const bool big = head || in_off[u0 + i + 1] - in_off[u0 + i] > big_in || out_off[u0 + i + 1] - out_off[u0 + i] > big_out;
                if (skip && skip[u0 + i]) { any_big = true; continue; }  // already decoded (speculative split): in neither list
                if ((int)big == pass) { ids.push_back((uint32_t)(i - cut[k])); (big ? n_big[k] : n_small[k])++; any_big |= big; }
            }
    }
    ids_at[nsub] = ids.size();
    // meta layout: in_off[n+1] out_off[n+1] out_lens[n] consumed[n] statuses[n] checks[2n] ids[n]
    const size_t m_inoff = 0, m_outoff = m_inoff + 8 * (n + 1), m_lens = m_outoff + 8 * (n + 1), m_cons = m_lens + 8 * n,
                 m_stat = m_cons + 8 * n, m_chk = align_up(m_stat + 4 * n, 8), m_ids = m_chk + 8 * n, m_total = m_ids + 4 * n;
    uint64_t ws_total = 0;
    std::vector<uint64_t> ws_off(nsub + 1, 0);
    for (size_t k = 0; k < nsub; k++) {
        ws_total += align_up(inflate_workspace_bytes(cut[k + 1] - cut[k], out_off[u0 + cut[k + 1]] - out_off[u0 + cut[k]]), 256) + 256;
        ws_off[k + 1] = ws_total;
    }
    if (!w.in.reserve(ie - ib + 16) || !w.out.reserve(oe - ob + 16) || !w.meta.reserve(m_total) || !w.ws.reserve(ws_total) ||
        !w.hres.reserve(28 * n + 64)) return CZ_E_MEM;
    w.res_n = n; w.res_u0 = u0;
    w.res_lens = w.hres.as<uint64_t>(); w.res_cons = w.res_lens + n; w.res_stat = (int32_t *)(w.res_cons + n);
    w.res_chk = (uint32_t *)(w.res_stat + n);
    std::vector<uint64_t> offs(2 * (n + 1));
    for (size_t i = 0; i <= n; i++) { offs[i] = in_off[u0 + i] - ib; offs[n + 1 + i] = out_off[u0 + i] - ob; }
    uint8_t *dm = w.meta.as<uint8_t>();
    if (!CZ_CUDA(cudaMemcpyAsync(dm, offs.data(), 16 * (n + 1), cudaMemcpyHostToDevice, w.streams[0]))) return CZ_E_MEM;
    if (any_big && !ids.empty() && !CZ_CUDA(cudaMemcpyAsync(dm + m_ids, ids.data(), 4 * ids.size(), cudaMemcpyHostToDevice, w.streams[0]))) return CZ_E_MEM;
    if (!CZ_CUDA(cudaStreamSynchronize(w.streams[0]))) return CZ_E_MEM;  // `offs`/`ids` are stack-lifetime pageable buffers
    static const int tracing = getenv("CZ_TRACE") ? 1 : 0;
    if (tracing) {
        cudaEventCreate(&w.trace_t0);
        cudaEventRecord(w.trace_t0, w.streams[0]);
        for (int i = 1; i < CZ_INFLATE_STREAMS; i++) cudaStreamWaitEvent(w.streams[i], w.trace_t0, 0);
    }
    for (size_t k = 0; k < nsub; k++) {
        cudaStream_t st = w.streams[k % CZ_INFLATE_STREAMS];
        const size_t a = cut[k], b = cut[k + 1], nk = b - a;
        InflateWork::TraceRec tr = {nullptr, nullptr, nullptr, 0, 0, (int)(k % CZ_INFLATE_STREAMS)};
        if (tracing) { cudaEventCreate(&tr.h2d); cudaEventCreate(&tr.k); cudaEventCreate(&tr.d2h); }
        const uint64_t ia = offs[a], ibk = offs[b], oa = offs[n + 1 + a], obk = offs[n + 1 + b];
        if (any_big && !n_small[k] && !n_big[k]) continue;  // every unit of this sub-batch was decoded by the speculative split
        if (!skip) {
            if (ibk > ia && !CZ_CUDA(cudaMemcpyAsync(w.in.as<uint8_t>() + ia, in + ib + ia, ibk - ia, cudaMemcpyHostToDevice, st))) return CZ_E_MEM;
        } else {
            // units already decoded by the block-parallel path need no copy: the runs between them do
            size_t i = a;
            while (i < b) {
                while (i < b && skip[u0 + i]) i++;
                size_t e = i;
                while (e < b && !skip[u0 + e]) e++;
                const uint64_t ra = offs[i], rb = offs[e];
                if (rb > ra && !CZ_CUDA(cudaMemcpyAsync(w.in.as<uint8_t>() + ra, in + ib + ra, rb - ra, cudaMemcpyHostToDevice, st))) return CZ_E_MEM;
                i = e;
            }
        }
        if (tracing) { cudaEventRecord(tr.h2d, st); tr.in_bytes = ibk - ia; tr.out_bytes = obk - oa; }
        const uint32_t *d_ids = any_big ? (const uint32_t *)(dm + m_ids) + ids_at[k] : nullptr;
        uint8_t *wsk = w.ws.as<uint8_t>() + ws_off[k];
        const uint64_t wsk_bytes = ws_off[k + 1] - ws_off[k] - 256;
        for (int big = 0; big < 2; big++) {
            if (big && !n_big[k]) continue;
            if (!big && any_big && !n_small[k]) continue;
            int r = launch_inflate(st, ctx, nk, w.in.as<uint8_t>(), (const uint64_t *)(dm + m_inoff) + a, w.out.as<uint8_t>(),
                                   (const uint64_t *)(dm + m_outoff) + a, (uint64_t *)(dm + m_lens) + a, (int32_t *)(dm + m_stat) + a,
                                   (uint64_t *)(dm + m_cons) + a, checks ? (uint32_t *)(dm + m_chk) + 2 * a : nullptr, window_bits,
                                   segment_mode, checks ? 3 : 0, big ? wsk + wsk_bytes : wsk, big ? 256 : wsk_bytes, obk - oa,
                                   d_ids ? d_ids + (big ? n_small[k] : 0) : nullptr, big ? n_big[k] : n_small[k],
                                   big | (nsub == 1 && !skip ? 2 : 0));
            if (r) return r;
        }
        if (tracing) cudaEventRecord(tr.k, st);
        if (!skip) {
            if (obk > oa && !CZ_CUDA(cudaMemcpyAsync(out + ob + oa, w.out.as<uint8_t>() + oa, obk - oa, cudaMemcpyDeviceToHost, st))) return CZ_E_MEM;
        } else {
            // units already decoded in place by the speculative split must not be overwritten: copy the runs between them
            size_t i = a;
            while (i < b) {
                while (i < b && skip[u0 + i]) i++;
                size_t e = i;
                while (e < b && !skip[u0 + e]) e++;
                const uint64_t ra = offs[n + 1 + i], rb = offs[n + 1 + e];
                if (rb > ra && !CZ_CUDA(cudaMemcpyAsync(out + ob + ra, w.out.as<uint8_t>() + ra, rb - ra, cudaMemcpyDeviceToHost, st))) return CZ_E_MEM;
                i = e;
            }
        }
        if (!CZ_CUDA(cudaMemcpyAsync(w.res_lens + a, dm + m_lens + 8 * a, 8 * nk, cudaMemcpyDeviceToHost, st))) return CZ_E_MEM;
        if (!CZ_CUDA(cudaMemcpyAsync(w.res_stat + a, dm + m_stat + 4 * a, 4 * nk, cudaMemcpyDeviceToHost, st))) return CZ_E_MEM;
        if (in_consumed && !CZ_CUDA(cudaMemcpyAsync(w.res_cons + a, dm + m_cons + 8 * a, 8 * nk, cudaMemcpyDeviceToHost, st))) return CZ_E_MEM;
        if (checks && !CZ_CUDA(cudaMemcpyAsync(w.res_chk + 2 * a, dm + m_chk + 8 * a, 8 * nk, cudaMemcpyDeviceToHost, st))) return CZ_E_MEM;
        if (tracing) { cudaEventRecord(tr.d2h, st); w.trace.push_back(tr); }
    }
    return 0;
}

// Contiguous shards balanced by compressed bytes (contiguous per GPU keeps H2D one large copy).
static void split_by_bytes(size_t n, const uint64_t *off, int parts, std::vector<size_t> &cuts) {
    cuts.assign(parts + 1, n);
    cuts[0] = 0;
    const uint64_t total = off[n] - off[0];
    size_t u = 0;
    for (int p = 1; p < parts; p++) {
        const uint64_t target = off[0] + total * p / parts;
        while (u < n && off[u] < target) u++;
        cuts[p] = u;
    }
    cuts[parts] = n;
}

static std::atomic<uint64_t> g_split_ok{0}, g_split_tried{0};  // long units: tried / decoded by the block-parallel path

int inflate_batch_host(size_t n, const uint8_t *in, const uint64_t *in_off, uint8_t *out, const uint64_t *out_off,
                       uint64_t *out_lens, int32_t *statuses, uint64_t *in_consumed, int window_bits, int segment_mode,
                       uint32_t *checks, uint32_t devices_mask) {
    if (probe_devices() == 0) { set_error("no CUDA device"); return CZ_E_NO_DEVICE; }
    if (!devices_mask) devices_mask = 1;
    std::vector<int> devs;
    for (int d = 0; d < 32; d++)
        if (devices_mask >> d & 1) devs.push_back(d);
    for (int d : devs)
        if (!device_ctx(d)) return CZ_E_NO_DEVICE;
    std::vector<size_t> cuts;
    split_by_bytes(n, in_off, (int)devs.size(), cuts);
    // Long streams take the block-parallel path (inflate_runs.cuh), run by run on their shard's device, WHILE the ordinary units
    // of every shard go through the serial kernels: pass 1 enqueues the shards without the long units (copies and kernels are
    // asynchronous), the host then drives the long units on other streams, and whatever the block-parallel path declines (errors,
    // truncation, slots that are too small) goes through the serial kernels in a second pass.
    struct Done { size_t i; uint64_t len, cons; int32_t st; };
    std::vector<Done> done;
    std::vector<uint8_t> skip;        // pass 1: the long units
    std::vector<uint8_t> declined;    // long units the block-parallel path left alone
    std::vector<std::vector<size_t>> long_ids(devs.size());
    static const bool no_runs = getenv("CZ_NO_RUNS") != nullptr || getenv("CZ_NO_SPLIT") != nullptr;
    if (!segment_mode && !checks && !no_runs) {
        bool any = false;
        for (size_t k = 0; k < devs.size(); k++)
            for (size_t i = cuts[k]; i < cuts[k + 1]; i++)
                if (in_off[i + 1] - in_off[i] >= runs_min_unit_bytes()) { long_ids[k].push_back(i); any = true; }
        if (any) {
            skip.assign(n, 0);
            for (size_t k = 0; k < devs.size(); k++)
                for (size_t i : long_ids[k]) skip[i] = 1;
        }
    }
    // one work object (streams, device and pinned buffers) per device, borrowed from a bounded per-device pool
    static DevicePool<InflateWork, 2> pool;
    InflateWork *works[32] = {};
    struct Lease {
        InflateWork **w; DevicePool<InflateWork, 2> *p;
        ~Lease() { for (int d = 0; d < 32; d++) if (w[d]) p->release(d, w[d]); }
    } lease{works, &pool};
    for (int d : devs)
        if (!(works[d] = pool.acquire(d))) { set_error("out of memory"); return CZ_E_MEM; }
    int prev = 0;
    cudaGetDevice(&prev);
    struct RestoreDev { int d; ~RestoreDev() { cudaSetDevice(d); } } restore_dev{prev};
    int rc = 0;
    // waits for the shards of one pass and collects the per-unit results of the units that pass handled
    auto finish_pass = [&](const uint8_t *skipped) {
        for (size_t k = 0; k < devs.size(); k++) {
            if (!works[devs[k]]->streams[0]) continue;
            cudaSetDevice(devs[k]);
            InflateWork &w = *works[devs[k]];
            if (!w.sync_all() && !rc) rc = CZ_E_MEM;
            if (!rc && w.res_n) {
                if (!skipped) {
                    memcpy(out_lens + w.res_u0, w.res_lens, 8 * w.res_n);
                    memcpy(statuses + w.res_u0, w.res_stat, 4 * w.res_n);
                    if (in_consumed) memcpy(in_consumed + w.res_u0, w.res_cons, 8 * w.res_n);
                    if (checks) memcpy(checks + 2 * w.res_u0, w.res_chk, 8 * w.res_n);
                } else {
                    for (size_t j = 0; j < w.res_n; j++) {
                        const size_t i = w.res_u0 + j;
                        if (skipped[i]) continue;
                        out_lens[i] = w.res_lens[j];
                        statuses[i] = w.res_stat[j];
                        if (in_consumed) in_consumed[i] = w.res_cons[j];
                        if (checks) { checks[2 * i] = w.res_chk[2 * j]; checks[2 * i + 1] = w.res_chk[2 * j + 1]; }
                    }
                }
            }
            w.res_n = 0;
        }
    };
    // pass 1: enqueue every shard (copies and kernels of different devices overlap) ...
    for (size_t k = 0; k < devs.size() && !rc; k++)
        rc = inflate_shard(*works[devs[k]], devs[k], cuts[k], cuts[k + 1], in, in_off, out, out_off, out_lens, statuses,
                           in_consumed, window_bits, segment_mode, checks, skip.empty() ? nullptr : skip.data());
    // ... the long units meanwhile (one host thread per device) ...
    if (!skip.empty() && !rc) {
        std::vector<uint8_t> ok(n, 0);
        std::vector<uint64_t> cons(n, 0);
        std::vector<int> rcs(devs.size(), 0);
        auto work = [&](size_t k) {
            rcs[k] = inflate_long_units(devs[k], long_ids[k], in, in_off, out, out_off, out_lens, statuses, cons.data(), window_bits, ok.data());
        };
        std::vector<std::thread> th;
        for (size_t k = 1; k < devs.size(); k++)
            if (!long_ids[k].empty()) th.emplace_back(work, k);
        if (!long_ids[0].empty()) work(0);
        for (auto &t : th) t.join();
        for (size_t k = 0; k < devs.size(); k++) {
            if (rcs[k] && !rc) rc = rcs[k];
            for (size_t i : long_ids[k]) {
                g_split_tried++;
                if (ok[i]) { g_split_ok++; done.push_back(Done{i, out_lens[i], cons[i], statuses[i]}); }
                else { if (declined.empty()) declined.assign(n, 1); declined[i] = 0; }  // (declined[] is a skip list: 0 = decode it)
            }
        }
    }
    // ... and wait
    finish_pass(skip.empty() ? nullptr : skip.data());
    for (const Done &d : done) {
        out_lens[d.i] = d.len;
        statuses[d.i] = d.st;
        if (in_consumed) in_consumed[d.i] = d.cons;
    }
    // pass 2: what the block-parallel path declined
    if (!declined.empty() && !rc) {
        for (size_t k = 0; k < devs.size() && !rc; k++)
            rc = inflate_shard(*works[devs[k]], devs[k], cuts[k], cuts[k + 1], in, in_off, out, out_off, out_lens, statuses,
                               in_consumed, window_bits, segment_mode, checks, declined.data());
        finish_pass(declined.data());
    }
    return rc;
}

}  // namespace czh

using namespace czh;

// ------------------------------------------------------------------------------------------------------------------
extern "C" int cz_device_count(void) { return usable_device_count(); }
extern "C" const char *cz_version(void) { return "compu-b200 0.1 (sm_100a)"; }
extern "C" const char *cz_last_error(void) { return g_err; }
extern "C" uint64_t cz_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" void *cz_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (probe_devices() == 0) { set_error("no CUDA device"); return nullptr; }
    if (!CZ_CUDA(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable))) return nullptr;
    return p;
}
extern "C" void cz_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

extern "C" const char *cz_describe_error(int32_t code) {
    // same table as zError (zlib z_errmsg[]), so describe_error() parity holds for every code the backend returns
    switch (code) {
        case 2: return "need dictionary";
        case 1: return "stream end";
        case 0: return "";
        case -1: return "file error";
        case -2: return "stream error";
        case -3: return "data error";
        case -4: return "insufficient memory";
        case -5: return "buffer error";
        case -6: return "incompatible version";
        default: return "";
    }
}

extern "C" void cz_split_stats(uint64_t *tried, uint64_t *split) {
    if (tried) *tried = g_split_tried.load();
    if (split) *split = g_split_ok.load();
}

extern "C" int cz_partition_by_bytes(size_t n, const uint64_t *offsets, int parts, uint64_t *cuts) {
    if (!offsets || !cuts || parts < 1) return CZ_E_STREAM;
    std::vector<size_t> c;
    split_by_bytes(n, offsets, parts, c);
    for (int p = 0; p <= parts; p++) cuts[p] = c[p];
    return 0;
}

extern "C" uint32_t cz_adler32_combine(uint32_t a, uint32_t b, uint64_t len2) { return czk::adler32_combine_u(a, b, len2); }
extern "C" uint32_t cz_crc32_combine(uint32_t a, uint32_t b, uint64_t len2) { return czk::crc32_combine_u(a, b, len2); }

extern "C" int cz_inflate_batch(size_t n, const uint8_t *in, const uint64_t *in_offsets, uint8_t *out,
                                const uint64_t *out_offsets, uint64_t *out_lens, int32_t *statuses, uint64_t *in_consumed,
                                int window_bits, uint32_t devices_mask) {
    return inflate_batch_host(n, in, in_offsets, out, out_offsets, out_lens, statuses, in_consumed, window_bits, 0, nullptr,
                              devices_mask);
}

extern "C" int cz_inflate_batch_ptrs(size_t n, const uint8_t *const *in_ptrs, const size_t *in_lens, uint8_t *const *out_ptrs,
                                     const size_t *out_caps, size_t *out_lens, int32_t *statuses, int window_bits,
                                     uint32_t devices_mask) {
    // gather into the packed form (pinned), run, scatter
    std::vector<uint64_t> ioff(n + 1, 0), ooff(n + 1, 0), lens(n, 0);
    for (size_t i = 0; i < n; i++) { ioff[i + 1] = ioff[i] + in_lens[i]; ooff[i + 1] = ooff[i] + out_caps[i]; }
    struct PinPair { PinBuf in, out; };
    static DevicePool<PinPair, 2> pin_pool;
    PinPair *pp = pin_pool.acquire(0);
    if (!pp) return CZ_E_MEM;
    struct Lease { PinPair *p; DevicePool<PinPair, 2> *pool; ~Lease() { pool->release(0, p); } } lease{pp, &pin_pool};
    PinBuf &pin_in = pp->in, &pin_out = pp->out;
    if (!pin_in.reserve(ioff[n] + 16) || !pin_out.reserve(ooff[n] + 16)) return CZ_E_MEM;
    for (size_t i = 0; i < n; i++) memcpy(pin_in.as<uint8_t>() + ioff[i], in_ptrs[i], in_lens[i]);
    int rc = inflate_batch_host(n, pin_in.as<uint8_t>(), ioff.data(), pin_out.as<uint8_t>(), ooff.data(), lens.data(), statuses,
                                nullptr, window_bits, 0, nullptr, devices_mask);
    if (rc) return rc;
    for (size_t i = 0; i < n; i++) {
        memcpy(out_ptrs[i], pin_out.as<uint8_t>() + ooff[i], (size_t)lens[i]);
        out_lens[i] = (size_t)lens[i];
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Streaming decoder: zlib's inflate() contract, call for call (/root/reference/src/decoder/mod.rs:459-486).
//
// Every cz_decode that can make progress is ONE launch of the warp-per-stream kernel in resumable mode
// (inflate_kernel.cuh, ResumeState): the unit is [staged tail | this call's input], the output slot is exactly the caller's
// buffer, and the kernel stops where zlib's inflate() would return — out of input inside a header / a code / a stored run, or
// with the slot full in the middle of a match — leaving in the state what zlib keeps in its inflate_state. So device work is
// linear in the stream (nothing is ever decoded twice except a partial header or code), and what the decoder holds between
// calls is bounded: the bytes of the partial header or code that ended the last call (a few bytes; a gzip header with
// FEXTRA / FNAME can be longer), the 32 KiB history (on the device, in front of the next slot) and the state itself.
// Statuses and remainders are zlib's: all input taken unless the output filled up first, in which case the bytes after the
// last one whose bits were used go back to the caller (input_remain); no progress at all is Z_BUF_ERROR -> NeedOutput.
// The caller's buffers are read and written by the copy engine directly (cudaMemcpyAsync on the caller's pointers: zero
// staging copies of ours when they are cz_host_alloc'ed / pinned; the driver stages pageable memory itself).
// A stream that arrives whole in one call and is large takes the batched path first (speculative split, see above).
struct DecoderState {
    int window_bits = 0;
    int dev = 0;
    cudaStream_t stream = nullptr;
    DevBuf d_in, d_out, d_meta;
    PinBuf h_meta;                 // pinned: [ResumeState | in_off[2] | out_off[2] | out_len | status] up and down
    std::vector<uint8_t> carry;    // staged tail: bytes from the one that holds rs.bit_pos on
    czk::ResumeState rs;
    uint64_t wpos = 0;             // d_out[wpos - rs.hist_len, wpos) is the history, the next slot starts at wpos
    bool done = false;
    int32_t error = 0;             // sticky (<0, or 3 = need dictionary)
    void reset_stream() {
        memset(&rs, 0, sizeof rs);
        rs.adler = 1;
        carry.clear();
        wpos = 0;
        done = false;
        error = 0;
    }
    ~DecoderState() {
        if (stream) { cudaSetDevice(dev); cudaStreamDestroy(stream); }
    }
};

static std::atomic<int> g_stream_device{0};
namespace czh { int stream_device() { return g_stream_device.load(std::memory_order_relaxed); } }

extern "C" int cz_set_stream_device(int device) {
    if (probe_devices() == 0 || !device_ctx(device)) {
        if (!g_err[0]) set_error("no usable sm_100 CUDA device %d", device);
        return CZ_E_NO_DEVICE;
    }
    g_stream_device.store(device, std::memory_order_relaxed);
    return 0;
}

extern "C" void *cz_decoder_new(int window_bits) {
    if (!(window_bits == -15 || window_bits == 15 || window_bits == 31 || window_bits == 47)) {
        set_error("unsupported window_bits %d", window_bits);
        return nullptr;
    }
    const int dev = stream_device();
    if (probe_devices() == 0 || !device_ctx(dev)) {
        if (!g_err[0]) set_error("no usable sm_100 CUDA device");
        return nullptr;  // => Interface::zlib_cuda(mode) returns None; there is no CPU path
    }
    DecoderState *s = new (std::nothrow) DecoderState();
    if (!s) return nullptr;
    s->window_bits = window_bits;
    s->dev = dev;
    s->reset_stream();
    return s;
}

extern "C" void *cz_decoder_reset(void *state) {
    DecoderState *s = (DecoderState *)state;
    if (!s) return nullptr;
    s->reset_stream();  // keeps the stream and the device / pinned allocations (cheap reset)
    return s;
}

extern "C" void cz_decoder_free(void *state) { delete (DecoderState *)state; }

static const uint32_t kWindow = 32768;

// One resumable launch: decodes [carry | in) into out[0, out_len). Returns 0 and the kernel's verdict, or a CZ_E_* code.
static int decoder_launch(DecoderState *s, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len, uint64_t *produced,
                          int32_t *kstatus) {
    DeviceCtx *ctx = device_ctx(s->dev);
    if (!ctx) return CZ_E_NO_DEVICE;
    int prev = 0;
    cudaGetDevice(&prev);
    struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev};
    if (!CZ_CUDA(cudaSetDevice(s->dev))) return CZ_E_MEM;
    if (!s->stream && !CZ_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking))) return CZ_E_MEM;
    cudaStream_t st = s->stream;
    const size_t nc = s->carry.size(), unit = nc + in_len;
    // meta (device): [ResumeState 384 | in_off 16 | out_off 16 | out_len 8 | status 8 | pad -> 512 | counter 256]
    const size_t m_rs = 0, m_ioff = sizeof(czk::ResumeState), m_ooff = m_ioff + 16, m_len = m_ooff + 16, m_stat = m_len + 8, m_cnt = 512,
                 m_total = 768;
    static_assert(sizeof(czk::ResumeState) + 48 <= 512, "meta layout");
    if (!s->d_meta.reserve(m_total) || !s->h_meta.reserve(m_total + 512) || !s->d_in.reserve(unit + 64)) return CZ_E_MEM;
    // output ring: history directly in front of the slot. When the slot does not fit behind the cursor, the history moves to
    // the front (through a bounce area at the end of the buffer when the two ranges would overlap).
    const uint32_t hist = s->rs.hist_len;
    if (s->wpos + out_len + 16 > s->d_out.cap || s->wpos < hist) {
        const size_t need = (size_t)kWindow * 3 + out_len + 64;  // (the bounce area at the end stays clear of cursor + slot)
        if (need > s->d_out.cap) {
            DevBuf nb;
            if (!nb.reserve(need + (need >> 1))) return CZ_E_MEM;
            if (hist && !CZ_CUDA(cudaMemcpyAsync(nb.p, s->d_out.as<uint8_t>() + s->wpos - hist, hist, cudaMemcpyDeviceToDevice, st))) return CZ_E_MEM;
            if (!CZ_CUDA(cudaStreamSynchronize(st))) return CZ_E_MEM;
            std::swap(nb.p, s->d_out.p);
            std::swap(nb.cap, s->d_out.cap);
        } else if (hist) {
            uint8_t *base = s->d_out.as<uint8_t>();
            if (s->wpos - hist >= hist) {
                if (!CZ_CUDA(cudaMemcpyAsync(base, base + s->wpos - hist, hist, cudaMemcpyDeviceToDevice, st))) return CZ_E_MEM;
            } else {  // ranges overlap: bounce through the last 32 KiB of the buffer (beyond any slot in use: cap >= 2 W + ...)
                uint8_t *bounce = base + s->d_out.cap - kWindow;
                if (!CZ_CUDA(cudaMemcpyAsync(bounce, base + s->wpos - hist, hist, cudaMemcpyDeviceToDevice, st)) ||
                    !CZ_CUDA(cudaMemcpyAsync(base, bounce, hist, cudaMemcpyDeviceToDevice, st))) return CZ_E_MEM;
            }
        }
        s->wpos = hist;
    }
    uint8_t *hm = s->h_meta.as<uint8_t>();
    memcpy(hm + m_rs, &s->rs, sizeof s->rs);
    uint64_t *ioff = (uint64_t *)(hm + m_ioff), *ooff = (uint64_t *)(hm + m_ooff);
    ioff[0] = 0; ioff[1] = unit; ooff[0] = 0; ooff[1] = out_len;
    uint8_t *dm = s->d_meta.as<uint8_t>();
    if (!CZ_CUDA(cudaMemcpyAsync(dm, hm, m_len, cudaMemcpyHostToDevice, st))) return CZ_E_MEM;
    if (nc) {
        memcpy(hm + m_total, s->carry.data(), nc <= 512 ? nc : 0);
        const void *src = nc <= 512 ? (const void *)(hm + m_total) : (const void *)s->carry.data();
        if (!CZ_CUDA(cudaMemcpyAsync(s->d_in.p, src, nc, cudaMemcpyHostToDevice, st))) return CZ_E_MEM;
    }
    if (in_len && !CZ_CUDA(cudaMemcpyAsync(s->d_in.as<uint8_t>() + nc, in, in_len, cudaMemcpyHostToDevice, st))) return CZ_E_MEM;
    int rc = launch_inflate_resume(st, ctx, s->d_in.as<uint8_t>(), (const uint64_t *)(dm + m_ioff), s->d_out.as<uint8_t>() + s->wpos,
                                   (const uint64_t *)(dm + m_ooff), (uint64_t *)(dm + m_len), (int32_t *)(dm + m_stat), s->window_bits,
                                   (czk::ResumeState *)(dm + m_rs), dm + m_cnt);
    if (rc) return rc;
    if (!CZ_CUDA(cudaMemcpyAsync(hm, dm, m_stat + 8, cudaMemcpyDeviceToHost, st)) || !CZ_CUDA(cudaStreamSynchronize(st))) return CZ_E_MEM;
    memcpy(&s->rs, hm + m_rs, sizeof s->rs);
    *produced = *(const uint64_t *)(hm + m_len);
    *kstatus = *(const int32_t *)(hm + m_stat);
    if (*produced > out_len) { set_error("internal: resumable launch overran its slot"); return CZ_E_MEM; }
    if (*produced) {
        if (!CZ_CUDA(cudaMemcpyAsync(out, s->d_out.as<uint8_t>() + s->wpos, *produced, cudaMemcpyDeviceToHost, st)) ||
            !CZ_CUDA(cudaStreamSynchronize(st))) return CZ_E_MEM;
        s->wpos += *produced;
    }
    return 0;
}

extern "C" cz_result cz_decode(void *state, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len) {
    DecoderState *s = (DecoderState *)state;
    cz_result r;
    r.input_remain = in_len;
    r.output_remain = out_len;
    if (!s) { r.status = CZ_E_STREAM; return r; }
    if (s->error) { r.status = s->error; return r; }         // zlib: mode BAD stays BAD
    if (s->done) { r.status = CZ_DECODE_FINISHED; return r; }  // zlib: mode DONE returns Z_STREAM_END again, nothing consumed
    // a large stream that arrives whole at the very start: the batched path (speculative split at verified full-flush points)
    static const bool no_runs = getenv("CZ_NO_RUNS") != nullptr || getenv("CZ_NO_SPLIT") != nullptr;
    const bool fresh = s->rs.phase == czk::RP_HEADER && s->rs.total_out == 0 && s->carry.empty();
    if (fresh && !no_runs && in_len >= (1u << 20)) {
        const uint64_t ioff[2] = {0, in_len}, ooff[2] = {0, out_len};
        uint64_t got = 0, used = 0;
        int32_t st = 0;
        uint8_t done1 = 0;
        int prev = 0;
        cudaGetDevice(&prev);
        g_split_tried++;
        const int rc = inflate_long_units(s->dev, std::vector<size_t>(1, 0), in, ioff, out, ooff, &got, &st, &used, s->window_bits, &done1);
        cudaSetDevice(prev);
        if (rc == 0 && done1) {  // the whole stream, bit-exact and checked: Finished; what follows it goes back to the caller
            g_split_ok++;
            s->done = true;
            r.output_remain = out_len - (size_t)got;
            r.input_remain = in_len - (size_t)used;
            r.status = CZ_DECODE_FINISHED;
            return r;
        }
    }
    uint64_t produced = 0;
    int32_t ks = 0;
    int rc = decoder_launch(s, in, in_len, out, out_len, &produced, &ks);
    if (rc) { s->error = rc; r.status = rc; return r; }
    const size_t nc = s->carry.size(), unit = nc + in_len;
    const uint64_t B = s->rs.bit_pos;
    r.output_remain = out_len - (size_t)produced;
    if (ks == CZ_DECODE_FINISHED) {
        s->done = true;
        const uint64_t used = (B + 7) >> 3;               // the trailer ends on a byte boundary
        r.input_remain = unit > used ? (size_t)(unit - used) : 0;  // bytes after the end of the stream belong to the caller
        if (r.input_remain > in_len) r.input_remain = in_len;
        s->carry.clear();
        r.status = CZ_DECODE_FINISHED;
        return r;
    }
    if (ks < 0 || ks == CZ_DECODE_NEED_DICT) {  // everything decoded before the error has been written (as zlib does)
        s->error = ks;
        r.input_remain = 0;
        r.status = ks;
        return r;
    }
    // Z_OK: rebuild the staged tail from the byte that holds the next bit
    std::vector<uint8_t> tail;
    size_t keep_from = (size_t)(B >> 3), keep_to;
    if (ks == CZ_DECODE_NEED_INPUT) { keep_to = unit; r.input_remain = 0; }        // everything is taken; the partial item stays staged
    else { keep_to = (size_t)((B + 7) >> 3); r.input_remain = unit - keep_to; }  // slot full: only the byte in use stays
    if (r.input_remain > in_len) { set_error("internal: resumable decoder gave back more than it was given"); s->error = CZ_E_MEM; r.status = CZ_E_MEM; return r; }
    tail.reserve(keep_to > keep_from ? keep_to - keep_from : 0);
    for (size_t k = keep_from; k < keep_to; k++) tail.push_back(k < nc ? s->carry[k] : in[k - nc]);
    s->carry.swap(tail);
    s->rs.bit_pos = B & 7;
    const size_t used_in = in_len - r.input_remain;
    if (r.input_remain == 0 && !(used_in == 0 && produced == 0)) r.status = CZ_DECODE_NEED_INPUT;
    else r.status = CZ_DECODE_NEED_OUTPUT;  // output full — or no progress at all: Z_BUF_ERROR -> NeedOutput (mod.rs:481)
    return r;
}
