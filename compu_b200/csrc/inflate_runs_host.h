// inflate_runs_host.h — host orchestration of the block-parallel inflate of long streams (kernels: inflate_runs.cuh).
//
// Written against a small backend interface so that the SAME logic drives the CUDA kernels (inflate.cu) and the CPU emulator
// of the kernel sources (tests/cusim): chunking, run construction from the candidates, verification of the run chain (run r
// must end exactly where run r+1 starts; a candidate the chain passes over is a false positive and its range is decoded again
// from the true boundary), layout, and the container trailer check with per-run Adler-32 / CRC-32 folded on the host.
//
// The parallel path only ever reports a stream as FINISHED with exactly its bytes. Anything else — a data error, a truncated
// stream, an output slot that is too small, a chain that does not close — marks the unit `ok = false` and the caller decodes
// it on the serial path, which reproduces zlib's partial output, status and consumed count for those cases.
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

#include "inflate_runs.cuh"

namespace czh {

struct BigUnit {
    // in
    const uint8_t *h_in = nullptr;   // the stream's bytes on the host (container header / trailer are read there)
    uint64_t in_len = 0;
    uint64_t d_in_lo = 0;            // byte offset of the stream in the backend's device input buffer
    uint64_t out_cap = 0;            // the caller's slot: a stream that needs more is left to the serial path
    uint64_t d_out_off = 0;          // byte offset of its slot in the backend's device output buffer (slots must not overlap;
                                     // mirroring the caller's layout lets neighbouring units leave in one copy)
    int window_bits = 15;            // -15 raw, 15 zlib, 31 gzip, 47 auto
    // out
    bool ok = false;
    uint64_t out_len = 0, in_consumed = 0;
    uint32_t n_runs = 0;
};

// Backend: device memory + kernel launches. All calls are synchronous from the caller's point of view (results of a launch are
// complete when d2h returns). scratch() hands out 256-byte aligned device memory from an arena that scratch_reset() rewinds.
//   uint8_t *d_in(); uint8_t *d_out();
//   void scratch_reset(); bool scratch_need(size_t total); void *scratch(size_t bytes);   // scratch_need right after a reset:
//                                                       // make room for `total` bytes of scratch() calls (nullptr / false: out of memory)
//   bool out_need(size_t bytes);                        // size the device output buffer (before anything is written to it)
//   void *tok_buffer(size_t bytes);                     // the token buffer of this batch (kept until the next batch)
//   bool h2d(void *d, const void *h, size_t n); bool d2h(void *h, const void *d, size_t n); bool zero(void *d, size_t n);
//   bool candidates(const czk::CandChunk *d_chunks, uint32_t n, uint64_t *d_cand);
//   bool tok(const czk::TwoPhaseParams &Q); bool lz16(const czk::TwoPhaseParams &Q, uint16_t *d_sym);
//   bool tail_markers(const uint64_t *d_run_off, uint32_t n_runs, const uint16_t *d_sym, uint8_t *d_flags);
//   bool window(const czk::RunStream *d_streams, uint32_t n_streams, const uint64_t *d_run_off, const uint16_t *d_sym, uint8_t *d_win, uint32_t *d_bad);
//   bool resolve(const czk::RunSlice *d_slices, uint32_t n_slices, const uint64_t *d_run_off, const uint64_t *d_final_off,
//                const uint8_t *d_first, const uint16_t *d_sym, const uint8_t *d_win, uint8_t *d_out, uint32_t *d_bad);
//   bool check(const uint64_t *d_run_off, const uint64_t *d_final_off, uint32_t n_runs, const uint8_t *d_out, int kind, uint32_t *d_checks);
//   const czk::CrcTables *crc();
//   void mark(const char *phase);   // optional timeline hook (CZ_TRACE): called after each phase has been enqueued

static inline uint64_t runs_align(uint64_t v) { return (v + 255) & ~255ull; }

// wrap kind and the offset of the container trailer check; returns false when the trailer is incomplete or wrong
static inline bool runs_trailer_ok(const BigUnit &u, int wrap, uint64_t end_byte, uint32_t adler, uint32_t crc, uint64_t total, uint64_t *consumed) {
    const uint8_t *t = u.h_in + end_byte;
    if (wrap == 0) { *consumed = end_byte; return true; }
    if (wrap == 1) {
        if (end_byte + 4 > u.in_len) return false;
        if (((uint32_t)t[0] << 24 | (uint32_t)t[1] << 16 | (uint32_t)t[2] << 8 | t[3]) != adler) return false;
        *consumed = end_byte + 4;
        return true;
    }
    if (end_byte + 8 > u.in_len) return false;
    const uint32_t c = (uint32_t)t[0] | (uint32_t)t[1] << 8 | (uint32_t)t[2] << 16 | (uint32_t)t[3] << 24;
    const uint32_t isz = (uint32_t)t[4] | (uint32_t)t[5] << 8 | (uint32_t)t[6] << 16 | (uint32_t)t[7] << 24;
    if (c != crc || isz != (uint32_t)total) return false;
    *consumed = end_byte + 8;
    return true;
}

template <class BK>
int inflate_runs_batch(BK &bk, std::vector<BigUnit> &units, uint64_t chunk_bytes) {
    using namespace czk;
    const size_t nu = units.size();
    if (!nu) return 0;
    if (chunk_bytes < 4096) chunk_bytes = 4096;
    bk.scratch_reset();
    // ---- 1. chunks and candidates (chunk 0 of a stream needs none: run 0 starts at the container header)
    std::vector<CandChunk> chunks;
    std::vector<uint32_t> chunk_unit;
    for (size_t u = 0; u < nu; u++) {
        units[u].ok = false;
        const uint64_t lo = units[u].d_in_lo, hi = lo + units[u].in_len;
        for (uint64_t a = lo + chunk_bytes; a < hi; a += chunk_bytes) {
            CandChunk c;
            c.lo_bit = a * 8;
            c.hi_bit = (a + chunk_bytes < hi ? a + chunk_bytes : hi) * 8;
            c.end_bit = hi * 8;
            chunks.push_back(c);
            chunk_unit.push_back((uint32_t)u);
        }
    }
    std::vector<uint64_t> cand(chunks.size(), ~0ull);
    bk.mark("start");
    if (!chunks.empty()) {
        if (!bk.scratch_need(chunks.size() * (sizeof(CandChunk) + 8) + 1024)) return -4;
        CandChunk *d_chunks = (CandChunk *)bk.scratch(sizeof(CandChunk) * chunks.size());
        uint64_t *d_cand = (uint64_t *)bk.scratch(8 * chunks.size());
        if (!d_chunks || !d_cand) return -4;
        if (!bk.h2d(d_chunks, chunks.data(), sizeof(CandChunk) * chunks.size())) return -4;
        if (!bk.candidates(d_chunks, (uint32_t)chunks.size(), d_cand)) return -4;
        if (!bk.d2h(cand.data(), d_cand, 8 * chunks.size())) return -4;
    }
    bk.mark("candidates");
    // ---- 2. runs per unit: start bits relative to the stream's first byte. Every run is decoded ONCE, straight into tokens:
    // sizes are not known yet, so its token area is bounded by its compressed span (a run that ends where the next one starts
    // has at most ~5 token words per compressed byte in theory and ~0.5 on real data: 2 words per byte + slack; a run that
    // needs more — dense tokens, or it ran past a false next candidate — counts its words and is emitted again exactly).
    struct Run {
        uint32_t unit; uint64_t start, target; uint32_t mid;
        uint64_t end, out_len; int32_t status; uint32_t fin, ntok;
        uint64_t tok_off, tok_cap;
        bool counted, reemit;
    };
    std::vector<std::vector<Run>> ur(nu);
    for (size_t u = 0; u < nu; u++) ur[u].push_back(Run{(uint32_t)u, 0, ~0ull, 0, 0, 0, 0, 0, 0, 0, 0, false, false});
    for (size_t c = 0; c < chunks.size(); c++) {
        if (cand[c] == ~0ull) continue;
        const uint32_t u = chunk_unit[c];
        ur[u].push_back(Run{u, cand[c] - units[u].d_in_lo * 8, ~0ull, 1, 0, 0, 0, 0, 0, 0, 0, false, false});
    }
    uint64_t tok_next = 0;
    for (size_t u = 0; u < nu; u++) {
        std::vector<Run> &R = ur[u];
        for (size_t r = 0; r < R.size(); r++) {
            const uint64_t span_bits = (r + 1 < R.size() ? R[r + 1].start : units[u].in_len * 8) - R[r].start;
            R[r].tok_off = tok_next; R[r].tok_cap = 2 * ((span_bits + 7) >> 3) + 1024;
            tok_next += R[r].tok_cap;
        }
    }
    // room for repairs (runs decoded again from a true boundary, runs emitted again with their exact size)
    const uint64_t tok_total = tok_next + (tok_next >> 2) + (1u << 20);
    uint32_t *d_tok = (uint32_t *)bk.tok_buffer(4 * tok_total);
    if (!d_tok) return -4;
    // ---- 3. decode rounds with chain verification; runs that turn out to be needed are decoded in further rounds
    std::vector<uint8_t> dead(nu, 0);
    for (int round = 0; round < 6; round++) {
        std::vector<Run *> todo;
        for (size_t u = 0; u < nu; u++) {
            if (dead[u]) continue;
            std::vector<Run> &R = ur[u];
            for (size_t r = 0; r < R.size(); r++) {
                if (!R[r].reemit) R[r].target = r + 1 < R.size() ? R[r + 1].start : ~0ull;
                if (!R[r].counted || R[r].reemit) todo.push_back(&R[r]);
            }
        }
        if (todo.empty()) break;
        const size_t n = todo.size();
        bk.scratch_reset();
        if (!bk.scratch_need(n * (sizeof(RunDesc) + sizeof(RunResult) + sizeof(TokMeta) + 8 + 8 + 4 + 4) + 16 * 256 + 1024)) return -4;
        RunDesc *d_runs = (RunDesc *)bk.scratch(sizeof(RunDesc) * n);
        RunResult *d_res = (RunResult *)bk.scratch(sizeof(RunResult) * n);
        TokMeta *d_meta = (TokMeta *)bk.scratch(sizeof(TokMeta) * n);
        uint64_t *d_off = (uint64_t *)bk.scratch(8 * (n + 1));
        uint64_t *d_lens = (uint64_t *)bk.scratch(8 * n);
        int32_t *d_stat = (int32_t *)bk.scratch(4 * n);
        unsigned long long *d_cnt = (unsigned long long *)bk.scratch(256);
        uint32_t *d_ids = (uint32_t *)bk.scratch(4 * n);
        if (!d_runs || !d_res || !d_meta || !d_off || !d_lens || !d_stat || !d_cnt || !d_ids) return -4;
        std::vector<RunDesc> hr(n);
        std::vector<uint64_t> hoff(n + 1);
        for (size_t i = 0; i < n; i++) {
            const BigUnit &U = units[todo[i]->unit];
            hr[i].in_lo = U.d_in_lo; hr[i].in_hi = U.d_in_lo + U.in_len;
            hr[i].start_bit = todo[i]->start; hr[i].target_bit = todo[i]->target; hr[i].mid_stream = todo[i]->mid; hr[i].pad = 0;
            hr[i].tok_off = todo[i]->tok_off; hr[i].tok_cap = todo[i]->tok_cap;
            hoff[i] = (uint64_t)i << 40;  // "unlimited" output slots: sizes are what this pass finds out
        }
        hoff[n] = (uint64_t)n << 40;
        if (!bk.h2d(d_runs, hr.data(), sizeof(RunDesc) * n) || !bk.h2d(d_off, hoff.data(), 8 * (n + 1))) return -4;
        TwoPhaseParams Q;
        memset(&Q, 0, sizeof Q);
        Q.base.in = bk.d_in(); Q.base.in_off = nullptr; Q.base.out = nullptr; Q.base.out_off = d_off; Q.base.out_lens = d_lens;
        Q.base.statuses = d_stat; Q.base.counter = d_cnt; Q.base.crc = bk.crc();
        Q.meta = d_meta; Q.counter_b = d_cnt + 16; Q.count_only = 0; Q.tok = d_tok; Q.runs = d_runs; Q.run_res = d_res;
        // window_bits is a launch parameter: the runs of each container kind go in one launch each
        std::vector<TokMeta> hm(n);
        std::vector<RunResult> hres(n);
        {
            const int kinds[4] = {-15, 15, 31, 47};
            std::vector<uint32_t> ids;
            for (int k = 0; k < 4; k++) {
                ids.clear();
                for (size_t i = 0; i < n; i++)
                    if (units[todo[i]->unit].window_bits == kinds[k]) ids.push_back((uint32_t)i);
                if (ids.empty()) continue;
                if (!bk.h2d(d_ids, ids.data(), 4 * ids.size()) || !bk.zero(d_cnt, 256)) return -4;
                Q.base.ids = d_ids; Q.base.n = (uint32_t)ids.size(); Q.base.window_bits = kinds[k];
                if (!bk.tok(Q)) return -4;
                if (!bk.d2h(hm.data(), d_meta, sizeof(TokMeta))) return -4;  // (synchronises before d_ids is reused)
            }
        }
        if (!bk.d2h(hm.data(), d_meta, sizeof(TokMeta) * n) || !bk.d2h(hres.data(), d_res, sizeof(RunResult) * n)) return -4;
        for (size_t i = 0; i < n; i++) {
            Run &r = *todo[i];
            r.counted = true; r.status = hm[i].status; r.out_len = hm[i].out_len; r.ntok = hm[i].ntok;
            r.end = hres[i].end_bit; r.fin = hres[i].final_block;
            r.reemit = hres[i].tok_overflow != 0;
            if (r.reemit) {  // sizes and positions are valid, the tokens are not: once more, with exactly the room it needs
                r.tok_off = tok_next; r.tok_cap = (uint64_t)r.ntok + 2 * 256 + 64;
                tok_next += r.tok_cap;
                r.target = r.end;
                if (tok_next > tok_total) dead[r.unit] = 1;
            }
        }
        // chain walk per unit
        for (size_t u = 0; u < nu; u++) {
            if (dead[u]) continue;
            std::vector<Run> &R = ur[u];
            std::vector<Run> keep;
            bool ok = true;
            size_t r = 0;
            while (r < R.size()) {
                Run cur = R[r];
                if (!cur.counted) { keep.push_back(cur); for (size_t k = r + 1; k < R.size(); k++) keep.push_back(R[k]); break; }
                keep.push_back(cur);
                if (cur.status == ST_FINISHED && cur.fin) break;  // the stream ends here: later candidates lie in trailing bytes
                if (cur.status != CZK_ST_RUN_END) { ok = false; break; }  // error / truncated on the true chain: serial path
                // the next run must start exactly at cur.end; runs the chain passed over are false positives
                size_t k = r + 1;
                while (k < R.size() && R[k].start < cur.end) k++;
                if (k < R.size() && R[k].start == cur.end) { r = k; continue; }
                // nothing starts at the true boundary: decode from there up to the next surviving candidate
                Run nr{(uint32_t)u, cur.end, ~0ull, 1, 0, 0, 0, 0, 0, 0, 0, false, false};
                const uint64_t span_bits = (k < R.size() ? R[k].start : units[u].in_len * 8) - cur.end;
                nr.tok_off = tok_next; nr.tok_cap = 2 * ((span_bits + 7) >> 3) + 1024;
                tok_next += nr.tok_cap;
                if (tok_next > tok_total) { ok = false; break; }
                keep.push_back(nr);
                for (size_t q = k; q < R.size(); q++) keep.push_back(R[q]);
                break;
            }
            if (!ok) { dead[u] = 1; continue; }
            R.swap(keep);
        }
        bk.mark("decode round");
    }
    // ---- 4. layout of the units whose chain closed
    struct LRun { uint32_t unit; uint64_t start, target; uint32_t mid; uint64_t out_len; uint32_t first, ntok; int32_t status; uint64_t tok_off, tok_cap, end; };
    std::vector<LRun> L;
    std::vector<RunStream> streams;
    std::vector<uint32_t> stream_unit;
    for (size_t u = 0; u < nu; u++) {
        if (dead[u]) continue;
        const std::vector<Run> &R = ur[u];
        bool closed = !R.empty() && R.back().counted && R.back().status == ST_FINISHED && R.back().fin;
        uint64_t total = 0;
        for (size_t r = 0; r < R.size() && closed; r++) {
            if (!R[r].counted || R[r].reemit) closed = false;
            total += R[r].out_len;
        }
        if (!closed || total > units[u].out_cap) continue;  // (too small a slot: the serial path reports NeedOutput exactly)
        RunStream s;
        s.first_run = (uint32_t)L.size(); s.n_runs = (uint32_t)R.size(); s.stream_start = 0;
        for (size_t r = 0; r < R.size(); r++)
            L.push_back(LRun{(uint32_t)u, R[r].start, r + 1 < R.size() ? R[r + 1].start : ~0ull, R[r].mid, R[r].out_len, r == 0 ? 1u : 0u,
                             R[r].ntok, R[r].status, R[r].tok_off, R[r].tok_cap, R[r].end});
        streams.push_back(s);
        stream_unit.push_back((uint32_t)u);
        units[u].out_len = total;
        units[u].n_runs = (uint32_t)R.size();
    }
    const size_t n = L.size();
    if (!n) return 0;
    std::vector<uint64_t> run_off(n + 1, 0), final_off(n, 0);
    std::vector<uint8_t> is_first(n, 0);
    std::vector<RunDesc> hr(n);
    std::vector<TokMeta> hmeta(n);
    std::vector<RunSlice> slices;
    {
        uint64_t within = 0, out_total = 0;
        for (size_t i = 0; i < n; i++) {
            BigUnit &U = units[L[i].unit];
            if (L[i].first) { within = 0; if (U.d_out_off + U.out_len > out_total) out_total = U.d_out_off + U.out_len; }
            run_off[i + 1] = run_off[i] + L[i].out_len;
            final_off[i] = U.d_out_off + within;
            within += L[i].out_len;
            is_first[i] = (uint8_t)L[i].first;
            hr[i].in_lo = U.d_in_lo; hr[i].in_hi = U.d_in_lo + U.in_len;
            hr[i].start_bit = L[i].start; hr[i].target_bit = L[i].target; hr[i].mid_stream = L[i].mid; hr[i].pad = 0;
            hr[i].tok_off = L[i].tok_off; hr[i].tok_cap = L[i].tok_cap;
            memset(&hmeta[i], 0, sizeof(TokMeta));
            hmeta[i].ntok = L[i].ntok; hmeta[i].status = L[i].status; hmeta[i].out_len = L[i].out_len;
            for (uint64_t s = 0; s * 65536 < L[i].out_len; s++) slices.push_back(RunSlice{(uint32_t)i, (uint32_t)s});
        }
        if (!bk.out_need(out_total + 256)) return -4;
    }
    const uint64_t total_sym = run_off[n];
    bk.scratch_reset();
    if (!bk.scratch_need(n * (sizeof(RunDesc) + sizeof(TokMeta) + 8 + 8 + 1 + 8 + 4 + 8 + 32768 + sizeof(RunStream) + 1) +
                         sizeof(RunSlice) * (slices.size() + 1) + 2 * total_sym + 24 * 256 + 4096)) return -4;
    RunDesc *d_runs = (RunDesc *)bk.scratch(sizeof(RunDesc) * n);
    TokMeta *d_meta = (TokMeta *)bk.scratch(sizeof(TokMeta) * n);
    uint64_t *d_run_off = (uint64_t *)bk.scratch(8 * (n + 1));
    uint64_t *d_final_off = (uint64_t *)bk.scratch(8 * n);
    uint8_t *d_first = (uint8_t *)bk.scratch(n);
    uint64_t *d_lens = (uint64_t *)bk.scratch(8 * n);
    int32_t *d_stat = (int32_t *)bk.scratch(4 * n);
    uint32_t *d_checks = (uint32_t *)bk.scratch(8 * n);
    unsigned long long *d_cnt = (unsigned long long *)bk.scratch(256);
    uint32_t *d_bad = (uint32_t *)bk.scratch(256);
    RunStream *d_streams = (RunStream *)bk.scratch(sizeof(RunStream) * n);  // chains: at most one per run
    uint8_t *d_flags = (uint8_t *)bk.scratch(n);
    RunSlice *d_slices = (RunSlice *)bk.scratch(sizeof(RunSlice) * (slices.size() + 1));
    uint16_t *d_sym = (uint16_t *)bk.scratch(2 * total_sym + 256);
    uint8_t *d_win = (uint8_t *)bk.scratch(32768ull * n);
    if (!d_runs || !d_meta || !d_run_off || !d_final_off || !d_first || !d_lens || !d_stat || !d_checks || !d_cnt ||
        !d_bad || !d_streams || !d_flags || !d_slices || !d_sym || !d_win) return -4;
    if (!bk.h2d(d_runs, hr.data(), sizeof(RunDesc) * n) || !bk.h2d(d_meta, hmeta.data(), sizeof(TokMeta) * n) ||
        !bk.h2d(d_run_off, run_off.data(), 8 * (n + 1)) ||
        !bk.h2d(d_final_off, final_off.data(), 8 * n) || !bk.h2d(d_first, is_first.data(), n) ||
        !bk.h2d(d_slices, slices.data(), sizeof(RunSlice) * slices.size()) || !bk.zero(d_bad, 256)) return -4;
    // ---- 5. tokens -> symbols, windows, bytes, checks
    TwoPhaseParams Q;
    memset(&Q, 0, sizeof Q);
    Q.base.in = bk.d_in(); Q.base.out = nullptr; Q.base.out_off = d_run_off; Q.base.out_lens = d_lens; Q.base.statuses = d_stat;
    Q.base.counter = d_cnt; Q.base.crc = bk.crc(); Q.base.ids = nullptr; Q.base.n = (uint32_t)n;
    Q.meta = d_meta; Q.counter_b = d_cnt + 16; Q.count_only = 0; Q.tok = d_tok; Q.runs = d_runs;
    if (!bk.zero(d_cnt, 256)) return -4;
    if (!bk.lz16(Q, d_sym)) return -4;
    bk.mark("lz16");
    // chains for the window pass: a stream's runs, cut at every run whose window does not depend on what precedes it
    std::vector<RunStream> chains;
    {
        std::vector<uint8_t> flags(n, 1);
        if (!bk.tail_markers(d_run_off, (uint32_t)n, d_sym, d_flags) || !bk.d2h(flags.data(), d_flags, n)) return -4;
        for (size_t si = 0; si < streams.size(); si++) {
            const uint32_t f = streams[si].first_run, k = streams[si].n_runs;
            for (uint32_t r = f; r < f + k; r++) {
                if (r == f || !flags[r]) chains.push_back(RunStream{r, 0, run_off[f]});
                chains.back().n_runs++;
            }
        }
        if (!bk.h2d(d_streams, chains.data(), sizeof(RunStream) * chains.size())) return -4;
    }
    if (!bk.window(d_streams, (uint32_t)chains.size(), d_run_off, d_sym, d_win, d_bad)) return -4;
    bk.mark("window");
    if (!bk.resolve(d_slices, (uint32_t)slices.size(), d_run_off, d_final_off, d_first, d_sym, d_win, bk.d_out(), d_bad)) return -4;
    bk.mark("resolve");
    int check_kind = 0;  // Adler-32 for zlib containers, CRC-32 for gzip (raw streams need neither)
    for (size_t si = 0; si < streams.size(); si++) {
        const BigUnit &U = units[stream_unit[si]];
        check_kind |= U.window_bits == 15 ? 1 : U.window_bits == 31 ? 2 : U.window_bits == 47 ? 3 : 0;
    }
    if (check_kind && !bk.check(d_run_off, d_final_off, (uint32_t)n, bk.d_out(), check_kind, d_checks)) return -4;
    if (!check_kind && !bk.zero(d_checks, 8 * n)) return -4;
    std::vector<uint32_t> checks(2 * n);
    uint32_t bad = 0;
    if (!bk.d2h(checks.data(), d_checks, 8 * n) || !bk.d2h(&bad, d_bad, 4)) return -4;
    bk.mark("checks");
    // ---- 6. per stream: fold the checks; container trailer
    for (size_t si = 0; si < streams.size(); si++) {
        BigUnit &U = units[stream_unit[si]];
        if (bad) continue;  // a marker pointed in front of a stream: some stream is corrupt — all go to the serial path
        uint32_t adler = 1, crc = 0;
        for (uint32_t k = 0; k < streams[si].n_runs; k++) {
            const size_t i = streams[si].first_run + k;
            adler = adler32_combine_u(adler, checks[2 * i], L[i].out_len);
            crc = crc32_combine_u(crc, checks[2 * i + 1], L[i].out_len);
        }
        const size_t last = streams[si].first_run + streams[si].n_runs - 1;
        const uint64_t end_byte = (L[last].end + 7) >> 3;
        const int wrap = U.window_bits < 0 ? 0 : U.window_bits == 15 ? 1 : U.window_bits == 31 ? 2 :
                         (U.in_len >= 2 && U.h_in[0] == 0x1f && U.h_in[1] == 0x8b) ? 2 : 1;
        uint64_t consumed = 0;
        if (!runs_trailer_ok(U, wrap, end_byte, adler, crc, U.out_len, &consumed)) continue;
        U.in_consumed = consumed;
        U.ok = true;
    }
    return 0;
}

}  // namespace czh
