// synth.cu — synthetic-data entry points of the C ABI (cz_synth_*), see synth.cuh.
#include <math.h>

#include <vector>

#include "host_common.h"
#include "synth.cuh"

using namespace czh;

extern "C" uint64_t cz_synth_model_bytes(void) { return sizeof(czk::SynthModel); }

extern "C" int cz_synth_build_model(const uint8_t *corpus, uint64_t corpus_len, uint8_t *model_out) {
    if (!corpus || corpus_len < 3 || !model_out) return CZ_E_STREAM;
    czk::SynthModel *m = (czk::SynthModel *)model_out;
    memset(m, 0, sizeof *m);
    // order-2 successor counts
    std::vector<uint32_t> freq((size_t)65536 * 256, 0);
    for (uint64_t i = 2; i < corpus_len; i++) {
        uint32_t ctx = ((uint32_t)corpus[i - 2] << 8) | corpus[i - 1];
        freq[(size_t)ctx * 256 + corpus[i]]++;
    }
    m->start_ctx = ((uint32_t)corpus[0] << 8) | corpus[1];
    uint32_t n = 0;
    for (uint32_t ctx = 0; ctx < 65536; ctx++) {
        uint32_t cum = 0, cnt = 0, off = n;
        for (uint32_t b = 0; b < 256; b++) {
            uint32_t f = freq[(size_t)ctx * 256 + b];
            if (!f) continue;
            if (n >= CZK_SYNTH_MAX_ENTRIES) { set_error("corpus has too many distinct order-2 transitions"); return CZ_E_MEM; }
            cum += f;
            m->entries[n++] = (cum << 8) | b;
            cnt++;
        }
        m->ctx_index[ctx] = (off << 12) | cnt;
        m->ctx_total[ctx] = cum;
    }
    m->n_entries = n;
    // Zipf(1.2) over 4096 ranks: host doubles here, the kernels only see the integer table
    double tot = 0;
    for (uint32_t r = 1; r <= CZK_SYNTH_PHRASES; r++) tot += pow((double)r, -1.2);
    double acc = 0;
    for (uint32_t r = 1; r <= CZK_SYNTH_PHRASES; r++) {
        acc += pow((double)r, -1.2);
        double v = acc / tot * 4294967295.0;
        m->zipf_cum[r - 1] = v >= 4294967295.0 ? 0xffffffffu : (uint32_t)v;
    }
    m->zipf_cum[CZK_SYNTH_PHRASES - 1] = 0xffffffffu;
    return 0;
}

extern "C" int cz_synth_fill_device(void *cuda_stream, int kind, uint64_t base_seed, size_t n, uint8_t *d_out,
                                    const uint64_t *d_offsets, const uint8_t *d_model) {
    if (kind < 0 || kind > 3 || n > 0xfffffff0u) return CZ_E_STREAM;
    if (!n) return 0;
    czk::synth_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)cuda_stream>>>((const czk::SynthModel *)d_model, kind,
                                                                                         base_seed, (uint32_t)n, d_out, d_offsets);
    return CZ_CUDA(cudaGetLastError()) ? 0 : CZ_E_MEM;
}

extern "C" int cz_synth_fill_host(int kind, uint64_t base_seed, size_t n, uint8_t *out, const uint64_t *offsets,
                                  const uint8_t *model) {
    if (kind < 0 || kind > 3) return CZ_E_STREAM;
    const czk::SynthModel *m = (const czk::SynthModel *)model;
#pragma omp parallel for schedule(dynamic, 4)
    for (long u = 0; u < (long)n; u++)
        czk::synth_fill_unit(m, kind, base_seed, (uint64_t)u, out + offsets[u], offsets[u + 1] - offsets[u], offsets[u]);
    return 0;
}
