// synth.cu — synthetic-data entry points of the C ABI (cz_synth_*), see synth.cuh.
#include <math.h>

#include <vector>

#include "host_common.h"
#include "synth.cuh"

using namespace czh;

extern "C" uint64_t cz_synth_model_bytes(void) { return sizeof(czk::SynthModel); }

extern "C" int cz_synth_build_model(const uint8_t *corpus, uint64_t corpus_len, uint8_t *model_out) {
    if (!corpus || corpus_len < 3 || !model_out) return CZ_E_STREAM;
    int rc = czk::synth_build_model(corpus, corpus_len, (czk::SynthModel *)model_out);
    if (rc == -4) { set_error("corpus has too many distinct order-2 transitions"); return CZ_E_MEM; }
    return rc ? CZ_E_STREAM : 0;
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
