// synth_host.cpp — host-only build of the synthetic-data generator (TEST / BENCH INFRASTRUCTURE, not the product).
//
// bench.py's `--impl reference` arm and the CPU-only tests need the same bytes the GPU arm generates (SURVEY.md 8d:
// "identical bytes in C, CUDA and Python") without loading the product's CUDA library. This file compiles the generator's
// __host__ __device__ source (compu_b200/csrc/synth.cuh, integer arithmetic only) with g++ into oracle/libcompu_synth.so.
// It contains no codec code.
#define CZK_MODEL 1
#include <math.h>
#include <string.h>

#include <vector>

#include "../compu_b200/csrc/synth.cuh"

extern "C" uint64_t oz_synth_model_bytes(void) { return sizeof(czk::SynthModel); }

extern "C" int oz_synth_build_model(const uint8_t *corpus, uint64_t corpus_len, uint8_t *model_out) {
    return czk::synth_build_model(corpus, corpus_len, (czk::SynthModel *)model_out);
}

extern "C" int oz_synth_fill(int kind, uint64_t base_seed, size_t n, uint8_t *out, const uint64_t *offsets, const uint8_t *model) {
    if (kind < 0 || kind > 3) return -2;
    const czk::SynthModel *m = (const czk::SynthModel *)model;
#pragma omp parallel for schedule(dynamic, 4)
    for (long u = 0; u < (long)n; u++)
        czk::synth_fill_unit(m, kind, base_seed, (uint64_t)u, out + offsets[u], offsets[u + 1] - offsets[u], offsets[u]);
    return 0;
}
