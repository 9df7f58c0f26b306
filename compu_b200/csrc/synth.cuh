// synth.cuh — synthetic byte streams with controlled entropy, identical bytes on host and device.
//
// kind 0 (T) Markov text: order-2 byte Markov chain whose transition table is built from a corpus
// kind 1 (R) repeated substrings: 4096 printable phrases of length 8..64 drawn Zipf(1.2)
// kind 2 (N) near-random mix: 4 KiB blocks, 90 % uniform random bytes / 10 % Markov text
// kind 3     round-robin of T/R/N in 1 MiB runs of the output buffer
//
// Integer arithmetic only (splitmix64 + table lookups), so the C, CUDA and emulated builds agree bit for bit.
// A "unit" is an independently seeded piece: unit i covers out[offsets[i], offsets[i+1]) with seed base_seed + i.
#pragma once
#include <math.h>
#include <string.h>

#include <vector>

#include "czk_common.cuh"

namespace czk {

#define CZK_SYNTH_MAX_ENTRIES 65536u
#define CZK_SYNTH_PHRASES 4096u

struct SynthModel {
    uint32_t start_ctx;                       // (corpus[0] << 8) | corpus[1]
    uint32_t n_entries;
    uint32_t ctx_index[65536];                // offset << 12 | count   (count <= 256 -> 12 bits are enough; offset < 2^20)
    uint32_t ctx_total[65536];                // total frequency of the context (0 = unseen)
    uint32_t entries[CZK_SYNTH_MAX_ENTRIES];  // cumulative frequency << 8 | byte, ascending inside a context
    uint32_t zipf_cum[CZK_SYNTH_PHRASES];     // cumulative Zipf(1.2) weights scaled to 2^32-1 at the last rank
};

__host__ __device__ inline uint64_t splitmix64(uint64_t &s) {
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

struct SynthGen {
    const SynthModel *m;
    uint64_t rng;
    uint32_t ctx;
    // phrase emission state (kind R)
    uint64_t phrase_rng;
    uint32_t phrase_left;

    __host__ __device__ inline void init(const SynthModel *model, uint64_t seed) {
        m = model;
        rng = seed * 0xd1342543de82ef95ull + 0x2545f4914f6cdd1dull;
        ctx = model->start_ctx;
        phrase_left = 0;
        phrase_rng = 0;
    }
    __host__ __device__ inline uint32_t markov() {
        uint32_t tot = m->ctx_total[ctx];
        if (!tot) { ctx = m->start_ctx; tot = m->ctx_total[ctx]; }
        uint32_t idx = m->ctx_index[ctx];
        uint32_t off = idx >> 12, cnt = idx & 0xfff;
        uint32_t x = (uint32_t)(((splitmix64(rng) >> 32) * (uint64_t)tot) >> 32);  // uniform in [0, tot)
        uint32_t lo = 0, hi = cnt - 1;                                              // first entry with cum > x
        while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if ((m->entries[off + mid] >> 8) > x) hi = mid; else lo = mid + 1;
        }
        uint32_t b = m->entries[off + lo] & 0xff;
        ctx = ((ctx << 8) | b) & 0xffff;
        return b;
    }
    __host__ __device__ inline uint32_t phrase_byte(uint64_t dict_seed) {
        if (!phrase_left) {
            uint32_t u = (uint32_t)(splitmix64(rng) >> 32);
            uint32_t lo = 0, hi = CZK_SYNTH_PHRASES - 1;  // first rank with cum >= u
            while (lo < hi) {
                uint32_t mid = (lo + hi) >> 1;
                if (m->zipf_cum[mid] >= u) hi = mid; else lo = mid + 1;
            }
            phrase_rng = dict_seed ^ (0x9e3779b97f4a7c15ull * (lo + 1));
            phrase_left = 8 + (uint32_t)(splitmix64(phrase_rng) % 57);
        }
        phrase_left--;
        if (!phrase_left) return ' ';
        return 33 + (uint32_t)(splitmix64(phrase_rng) % 94);  // printable, no space
    }
};

// Fills one unit. `abs_off` is the unit's offset inside the whole output buffer (selects the class for kind 3).
__host__ __device__ inline void synth_fill_unit(const SynthModel *m, int kind, uint64_t base_seed, uint64_t unit,
                                                uint8_t *out, uint64_t len, uint64_t abs_off) {
    SynthGen g;
    g.init(m, base_seed + unit);
    const uint64_t dict_seed = base_seed * 0x100000001b3ull + 0x6a09e667f3bcc909ull;
    uint32_t block_is_text = 0;
    const bool packed = ((uintptr_t)out & 7) == 0;  // 8-byte stores when the unit is aligned
    uint64_t acc = 0;
    for (uint64_t i = 0; i < len; i++) {
        int k = kind;
        if (kind == 3) k = (int)(((abs_off + i) >> 20) % 3);
        uint32_t b;
        if (k == 0) b = g.markov();
        else if (k == 1) b = g.phrase_byte(dict_seed);
        else {
            if (((abs_off + i) & 4095) == 0 || i == 0) block_is_text = (splitmix64(g.rng) >> 32) % 10 == 0;
            b = block_is_text ? g.markov() : (uint32_t)(splitmix64(g.rng) >> 56);
        }
        if (packed) {
            acc |= (uint64_t)b << (8 * (i & 7));
            if ((i & 7) == 7) { *(uint64_t *)(out + i - 7) = acc; acc = 0; }
            else if (i + 1 == len) { for (uint64_t kk = i & ~(uint64_t)7; kk <= i; kk++) out[kk] = (uint8_t)(acc >> (8 * (kk & 7))); }
        } else out[i] = (uint8_t)b;
    }
}

// Builds the model from a corpus (host): order-2 successor counts and the Zipf(1.2) table. Returns 0, or -4 when the corpus
// has more distinct order-2 transitions than the table holds. `zipf_pow(r)` = r^-1.2 as a double (host libm).
inline int synth_build_model(const uint8_t *corpus, uint64_t corpus_len, SynthModel *m) {
    if (!corpus || corpus_len < 3 || !m) return -2;
    memset(m, 0, sizeof *m);
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
            if (n >= CZK_SYNTH_MAX_ENTRIES) return -4;
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

#if !defined(CZK_MODEL)
__global__ void __launch_bounds__(128) synth_kernel(const SynthModel *m, int kind, uint64_t base_seed, uint32_t n, uint8_t *out,
                                                    const uint64_t *offsets) {
    uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    uint64_t o0 = offsets[u], o1 = offsets[u + 1];
    synth_fill_unit(m, kind, base_seed, u, out + o0, o1 - o0, o0);
}
#endif

}  // namespace czk
