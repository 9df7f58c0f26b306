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

__global__ void __launch_bounds__(128) synth_kernel(const SynthModel *m, int kind, uint64_t base_seed, uint32_t n, uint8_t *out,
                                                    const uint64_t *offsets) {
    uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    uint64_t o0 = offsets[u], o1 = offsets[u + 1];
    synth_fill_unit(m, kind, base_seed, u, out + o0, o1 - o0, o0);
}

}  // namespace czk
