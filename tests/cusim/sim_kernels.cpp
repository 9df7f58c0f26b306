// sim_kernels.cpp — runs the product's kernel sources on the CPU emulator (tests only; see cusim.h).
#include "cusim.h"
#include "../../compu_b200/csrc/inflate_kernel.cuh"
#include "../../compu_b200/csrc/inflate_lane_kernel.cuh"
#include "../../compu_b200/csrc/inflate_lc_kernel.cuh"

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
    P.in = in; P.in_off = in_off; P.out = out; P.out_off = out_off; P.out_lens = out_lens; P.statuses = statuses;
    P.in_consumed = in_consumed; P.checks = checks; P.counter = &counter; P.crc = &crc; P.n = (uint32_t)n;
    P.window_bits = window_bits; P.segment_mode = segment_mode; P.check_kind = check_kind;
    cusim::set_seed(seed);
    switch (D) {
        case 1: run_inflate<1, 2>(P, grid); break;
        case 2: run_inflate<2, 2>(P, grid); break;
        case 4: run_inflate<4, 2>(P, grid); break;
        case 8: run_inflate<8, 1>(P, grid); break;
        case 32: run_inflate<32, 1>(P, grid); break;
        case -1: cusim::launch(grid, 2 * 32, inflate_lc_smem_bytes<2>(), inflate_lc_kernel<2>, P); break;
        case -9: cusim::launch(grid, 2 * 32, inflate_lane_smem_bytes<9, 8, 2>(), inflate_lane_kernel<9, 8, 2>, P); break;
        case -8: cusim::launch(grid, 1 * 32, inflate_lane_smem_bytes<8, 7, 1>(), inflate_lane_kernel<8, 7, 1>, P); break;
        default: return -1;
    }
    return 0;
}

extern "C" uint32_t sim_crc32_combine(uint32_t a, uint32_t b, uint64_t len2) { return crc32_combine_u(a, b, len2); }
extern "C" uint32_t sim_adler32_combine(uint32_t a, uint32_t b, uint64_t len2) { return adler32_combine_u(a, b, len2); }
