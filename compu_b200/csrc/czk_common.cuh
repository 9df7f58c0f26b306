// czk_common.cuh — shared definitions for the compu-b200 CUDA kernels.
//
// The kernel sources (*.cuh) compile two ways:
//   * nvcc, -gencode arch=compute_100a,code=sm_100a  -> the product (libcompu_b200.so)
//   * g++ -DCUSIM with tests/cusim/cusim.h            -> a CPU emulation used ONLY by tests to debug kernel logic
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef CUSIM
#include "cusim.h"
#elif defined(CZK_MODEL)
// host-only build of the shared __host__ __device__ decision code (tests/model); no CUDA headers needed
#ifndef __host__
#define __host__
#define __device__
#endif
#else
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#define CZ_DYNAMIC_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#endif

namespace czk {

#define CZK_FULL 0xffffffffu

// status codes shared with include/compu_b200.h
enum : int32_t {
    ST_NEED_INPUT = 0,
    ST_NEED_OUTPUT = 1,
    ST_FINISHED = 2,
    ST_NEED_DICT = 3,
    ST_E_DATA = -3,
    ST_E_MEM = -4,
    ST_E_BUF = -5,
};

// CRC-32 (IEEE 802.3, reflected 0xEDB88320) helper tables, built on the host once (checksum.cpp) and read by kernels.
struct CrcTables {
    uint32_t table[256];   // byte-at-a-time table
    uint32_t pow128[34];   // pow128[k] = x^(8*128*k) mod P (reflected form), k = 0..33
};

#define CZK_CRC_POLY 0xedb88320u

// a(x) * b(x) mod P in the reflected representation (bit 31 = x^0).
__host__ __device__ inline uint32_t crc_mulmod(uint32_t a, uint32_t b) {
    uint32_t p = 0;
#pragma unroll 4
    for (int i = 0; i < 32; i++) {
        // bit (31 - i) of a is the coefficient of x^i
        uint32_t m = 0u - ((a >> (31 - i)) & 1u);
        p ^= b & m;
        b = (b >> 1) ^ (CZK_CRC_POLY & (0u - (b & 1u)));
    }
    return p;
}

// x^(8*nbytes) mod P
__host__ __device__ inline uint32_t crc_xpow8n(uint64_t nbytes) {
    uint32_t p = 0x80000000u;      // 1
    uint32_t sq = 0x00800000u;     // x^8  (bit 31-8)
    while (nbytes) {
        if (nbytes & 1) p = crc_mulmod(sq, p);
        sq = crc_mulmod(sq, sq);
        nbytes >>= 1;
    }
    return p;
}

__host__ __device__ inline uint32_t crc32_combine_u(uint32_t crc1, uint32_t crc2, uint64_t len2) {
    return crc_mulmod(crc_xpow8n(len2), crc1) ^ crc2;
}

inline void init_crc_tables(CrcTables *t) {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (CZK_CRC_POLY & (0u - (c & 1u)));
        t->table[i] = c;
    }
    uint32_t x1024 = crc_xpow8n(128);
    t->pow128[0] = 0x80000000u;
    for (int k = 1; k < 34; k++) t->pow128[k] = crc_mulmod(t->pow128[k - 1], x1024);
}

#define CZK_ADLER_BASE 65521u
__host__ __device__ inline uint32_t adler32_combine_u(uint32_t ad1, uint32_t ad2, uint64_t len2) {
    // A = a1 + a2 - 1 ; B = b1 + b2 + len2*(a1 - 1)   (mod 65521)
    uint32_t a1 = ad1 & 0xffff, b1 = ad1 >> 16, a2 = ad2 & 0xffff, b2 = ad2 >> 16;
    uint32_t rem = (uint32_t)(len2 % CZK_ADLER_BASE);
    uint32_t A = (a1 + a2 + CZK_ADLER_BASE - 1) % CZK_ADLER_BASE;
    uint64_t B = (uint64_t)b1 + b2 + (uint64_t)rem * ((a1 + CZK_ADLER_BASE - 1) % CZK_ADLER_BASE);
    return ((uint32_t)(B % CZK_ADLER_BASE) << 16) | A;
}

}  // namespace czk
