// lut_probe.cu — feasibility probe (development tool, not part of the product): how fast is a lane-per-stream decode loop whose
// per-lane Huffman lookup tables live in GLOBAL memory (L2-resident) instead of canonical structures in shared memory?
// Mimics phase A's access pattern on cfg2: 2 048 warps, every lane walks its own dependent chain of `iters` steps; a step is one
// table lookup in the lane's root table (2^R 16-bit entries), with probability ~pm a second dependent lookup in the lane's
// distance table (2^DR entries), with probability ~pl a third one (sub-table); every step stores one 4-byte token to the lane's
// own token stream and every other step loads the next 4-byte word of the lane's own input stream (prefetched two words ahead).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/lut_probe tools/lut_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int R, int DR>
__global__ void __launch_bounds__(448) probe(const uint16_t *__restrict__ lut, size_t lane_stride, const uint32_t *__restrict__ in,
                                             size_t in_stride, uint32_t *__restrict__ tok, size_t tok_stride, int iters, uint32_t pm,
                                             uint32_t pl, uint32_t *sink, int lanes_total) {
    const int gl = blockIdx.x * blockDim.x + threadIdx.x;
    if (gl >= lanes_total) return;
    const uint16_t *my = lut + (size_t)gl * lane_stride;
    const uint32_t *w = in + (size_t)gl * in_stride;
    uint32_t *tp = tok + (size_t)gl * tok_stride;
    uint64_t buf = w[0] | ((uint64_t)w[1] << 32);
    uint32_t cnt = 64, widx = 2, nextw = w[2], nextw2 = w[3];
    uint32_t acc = 0;
    for (int i = 0; i < iters; i++) {
        if (cnt <= 32) {
            buf |= (uint64_t)nextw << cnt;
            cnt += 32;
            widx++;
            nextw = nextw2;
            nextw2 = __ldg(w + widx + 1);
        }
        uint32_t e = my[(uint32_t)buf & ((1u << R) - 1u)];
        uint32_t cl = (e & 7u) + 4u;
        buf >>= cl; cnt -= cl;
        if ((e >> 3) % 1000u < pl) {  // long code: sub-table
            e = my[(1u << R) + (1u << DR) + ((uint32_t)buf & 63u)];
            buf >>= 2; cnt -= 2;
        }
        uint32_t t = e;
        if ((e >> 3) % 100u < pm) {  // match: distance table
            uint32_t eb = (e >> 6) & 3u;
            buf >>= eb; cnt -= eb;
            if (cnt <= 32) {
                buf |= (uint64_t)nextw << cnt;
                cnt += 32;
                widx++;
                nextw = nextw2;
                nextw2 = __ldg(w + widx + 1);
            }
            uint32_t d = my[(1u << R) + ((uint32_t)buf & ((1u << DR) - 1u))];
            uint32_t dl = (d & 7u) + 3u;
            buf >>= dl; cnt -= dl;
            t = e | (d << 16);
        }
        *tp++ = t;
        acc += t;
    }
    if (acc == 0x12345u) sink[0] = acc;
}

int main(int argc, char **argv) {
    const int lanes = 65536, iters = argc > 1 ? atoi(argv[1]) : 18200;
    const int R = argc > 2 ? atoi(argv[2]) : 9;
    const uint32_t pm = argc > 3 ? atoi(argv[3]) : 63, pl = argc > 4 ? atoi(argv[4]) : 10;
    const size_t lane_stride = argc > 5 ? atoi(argv[5]) : 3072;  // 16-bit entries per lane (allocated, not all touched)
    const size_t in_stride = 8192 + 64, tok_stride = iters + 8;
    uint16_t *lut; uint32_t *in, *tok, *sink;
    cudaMalloc(&lut, lanes * lane_stride * 2);
    cudaMalloc(&in, lanes * in_stride * 4);
    cudaMalloc(&tok, lanes * tok_stride * 4);
    cudaMalloc(&sink, 4);
    // random contents
    {
        size_t n = lanes * lane_stride;
        uint16_t *h = (uint16_t *)malloc(n * 2);
        uint32_t s = 12345;
        for (size_t i = 0; i < n; i++) { s = s * 1664525u + 1013904223u; h[i] = (uint16_t)(s >> 13); }
        cudaMemcpy(lut, h, n * 2, cudaMemcpyHostToDevice);
        free(h);
        n = lanes * in_stride;
        uint32_t *hi = (uint32_t *)malloc(n * 4);
        for (size_t i = 0; i < n; i++) { s = s * 1664525u + 1013904223u; hi[i] = s ^ (s >> 15); }
        cudaMemcpy(in, hi, n * 4, cudaMemcpyHostToDevice);
        free(hi);
    }
    const int smem = argc > 6 ? atoi(argv[6]) : 60000;  // dynamic shared memory per CTA (shrinks L1 like the real kernel's scratch)
    cudaFuncSetAttribute(probe<8, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220000);
    cudaFuncSetAttribute(probe<9, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220000);
    cudaFuncSetAttribute(probe<10, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220000);
    cudaFuncSetAttribute(probe<11, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220000);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a);
        const int grid = (lanes + 447) / 448;
        if (R == 8) probe<8, 6><<<grid, 448, smem>>>(lut, lane_stride, in, in_stride, tok, tok_stride, iters, pm, pl, sink, lanes);
        else if (R == 9) probe<9, 7><<<grid, 448, smem>>>(lut, lane_stride, in, in_stride, tok, tok_stride, iters, pm, pl, sink, lanes);
        else if (R == 10) probe<10, 7><<<grid, 448, smem>>>(lut, lane_stride, in, in_stride, tok, tok_stride, iters, pm, pl, sink, lanes);
        else probe<11, 8><<<grid, 448, smem>>>(lut, lane_stride, in, in_stride, tok, tok_stride, iters, pm, pl, sink, lanes);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        printf("R=%d pm=%u%% pl=%u/1000 stride=%zu iters=%d: %.2f ms  (%.0f cycles per step at 1.965 GHz) err=%d\n", R, pm, pl, lane_stride, iters,
               ms, ms * 1.965e6 / iters, (int)cudaGetLastError());
    }
    return 0;
}
