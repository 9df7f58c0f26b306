// cusim.h — a tiny CPU emulator of the CUDA execution model, TEST INFRASTRUCTURE ONLY.
//
// Purpose: compile the *same* kernel source that nvcc compiles for sm_100a (compu_b200/csrc/*.cuh) with g++ and run
// it on the CPU, one fiber per CUDA thread, so kernel logic can be debugged and regression-tested in a container
// that has no GPU. It is never linked into the product library; the product has no CPU path.
//
// Model: CTAs run one after another; inside a CTA every thread is a fiber (hand-rolled x86-64 context switch).
// A fiber runs until it reaches a barrier or warp collective, then yields. The scheduler picks fibers in a
// pseudo-random (seeded) order, so code that forgets a __syncwarp()/__syncthreads() between a write and another
// lane's read fails here with high probability instead of "working" by lock-step luck.
//
// Supported: threadIdx/blockIdx/blockDim/gridDim (.x only... .y/.z = 0/1), static and dynamic shared memory,
// __syncthreads, __syncwarp, __shfl*_sync, __ballot_sync, __any/__all_sync, __match_any_sync, __reduce_*_sync,
// atomics (plain ops: one OS thread), bit intrinsics.
#pragma once
#ifndef CUSIM
#error "cusim.h is only for -DCUSIM host builds"
#endif
#if !defined(__x86_64__)
#error "cusim needs x86-64"
#endif

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __constant__ static
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define CZ_DYNAMIC_SMEM(name) uint8_t *name = ::cusim::g_dyn_smem

namespace cusim {

struct dim3_ {
    unsigned x = 1, y = 1, z = 1;
};

struct Warp;
struct Fiber {
    void *sp = nullptr;
    uint8_t *stack = nullptr;
    unsigned tid = 0;
    bool done = false;
    Warp *warp = nullptr;
};

struct BarEntry {
    uint32_t mask, arrived, gen;
};
struct Warp {
    uint64_t xchg[32];
    std::vector<BarEntry> bars;
};

extern Fiber *g_cur;
extern void *g_sched_sp;
extern uint8_t *g_dyn_smem;
extern dim3_ g_blockIdx, g_blockDim, g_gridDim;
extern unsigned g_cta_arrived, g_cta_gen, g_cta_live;
extern uint64_t g_progress;
extern std::function<void()> *g_body;

extern "C" void cusim_switch(void **save_sp, void *load_sp);

inline void yield() { cusim_switch(&g_cur->sp, g_sched_sp); }

inline void warp_barrier(uint32_t mask) {
    Warp &w = *g_cur->warp;
    unsigned lane = g_cur->tid & 31;
    if (!(mask >> lane & 1)) {
        fprintf(stderr, "cusim: lane %u not in its own mask %08x\n", lane, mask);
        abort();
    }
    size_t idx = 0;
    for (; idx < w.bars.size(); idx++)
        if (w.bars[idx].mask == mask) break;
    if (idx == w.bars.size()) w.bars.push_back({mask, 0, 0});
    uint32_t my_gen = w.bars[idx].gen;
    w.bars[idx].arrived |= 1u << lane;
    g_progress++;
    if (w.bars[idx].arrived == mask) {
        w.bars[idx].arrived = 0;
        w.bars[idx].gen++;
        return;
    }
    while (w.bars[idx].gen == my_gen) yield();
}

template <class T>
inline uint64_t to_bits(T v) {
    uint64_t b = 0;
    static_assert(sizeof(T) <= 8, "cusim: collective operand too wide");
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T>
inline T from_bits(uint64_t b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}

// deposit, barrier, compute from all deposits, barrier
template <class T, class F>
inline auto collective(uint32_t mask, T v, F f) -> decltype(f((const uint64_t *)nullptr)) {
    Warp &w = *g_cur->warp;
    unsigned lane = g_cur->tid & 31;
    w.xchg[lane] = to_bits(v);
    warp_barrier(mask);
    auto r = f((const uint64_t *)w.xchg);
    warp_barrier(mask);
    return r;
}

void launch_impl(dim3_ grid, dim3_ block, size_t dyn_smem, std::function<void()> body);
void set_seed(uint64_t s);

template <class K, class... A>
inline void launch(unsigned grid, unsigned block, size_t dyn_smem, K kernel, A... args) {
    dim3_ g, b;
    g.x = grid;
    b.x = block;
    launch_impl(g, b, dyn_smem, [=]() { kernel(args...); });
}

}  // namespace cusim

#define threadIdx (::cusim::dim3_{::cusim::g_cur->tid, 0, 0})
#define blockIdx (::cusim::g_blockIdx)
#define blockDim (::cusim::g_blockDim)
#define gridDim (::cusim::g_gridDim)
#define warpSize 32

inline void __syncwarp(unsigned mask = 0xffffffffu) { ::cusim::warp_barrier(mask); }

inline void __syncthreads() {
    using namespace cusim;
    unsigned my_gen = g_cta_gen;
    g_progress++;
    if (++g_cta_arrived == g_cta_live) {
        g_cta_arrived = 0;
        g_cta_gen++;
        return;
    }
    while (g_cta_gen == my_gen) yield();
}
inline void __threadfence() {}
inline void __threadfence_block() {}

template <class T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    unsigned lane = ::cusim::g_cur->tid & 31;
    unsigned base = lane & ~(unsigned)(width - 1);
    unsigned s = base + ((unsigned)src & (unsigned)(width - 1));
    return ::cusim::collective(mask, v, [=](const uint64_t *x) { return ::cusim::from_bits<T>(x[s]); });
}
template <class T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    unsigned lane = ::cusim::g_cur->tid & 31;
    unsigned base = lane & ~(unsigned)(width - 1);
    int s = (int)lane - (int)delta;
    if (s < (int)base) s = lane;
    return ::cusim::collective(mask, v, [=](const uint64_t *x) { return ::cusim::from_bits<T>(x[s]); });
}
template <class T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    unsigned lane = ::cusim::g_cur->tid & 31;
    unsigned base = lane & ~(unsigned)(width - 1);
    unsigned s = lane + delta;
    if (s >= base + (unsigned)width) s = lane;
    return ::cusim::collective(mask, v, [=](const uint64_t *x) { return ::cusim::from_bits<T>(x[s]); });
}
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    unsigned lane = ::cusim::g_cur->tid & 31;
    unsigned s = lane ^ (unsigned)lanemask;
    (void)width;
    return ::cusim::collective(mask, v, [=](const uint64_t *x) { return ::cusim::from_bits<T>(x[s]); });
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    return ::cusim::collective(mask, (uint32_t)(pred != 0), [=](const uint64_t *x) {
        unsigned r = 0;
        for (int i = 0; i < 32; i++)
            if ((mask >> i & 1) && x[i]) r |= 1u << i;
        return r;
    });
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
template <class T>
inline unsigned __match_any_sync(unsigned mask, T v) {
    uint64_t mine = ::cusim::to_bits(v);
    return ::cusim::collective(mask, v, [=](const uint64_t *x) {
        unsigned r = 0;
        for (int i = 0; i < 32; i++)
            if ((mask >> i & 1) && x[i] == mine) r |= 1u << i;
        return r;
    });
}
#define CUSIM_REDUCE(name, T, init, op)                                       \
    inline T name(unsigned mask, T v) {                                       \
        return ::cusim::collective(mask, v, [=](const uint64_t *x) {          \
            T r = init;                                                       \
            for (int i = 0; i < 32; i++)                                      \
                if (mask >> i & 1) { T e = ::cusim::from_bits<T>(x[i]); r = op; } \
            return r;                                                         \
        });                                                                   \
    }
CUSIM_REDUCE(__reduce_add_sync, unsigned, 0u, r + e)
CUSIM_REDUCE(__reduce_or_sync, unsigned, 0u, r | e)
CUSIM_REDUCE(__reduce_and_sync, unsigned, 0xffffffffu, r &e)
CUSIM_REDUCE(__reduce_xor_sync, unsigned, 0u, r ^ e)
CUSIM_REDUCE(__reduce_max_sync, unsigned, 0u, (r > e ? r : e))
CUSIM_REDUCE(__reduce_min_sync, unsigned, 0xffffffffu, (r < e ? r : e))
inline int __reduce_add_sync(unsigned mask, int v) { return (int)__reduce_add_sync(mask, (unsigned)v); }
inline int __reduce_max_sync(unsigned mask, int v) {
    return (int)(__reduce_max_sync(mask, (unsigned)v ^ 0x80000000u) ^ 0x80000000u);
}
inline int __reduce_min_sync(unsigned mask, int v) {
    return (int)(__reduce_min_sync(mask, (unsigned)v ^ 0x80000000u) ^ 0x80000000u);
}

// ---- atomics: one OS thread, fibers only switch at barriers => plain read-modify-write is atomic
template <class T, class U>
inline T atomicAdd(T *p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U>
inline T atomicSub(T *p, U v) { T o = *p; *p = (T)(o - (T)v); return o; }
template <class T, class U>
inline T atomicOr(T *p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U>
inline T atomicAnd(T *p, U v) { T o = *p; *p = (T)(o & (T)v); return o; }
template <class T, class U>
inline T atomicXor(T *p, U v) { T o = *p; *p = (T)(o ^ (T)v); return o; }
template <class T, class U>
inline T atomicMax(T *p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U>
inline T atomicMin(T *p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U>
inline T atomicExch(T *p, U v) { T o = *p; *p = (T)v; return o; }
template <class T, class U, class V>
inline T atomicCAS(T *p, U cmp, V v) { T o = *p; if (o == (T)cmp) *p = (T)v; return o; }

// ---- bit intrinsics
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline unsigned __brev(unsigned x) {
    x = (x >> 16) | (x << 16);
    x = ((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8);
    x = ((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2);
    x = ((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1);
    return x;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)(v >> (sh & 31));
}
inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned sh) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return sh >= 32 ? hi : (unsigned)(v >> sh);
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)((v << (sh & 31)) >> 32);
}
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    uint64_t v = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 0xf;
        unsigned b = (unsigned)(v >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) b = (b & 0x80) ? 0xff : 0;
        r |= b << (8 * i);
    }
    return r;
}
inline float __fdividef(float a, float b) { return a / b; }
inline unsigned __float2uint_rz(float x) { return (unsigned)x; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
template <class T>
inline T __ldg(const T *p) { return *p; }
template <typename T>
inline T __ldcs(const T *p) { return *p; }
template <typename T, typename U>
inline void __stcs(T *p, U v) { *p = (T)v; }
inline unsigned __vcmpeq4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int i = 0; i < 4; i++)
        if (((a >> (8 * i)) & 0xff) == ((b >> (8 * i)) & 0xff)) r |= 0xffu << (8 * i);
    return r;
}
using std::max;
using std::min;
inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
inline unsigned min(int a, unsigned b) { return (unsigned)a < b ? (unsigned)a : b; }
inline unsigned max(unsigned a, int b) { return a > (unsigned)b ? a : (unsigned)b; }
inline unsigned max(int a, unsigned b) { return (unsigned)a > b ? (unsigned)a : b; }
inline unsigned long long min(unsigned long long a, unsigned b) { return a < b ? a : b; }
inline unsigned long long min(unsigned a, unsigned long long b) { return a < b ? a : b; }

struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
