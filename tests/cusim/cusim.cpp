// cusim.cpp — scheduler and context switch of the CPU CUDA emulator (test infrastructure only; see cusim.h).
#include "cusim.h"

#include <sys/mman.h>

namespace cusim {

Fiber *g_cur = nullptr;
void *g_sched_sp = nullptr;
uint8_t *g_dyn_smem = nullptr;
dim3_ g_blockIdx, g_blockDim, g_gridDim;
unsigned g_cta_arrived = 0, g_cta_gen = 0, g_cta_live = 0;
uint64_t g_progress = 0;
std::function<void()> *g_body = nullptr;
static uint64_t g_rng = 0x9e3779b97f4a7c15ull;

void set_seed(uint64_t s) { g_rng = s * 0x9e3779b97f4a7c15ull + 0x1234567ull; }

static inline uint64_t rnd() {
    g_rng ^= g_rng << 13;
    g_rng ^= g_rng >> 7;
    g_rng ^= g_rng << 17;
    return g_rng;
}

// void cusim_switch(void **save_sp, void *load_sp)
asm(R"(
.text
.globl cusim_switch
.type cusim_switch,@function
cusim_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size cusim_switch,.-cusim_switch
)");

static void fiber_entry() {
    (*g_body)();
    g_cur->done = true;
    g_progress++;
    // a finished thread no longer takes part in __syncthreads (CUDA: exited threads count as arrived)
    g_cta_live--;
    if (g_cta_live && g_cta_arrived == g_cta_live) {
        g_cta_arrived = 0;
        g_cta_gen++;
    }
    cusim_switch(&g_cur->sp, g_sched_sp);
    abort();
}

static const size_t kStack = 256 * 1024;

void launch_impl(dim3_ grid, dim3_ block, size_t dyn_smem, std::function<void()> body) {
    unsigned nthreads = block.x;
    unsigned nwarps = (nthreads + 31) / 32;
    std::vector<Fiber> fibers(nthreads);
    std::vector<Warp> warps(nwarps);
    uint8_t *stacks = (uint8_t *)mmap(nullptr, kStack * nthreads, PROT_READ | PROT_WRITE,
                                      MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (stacks == MAP_FAILED) {
        perror("cusim mmap");
        abort();
    }
    std::vector<uint8_t> smem(dyn_smem + 64);
    uint8_t *smem_aligned = (uint8_t *)(((uintptr_t)smem.data() + 63) & ~(uintptr_t)63);
    g_body = &body;
    g_gridDim = grid;
    g_blockDim = block;
    std::vector<unsigned> order(nthreads);
    for (unsigned cta = 0; cta < grid.x; cta++) {
        g_blockIdx = dim3_{cta, 0, 0};
        g_dyn_smem = smem_aligned;
        memset(smem_aligned, 0xA5, dyn_smem);  // shared memory is NOT zero on a GPU
        g_cta_arrived = 0;
        g_cta_gen = 0;
        g_cta_live = nthreads;
        for (auto &w : warps) w.bars.clear();
        for (unsigned t = 0; t < nthreads; t++) {
            Fiber &f = fibers[t];
            f.tid = t;
            f.done = false;
            f.warp = &warps[t / 32];
            f.stack = stacks + (size_t)t * kStack;
            uintptr_t top = ((uintptr_t)f.stack + kStack) & ~(uintptr_t)15;
            uint64_t *sp = (uint64_t *)(top - 64);
            for (int i = 0; i < 6; i++) sp[i] = 0;
            sp[6] = (uint64_t)(uintptr_t)&fiber_entry;
            sp[7] = 0;
            f.sp = sp;
            order[t] = t;
        }
        unsigned remaining = nthreads;
        uint64_t idle_rounds = 0;
        while (remaining) {
            // pseudo-random order each round
            for (unsigned i = nthreads - 1; i > 0; i--) {
                unsigned j = (unsigned)(rnd() % (i + 1));
                std::swap(order[i], order[j]);
            }
            uint64_t before = g_progress;
            for (unsigned i = 0; i < nthreads; i++) {
                Fiber &f = fibers[order[i]];
                if (f.done) continue;
                g_cur = &f;
                cusim_switch(&g_sched_sp, f.sp);
                if (f.done) remaining--;
            }
            if (g_progress == before) {
                if (++idle_rounds > 4) {
                    fprintf(stderr, "cusim: deadlock in CTA %u (%u threads unfinished); barrier masks per warp:\n", cta,
                            remaining);
                    for (unsigned w = 0; w < nwarps; w++)
                        for (auto &b : warps[w].bars)
                            if (b.arrived) fprintf(stderr, "  warp %u mask %08x arrived %08x\n", w, b.mask, b.arrived);
                    fprintf(stderr, "  __syncthreads arrived %u of %u\n", g_cta_arrived, g_cta_live);
                    abort();
                }
            } else
                idle_rounds = 0;
        }
    }
    g_cur = nullptr;
    munmap(stacks, kStack * nthreads);
}

}  // namespace cusim
