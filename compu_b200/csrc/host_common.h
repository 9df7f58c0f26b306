// host_common.h — host-side plumbing shared by the C-ABI translation units (device contexts, buffers, errors).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/compu_b200.h"
#include "czk_common.cuh"

namespace czh {

void set_error(const char *fmt, ...);
bool cuda_ok(cudaError_t e, const char *what);
#define CZ_CUDA(call) ::czh::cuda_ok((call), #call)

// One per CUDA device, created lazily. Holds what kernels need that is device-resident and read-only.
struct DeviceCtx {
    int dev = -1;
    int sm_count = 0;
    czk::CrcTables *d_crc = nullptr;
    bool ok = false;
};
DeviceCtx *device_ctx(int dev);  // nullptr if the device is unusable (not sm_100, CUDA failure)
int usable_device_count();
int stream_device();  // the device new streaming Decoder / Encoder objects are placed on (cz_set_stream_device, default 0)

// Grow-only device / pinned-host buffers reused across calls (reset() keeps the allocation).
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool reserve(size_t bytes);
    void release();
    ~DevBuf() { release(); }
    template <class T>
    T *as() { return (T *)p; }
};
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool reserve(size_t bytes, bool keep = false, size_t keep_bytes = 0);
    void release();
    ~PinBuf() { release(); }
    template <class T>
    T *as() { return (T *)p; }
};

inline uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

// Bounded per-device pool of reusable work objects (device / pinned buffers, streams). A call borrows one object for its
// duration, so concurrent callers never share buffers, and at most KEEP idle objects per device outlive a call (a
// 32-thread caller does not leave 32 sets of buffers behind; the reference's handles are one-thread-at-a-time but several
// handles may run on several threads, SURVEY.md 8b "Threading").
template <class T, int KEEP = 2>
struct DevicePool {
    std::mutex mu;
    std::vector<T *> idle[64];
    T *acquire(int dev) {
        {
            std::lock_guard<std::mutex> lk(mu);
            std::vector<T *> &v = idle[dev & 63];
            if (!v.empty()) { T *w = v.back(); v.pop_back(); return w; }
        }
        return new (std::nothrow) T();
    }
    void release(int dev, T *w) {
        if (!w) return;
        {
            std::lock_guard<std::mutex> lk(mu);
            std::vector<T *> &v = idle[dev & 63];
            if ((int)v.size() < KEEP) { v.push_back(w); return; }
        }
        delete w;
    }
};

// kernels launched by this library since load (bench.py reports the count of the timed region as gpu_launches)
void count_launches(unsigned k);
bool profiling_on();  // cz_profile_enable (inflate.cu)
#define CZ_KL(...) do { __VA_ARGS__; ::czh::count_launches(1); } while (0)

// kernel launchers (defined in inflate.cu / deflate.cu)
int launch_inflate(cudaStream_t st, DeviceCtx *ctx, size_t n, const uint8_t *d_in, const uint64_t *d_in_off, uint8_t *d_out,
                   const uint64_t *d_out_off, uint64_t *d_out_lens, int32_t *d_statuses, uint64_t *d_in_consumed,
                   uint32_t *d_checks, int window_bits, int segment_mode, int check_kind, void *d_ws, uint64_t ws_bytes,
                   uint64_t total_out_bytes, const uint32_t *d_ids, size_t n_ids, int big);
uint64_t inflate_workspace_bytes(size_t n, uint64_t total_out_bytes);
}  // namespace czh
namespace czk { struct ResumeState; }
namespace czh {
int launch_inflate_resume(cudaStream_t st, DeviceCtx *ctx, const uint8_t *d_in, const uint64_t *d_in_off, uint8_t *d_out,
                          const uint64_t *d_out_off, uint64_t *d_out_lens, int32_t *d_statuses, int window_bits,
                          czk::ResumeState *d_resume, void *d_ws);

// block-parallel path for long streams (inflate.cu, inflate_runs_host.h)
uint64_t runs_min_unit_bytes();
int inflate_long_units(int dev, const std::vector<size_t> &ids, const uint8_t *in, const uint64_t *in_off, uint8_t *out,
                       const uint64_t *out_off, uint64_t *out_lens, int32_t *statuses, uint64_t *in_consumed, int window_bits,
                       uint8_t *done);

// batched inflate over host memory (host.cu); segment_mode / checks as in cz_inflate_segments_device
int inflate_batch_host(size_t n, const uint8_t *in, const uint64_t *in_off, uint8_t *out, const uint64_t *out_off,
                       uint64_t *out_lens, int32_t *statuses, uint64_t *in_consumed, int window_bits, int segment_mode,
                       uint32_t *checks, uint32_t devices_mask);

}  // namespace czh
