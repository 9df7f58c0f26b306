/*
 * compu_b200.h — C ABI of the B200-native DEFLATE-family backend for compu's `Interface` vtables.
 *
 * This is the drop-in boundary: every entry point below is what a Rust adapter of the
 * shape of /root/reference/src/decoder/zlib_ng.rs and src/encoder/zlib_ng.rs binds with `extern "C"`.
 * Plain pointers and sizes only; no torch, no C++ types. The implementing library
 * (compu_b200/libcompu_b200.so) is CUDA-only: with no usable sm_100 device every constructor returns
 * NULL and every batch call returns CZ_E_NO_DEVICE — there is no CPU fallback.
 *
 * Reference interfaces replaced (file:line under /root/reference):
 *   cz_decoder_new      <- decoder::Interface::zlib_ng(mode)            src/decoder/zlib_ng.rs:61-90
 *   cz_decode           <- decode_fn / internal_zlib_impl_decode!       src/decoder/zlib_ng.rs:94-96, src/decoder/mod.rs:459-486
 *   cz_decoder_reset    <- reset_fn (returned pointer replaces old)     src/decoder/zlib_ng.rs:99-108, src/decoder/mod.rs:433-441
 *   cz_decoder_free     <- drop_fn                                      src/decoder/zlib_ng.rs:111-115
 *   cz_describe_error   <- describe_error_fn (zError, 'static text)     src/decoder/zlib_ng.rs:118-123
 *   cz_encoder_new      <- encoder::Interface::zlib_ng(opts)            src/encoder/zlib_ng.rs:50-87
 *   cz_encode           <- encode_fn / internal_zlib_impl_encode!       src/encoder/zlib_ng.rs:90-92, src/encoder/mod.rs:334-370
 *   cz_encoder_reset    <- reset_fn                                     src/encoder/zlib_ng.rs:95-104
 *   cz_encoder_free     <- drop_fn                                      src/encoder/zlib_ng.rs:107-111
 *   cz_host_alloc/free  <- compu_malloc / compu_free (pinned variant)   src/mem.rs:27-49
 *   cz_*_batch*         <- NEW: the batched many-stream entry points named by BASELINE.json north_star
 */
#ifndef COMPU_B200_H
#define COMPU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors `Decode` (src/decoder/mod.rs:150-157) and `Encode` (src/encoder/mod.rs:42-49). */
typedef struct cz_result {
    size_t input_remain;  /* bytes of THIS call's input not consumed  */
    size_t output_remain; /* bytes of THIS call's output not written  */
    int32_t status;       /* see enums below                          */
} cz_result;

/* DecodeStatus (src/decoder/mod.rs:139-146); negative values are DecodeError(i32) with zlib numbering. */
enum {
    CZ_DECODE_NEED_INPUT = 0,
    CZ_DECODE_NEED_OUTPUT = 1,
    CZ_DECODE_FINISHED = 2,
    CZ_DECODE_NEED_DICT = 3, /* error: the adapter returns Err(DecodeError(2)) (zlib's Z_NEED_DICT; 2 is taken by Finished here) */
    CZ_E_STREAM = -2,         /* Z_STREAM_ERROR  */
    CZ_E_DATA = -3,           /* Z_DATA_ERROR    */
    CZ_E_MEM = -4,            /* Z_MEM_ERROR (also: CUDA failure / out of device memory) */
    CZ_E_BUF = -5,            /* Z_BUF_ERROR     */
    CZ_E_VERSION = -6,        /* Z_VERSION_ERROR */
    CZ_E_NO_DEVICE = -100     /* batch calls only: no usable sm_100 CUDA device */
};

/* EncodeStatus (src/encoder/mod.rs:27-38). */
enum { CZ_ENCODE_CONTINUE = 0, CZ_ENCODE_NEED_OUTPUT = 1, CZ_ENCODE_FINISHED = 2, CZ_ENCODE_ERROR = 3 };

/* EncodeOp (src/encoder/mod.rs:12-23). Flush is a sync-flush point, as in the reference (mod.rs:338). */
enum { CZ_OP_PROCESS = 0, CZ_OP_FLUSH = 1, CZ_OP_FINISH = 2 };

/* ZlibMode: window_bits = -15 raw deflate, 15 zlib, 31 gzip, 47 auto (decoder only)
   (src/decoder/zlib_common.rs:4-15, src/encoder/zlib_common.rs:28-37). */
/* ZlibStrategy (src/encoder/zlib_common.rs:5-16, mapped zlib_ng.rs:70-76). */
enum { CZ_STRATEGY_DEFAULT = 0, CZ_STRATEGY_FILTERED = 1, CZ_STRATEGY_HUFFMAN_ONLY = 2, CZ_STRATEGY_RLE = 3, CZ_STRATEGY_FIXED = 4 };

/* ---------------------------------------------------------------- library / device ---------------- */

/* Number of usable sm_100 devices (0 = none; product calls then fail loudly). */
int cz_device_count(void);
/* Library version string, static storage. */
const char *cz_version(void);
/* Text of the last CUDA/host failure on the calling thread ("" if none), static/thread storage. */
const char *cz_last_error(void);

/* Kernels launched by this library since it was loaded (all threads, all devices). bench.py reports the difference over
   its timed region as "gpu_launches". */
uint64_t cz_launch_count(void);

/* Pinned (page-locked) host memory for zero-staging transfers; the pinned analogue of compu_malloc/compu_free. */
void *cz_host_alloc(size_t bytes);
void cz_host_free(void *p);

/* ---------------------------------------------------------------- streaming decoder ---------------- */

/* The device on which cz_decoder_new / cz_encoder_new place the objects they create from now on (process-wide, default 0; an
 * object stays on the device it was created on). compu has no counterpart: its backends run where the caller's thread runs
 * (the rayon baseline creates one handle per worker, src/decoder/mod.rs:196-204); a worker bound to GPU k calls this once.
 * Returns 0, or CZ_E_NO_DEVICE. */
int cz_set_stream_device(int device);

void *cz_decoder_new(int window_bits); /* NULL on failure => Interface::zlib_cuda returns None */
cz_result cz_decode(void *state, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len);
void *cz_decoder_reset(void *state); /* instance to use from now on; NULL = failed */
void cz_decoder_free(void *state);
const char *cz_describe_error(int32_t code); /* static storage; non-NULL for code 0 */

/* ---------------------------------------------------------------- streaming encoder ---------------- */

/* level: -1 (default = 6), 0..9. mem_level accepted for signature parity (1..9), strategy: CZ_STRATEGY_*. */
void *cz_encoder_new(int level, int window_bits, int mem_level, int strategy);
cz_result cz_encode(void *state, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len, int op);
void *cz_encoder_reset(void *state);
void cz_encoder_free(void *state);

/* ---------------------------------------------------------------- batched entry points (host memory) */

/* Inflate n independent streams. Stream i is in[in_offsets[i] .. in_offsets[i+1]) and is written to
   out[out_offsets[i] .. out_offsets[i+1]) (the slot size is the capacity). Per stream: out_lens[i] = bytes
   produced, statuses[i] = CZ_DECODE_FINISHED | CZ_DECODE_NEED_INPUT (truncated) | CZ_DECODE_NEED_OUTPUT (slot too
   small) | <0 error. in_consumed may be NULL. `in`/`out` may be pageable or cz_host_alloc'ed (pinned: no staging).
   devices_mask: bit d = use CUDA device d; 0 = device 0 only. Streams are sharded over the devices by
   compressed size with no collective; the gather is host-side. Returns 0 or a negative CZ_E_* for whole-call failure. */
int cz_inflate_batch(size_t n, const uint8_t *in, const uint64_t *in_offsets, uint8_t *out, const uint64_t *out_offsets,
                     uint64_t *out_lens, int32_t *statuses, uint64_t *in_consumed, int window_bits, uint32_t devices_mask);

/* Pointer-array form (the shape proposed in SURVEY.md §8b); gathers into the packed form above. */
int cz_inflate_batch_ptrs(size_t n, const uint8_t *const *in_ptrs, const size_t *in_lens, uint8_t *const *out_ptrs,
                          const size_t *out_caps, size_t *out_lens, int32_t *statuses, int window_bits,
                          uint32_t devices_mask);

/* Worst-case output bytes for deflating `len` input bytes into one stream with the given container and segment size
   (segment_bytes 0 = library default). */
uint64_t cz_deflate_bound(uint64_t len, int window_bits, uint64_t segment_bytes);

/* Deflate n independent buffers, each into its own complete stream (container by window_bits). Slot i of `out` must
   hold cz_deflate_bound(len_i, ...) bytes to be safe; statuses[i] = CZ_ENCODE_FINISHED | CZ_ENCODE_NEED_OUTPUT | CZ_ENCODE_ERROR. */
int cz_deflate_batch(size_t n, const uint8_t *in, const uint64_t *in_offsets, uint8_t *out, const uint64_t *out_offsets,
                     uint64_t *out_lens, int32_t *statuses, int level, int window_bits, int strategy,
                     uint64_t segment_bytes, uint32_t devices_mask);

/* Deflate ONE buffer into ONE valid stream made of independent full-flush segments (SURVEY.md §8e recipe).
   seg_index (optional, capacity seg_index_cap entries): compressed byte offset of each segment's start inside `out`,
   followed by one terminal entry (offset of the final `03 00` block) — the side index for segment-parallel inflate.
   *n_segments receives the number of segments. Returns 0, or CZ_E_BUF if cap is too small, or another CZ_E_*. */
int cz_deflate_segmented(const uint8_t *in, uint64_t len, uint8_t *out, uint64_t cap, uint64_t *out_len, int level,
                         int window_bits, int strategy, uint64_t segment_bytes, uint32_t devices_mask,
                         uint64_t *seg_index, uint64_t seg_index_cap, uint64_t *n_segments);

/* Inflate ONE stream produced by cz_deflate_segmented using its side index: segments are decoded in parallel.
   seg_index has n_segments+1 entries as written by cz_deflate_segmented; segment_bytes is the value used to encode. */
int cz_inflate_segmented(const uint8_t *in, uint64_t len, uint8_t *out, uint64_t cap, uint64_t *out_len, int window_bits,
                         uint64_t segment_bytes, const uint64_t *seg_index, uint64_t n_segments, uint32_t devices_mask);

/* ---------------------------------------------------------------- batched entry points (device memory) */
/* Same contracts, all pointers are device pointers on the CURRENT CUDA device, work is enqueued on `cuda_stream`
   (a cudaStream_t passed as void*) and NOT synchronised. These are what bench.py times for the roofline figure. */

/* 1 when the library was built with -DCZ_EXPERIMENTS (make EXPERIMENTS=1): the kernel variants that measured slower than the
   defaults are compiled in and the tuning knobs below can select them. The default build holds only the product kernels;
   there the knobs accept the default configurations only and return CZ_E_STREAM for anything else. */
int cz_has_experiments(void);

/* Tuning knob (experiments): selects an inflate kernel variant; see compu_b200/csrc/inflate.cu. Default from CZ_INFLATE_CFG
   or -2,14 (two-phase: lane-per-stream token decode with 14 warps per SM, then warp-per-stream LZ77 resolution). */
int cz_tune_inflate(int slots_per_warp, int warps_per_cta);
/* Tuning knob (experiments): phase B of the two-phase inflate for units whose output slot is at most 64 KiB.
   cta_mode 0: warp per unit, output in global memory (inflate_lz_kernel); 1 / 2: one CTA of 8 / 4 warps per unit, output
   assembled in a shared-memory tile (inflate_lz_cta_kernel); spin_ns: back-off of a warp that found no ready token.
   Defaults from CZ_LZ_CTA / CZ_LZ_SPIN_NS. */
int cz_tune_inflate_lz(int cta_mode, int spin_ns);

/* Scratch bytes cz_inflate_batch_device needs for n streams whose output slots total total_out_bytes
   (= d_out_offsets[n] - d_out_offsets[0]; the token area of the two-phase kernel is 4 bytes per output byte). */
uint64_t cz_inflate_workspace_bytes(size_t n, uint64_t total_out_bytes);
int cz_inflate_batch_device(void *cuda_stream, size_t n, const uint8_t *d_in, const uint64_t *d_in_offsets, uint8_t *d_out,
                            const uint64_t *d_out_offsets, uint64_t total_out_bytes, uint64_t *d_out_lens, int32_t *d_statuses,
                            uint64_t *d_in_consumed, int window_bits, void *d_workspace, uint64_t workspace_bytes);

/* Raw-segment form used for segment-parallel inflate: every unit is a raw-deflate fragment that never sets BFINAL and
   ends exactly at the end of its input (what a full-flush segment is); finishing at end-of-input is success.
   d_checks (optional) receives per unit {adler32, crc32} of the produced bytes as two uint32. */
int cz_inflate_segments_device(void *cuda_stream, size_t n, const uint8_t *d_in, const uint64_t *d_in_offsets, uint8_t *d_out,
                               const uint64_t *d_out_offsets, uint64_t total_out_bytes, uint64_t *d_out_lens,
                               int32_t *d_statuses, uint32_t *d_checks, void *d_workspace, uint64_t workspace_bytes);

/* Measurement hook: with profiling enabled every two-phase inflate launch records CUDA events around its two kernels on the
   launching stream; cz_profile_read sums the durations (ms) of inflate_tok_kernel / inflate_lz_kernel over the launches since
   the last read, waits for them, and returns the number of launches. */
void cz_profile_enable(int on);
int cz_profile_read(double *ms_decode, double *ms_resolve);
/* Same for the deflate kernel chain: summed durations (ms) of the match-search kernel(s) and of the whole chain of every
   cz_deflate_* launch since the last read. */
int cz_profile_read_deflate(double *ms_match, double *ms_chain);

/* Deflate on device. Unit i = d_in[in_offsets[i]..in_offsets[i+1]) is compressed as ONE raw full-flush segment (no header,
   no BFINAL, ends byte-aligned with 00 00 ff ff) into d_out[out_offsets[i]..]; d_out_lens[i] = bytes written (or needed, with
   d_statuses[i] = CZ_ENCODE_NEED_OUTPUT, when the slot is smaller), d_checks[2i], d_checks[2i+1] = adler32, crc32 of the
   unit's input (d_checks may be NULL). Units must be <= cz_deflate_max_segment() bytes; total_in_bytes =
   in_offsets[n] - in_offsets[0] (the caller knows it; it sizes the workspace). */
uint64_t cz_deflate_max_segment(void);
uint64_t cz_deflate_segment_bound(uint64_t seg_len);
uint64_t cz_deflate_workspace_bytes(size_t n_segments, uint64_t total_in_bytes);
int cz_deflate_segments_device(void *cuda_stream, size_t n, const uint8_t *d_in, const uint64_t *d_in_offsets,
                               uint64_t total_in_bytes, uint8_t *d_out, const uint64_t *d_out_offsets, uint64_t *d_out_lens,
                               int32_t *d_statuses, uint32_t *d_checks, int level, int strategy, void *d_workspace,
                               uint64_t workspace_bytes);

/* Checksum combine over k segments on the host side of the gather (O(k) scalar work):
   folds per-segment {adler32, crc32, len} into whole-stream values. */
uint32_t cz_adler32_combine(uint32_t adler1, uint32_t adler2, uint64_t len2);
uint32_t cz_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2);

/* Long streams (>= 128 KiB compressed) given to the inflate entry points are cut into runs of blocks — at block headers found
   by a candidate search, whoever produced the stream, flush points or not — and the runs are decoded in parallel
   (compu_b200/csrc/inflate_runs.cuh); what that path cannot prove correct (data errors, truncation, slots that are too small)
   takes the serial kernels, which reproduce zlib's partial output and status. Counters since load: long units seen / long
   units decoded by the parallel path. */
void cz_split_stats(uint64_t *tried, uint64_t *split);

/* The multi-GPU partitioner used by the batched entry points: cuts n packed units (offsets[n+1]) into `parts`
   contiguous ranges balanced by bytes, no collective. cuts[parts+1] receives unit indices, cuts[0] = 0, cuts[parts] = n.
   Pure host arithmetic (no device needed), exported so that callers sharding over processes use the same cuts. */
int cz_partition_by_bytes(size_t n, const uint64_t *offsets, int parts, uint64_t *cuts);

/* Synthetic data with controlled entropy, generated on the current device. kind: 0 Markov text (needs
   d_model from cz_synth_model_bytes/cz_synth_build_model), 1 repeated-substring Zipf, 2 near-random mix, 3 round-robin
   of the three in 1 MiB runs. Unit i fills d_out[offsets[i]..offsets[i+1]) from seed base_seed + i. */
uint64_t cz_synth_model_bytes(void);
int cz_synth_build_model(const uint8_t *corpus, uint64_t corpus_len, uint8_t *model_out /* host, cz_synth_model_bytes() */);
int cz_synth_fill_device(void *cuda_stream, int kind, uint64_t base_seed, size_t n, uint8_t *d_out, const uint64_t *d_offsets,
                         const uint8_t *d_model);
/* Same generator on the host (bit-identical bytes), for CPU-side tests and baselines. */
int cz_synth_fill_host(int kind, uint64_t base_seed, size_t n, uint8_t *out, const uint64_t *offsets, const uint8_t *model);

#ifdef __cplusplus
}
#endif
#endif /* COMPU_B200_H */
