/*
 * compu_oracle.c — CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of compu's DEFLATE-family hot path: the glue that the reference's
 * `zlib` / `zlib-ng` / `zlib-rust` backends share, driven over madler zlib (the L0 of compu's
 * `zlib` feature backend, the only L0 present in this image — SURVEY.md §8c).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library. The product (compu_b200/) never links, imports or calls it.
 *
 * Parity pin: checked by tests/test_oracle.py against the reference's own golden vectors
 * (the .compressed.gz files under tests/golden, used by /root/reference/tests/decoder.rs:12-15,141-150) through the
 * reference's 5-step decode protocol (tests/decoder.rs:21-77) and its encoder protocols
 * (tests/encoder.rs:10-78, 115-173), and against the zlib 1.3 known answers of SURVEY.md §8c.
 *
 * What follows what (reference file:line):
 *   oz_decoder_new      <- src/decoder/zlib.rs:59-90  / zlib_ng.rs:61-90   (inflateInit2_(strm, mode.max_bits()))
 *   oz_decode           <- src/decoder/mod.rs:459-486 (internal_zlib_impl_decode!: inflate(strm, 0) + status map)
 *   oz_decoder_reset    <- src/decoder/zlib_ng.rs:41-45, 99-108 (inflateReset; same pointer back)
 *   oz_decoder_free     <- src/decoder/zlib_ng.rs:48-55, 111-115 (inflateEnd + free)
 *   oz_describe_error   <- src/decoder/zlib_ng.rs:118-123 (zError)
 *   oz_encoder_new      <- src/encoder/zlib_ng.rs:50-87 (deflateInit2_(level, Z_DEFLATED, mode, mem_level, strategy))
 *   oz_encode           <- src/encoder/mod.rs:334-370 (internal_zlib_impl_encode!: op map, deflate(strm, op), status map)
 *   oz_encoder_reset    <- src/encoder/zlib_ng.rs:30-35, 95-104 (deflateReset)
 *   oz_encoder_free     <- src/encoder/zlib_ng.rs:38-43, 107-111 (deflateEnd + free)
 *   oz_*_batch          <- the "rayon par_iter over streams" CPU baseline of BASELINE.md §5: one reusable
 *                          z_stream per thread with *Reset between items (Decoder::reset, src/decoder/mod.rs:433-441).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct oz_result {
    size_t input_remain;
    size_t output_remain;
    int32_t status; /* decoder: 0 NeedInput, 1 NeedOutput, 2 Finished, <0 zlib error code
                       encoder: 0 Continue, 1 NeedOutput, 2 Finished, 3 Error */
} oz_result;

enum { OZ_NEED_INPUT = 0, OZ_NEED_OUTPUT = 1, OZ_FINISHED = 2 };
enum { OZ_CONTINUE = 0, OZ_ENC_NEED_OUTPUT = 1, OZ_ENC_FINISHED = 2, OZ_ENC_ERROR = 3 };
enum { OZ_OP_PROCESS = 0, OZ_OP_FLUSH = 1, OZ_OP_FINISH = 2 };

/* ------------------------------------------------------------------ decoder */

void *oz_decoder_new(int window_bits) {
    z_stream *s = (z_stream *)calloc(1, sizeof(z_stream));
    if (!s) return NULL;
    if (inflateInit2(s, window_bits) != Z_OK) { /* src/decoder/zlib_ng.rs:80-89: non-zero => None */
        free(s);
        return NULL;
    }
    return s;
}

oz_result oz_decode(void *state, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len) {
    z_stream *s = (z_stream *)state;
    /* `$output_len as _` / `$input_len as _` : usize -> uInt truncation is part of the reference glue
       (src/decoder/mod.rs:464,467). Kept on purpose. */
    s->avail_out = (uInt)out_len;
    s->next_out = out;
    s->avail_in = (uInt)in_len;
    s->next_in = (Bytef *)in;
    int r = inflate(s, 0 /* DEFAULT_INFLATE, src/decoder/zlib_ng.rs:15 */);
    oz_result res;
    res.input_remain = s->avail_in;
    res.output_remain = s->avail_out;
    if (r == Z_OK) res.status = s->avail_in == 0 ? OZ_NEED_INPUT : OZ_NEED_OUTPUT;
    else if (r == Z_STREAM_END) res.status = OZ_FINISHED;
    else if (r == Z_BUF_ERROR) res.status = OZ_NEED_OUTPUT;
    else res.status = r; /* Err(DecodeError(other)) */
    return res;
}

void *oz_decoder_reset(void *state) {
    return inflateReset((z_stream *)state) == Z_OK ? state : NULL;
}

void oz_decoder_free(void *state) {
    if (!state) return;
    inflateEnd((z_stream *)state);
    free(state);
}

const char *oz_describe_error(int32_t code) { return zError(code); }

/* ------------------------------------------------------------------ encoder */

void *oz_encoder_new(int level, int window_bits, int mem_level, int strategy) {
    z_stream *s = (z_stream *)calloc(1, sizeof(z_stream));
    if (!s) return NULL;
    /* strategy: Default 0, Filtered 1, HuffmanOnly 2, Rle 3, Fixed 4 (src/encoder/zlib_ng.rs:70-76) */
    if (deflateInit2(s, level, Z_DEFLATED, window_bits, mem_level, strategy) != Z_OK) {
        free(s);
        return NULL;
    }
    return s;
}

oz_result oz_encode(void *state, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len, int op) {
    z_stream *s = (z_stream *)state;
    int flush = op == OZ_OP_PROCESS ? Z_NO_FLUSH : op == OZ_OP_FLUSH ? Z_SYNC_FLUSH : Z_FINISH;
    s->avail_out = (uInt)out_len;
    s->next_out = out;
    s->avail_in = (uInt)in_len;
    s->next_in = (Bytef *)in;
    int r = deflate(s, flush);
    oz_result res;
    res.input_remain = s->avail_in;
    res.output_remain = s->avail_out;
    if (r == Z_STREAM_END) res.status = OZ_ENC_FINISHED;
    else if (r == Z_OK) res.status = flush == Z_FINISH ? OZ_ENC_NEED_OUTPUT : OZ_CONTINUE;
    else if (r == Z_BUF_ERROR) res.status = OZ_ENC_NEED_OUTPUT;
    else res.status = OZ_ENC_ERROR;
    return res;
}

void *oz_encoder_reset(void *state) {
    return deflateReset((z_stream *)state) == Z_OK ? state : NULL;
}

void oz_encoder_free(void *state) {
    if (!state) return;
    deflateEnd((z_stream *)state);
    free(state);
}

/* ------------------------------------------------------------------ checksums (for the parallel-combine checks) */

uint32_t oz_crc32(uint32_t crc, const uint8_t *p, size_t n) {
    while (n) { uInt k = n > (1u << 30) ? (1u << 30) : (uInt)n; crc = (uint32_t)crc32(crc, p, k); p += k; n -= k; }
    return crc;
}
uint32_t oz_adler32(uint32_t a, const uint8_t *p, size_t n) {
    while (n) { uInt k = n > (1u << 30) ? (1u << 30) : (uInt)n; a = (uint32_t)adler32(a, p, k); p += k; n -= k; }
    return a;
}
uint32_t oz_crc32_combine(uint32_t a, uint32_t b, uint64_t len_b) { return (uint32_t)crc32_combine(a, b, (z_off_t)len_b); }
uint32_t oz_adler32_combine(uint32_t a, uint32_t b, uint64_t len_b) { return (uint32_t)adler32_combine(a, b, (z_off_t)len_b); }

/* ------------------------------------------------------------------ batched CPU baseline (OpenMP stand-in for rayon) */

int oz_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Inflate n independent streams; stream i = in[in_off[i] .. in_off[i+1]) -> out[out_off[i] .. out_off[i+1]).
   One decoder per thread, reset between items. The loop per item is the reference caller's loop
   (README.md:33-49): call decode until Finished / error / no progress. Returns #streams that did not Finish. */
long oz_inflate_batch(size_t n, const uint8_t *in, const uint64_t *in_off, uint8_t *out, const uint64_t *out_off,
                      uint64_t *out_lens, int32_t *statuses, int window_bits, int threads) {
    long bad = 0;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads) reduction(+ : bad)
#endif
    {
        void *d = oz_decoder_new(window_bits);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 8)
#endif
        for (long i = 0; i < (long)n; i++) {
            const uint8_t *ip = in + in_off[i];
            size_t il = (size_t)(in_off[i + 1] - in_off[i]);
            uint8_t *op = out + out_off[i];
            size_t ol = (size_t)(out_off[i + 1] - out_off[i]);
            size_t produced = 0;
            int32_t st = OZ_NEED_INPUT;
            for (;;) {
                /* feed < 4 GiB per call: the glue truncates to uInt */
                size_t ci = il > 0x40000000u ? 0x40000000u : il, co = ol > 0x40000000u ? 0x40000000u : ol;
                oz_result r = oz_decode(d, ip, ci, op, co);
                size_t used = ci - r.input_remain, made = co - r.output_remain;
                ip += used; il -= used; op += made; ol -= made; produced += made;
                st = r.status;
                if (st == OZ_FINISHED || st < 0) break;
                if (used == 0 && made == 0) break;           /* no progress: truncated input or full output */
                if (st == OZ_NEED_INPUT && il == 0) break;   /* truncated */
                if (st == OZ_NEED_OUTPUT && ol == 0) break;  /* output capacity exhausted */
            }
            out_lens[i] = produced;
            statuses[i] = st;
            if (st != OZ_FINISHED) bad++;
            d = oz_decoder_reset(d);
        }
        oz_decoder_free(d);
    }
    return bad;
}

/* Deflate n independent buffers, each into its own stream with Finish. */
long oz_deflate_batch(size_t n, const uint8_t *in, const uint64_t *in_off, uint8_t *out, const uint64_t *out_off,
                      uint64_t *out_lens, int32_t *statuses, int level, int window_bits, int mem_level, int strategy,
                      int threads) {
    long bad = 0;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads) reduction(+ : bad)
#endif
    {
        void *e = oz_encoder_new(level, window_bits, mem_level, strategy);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (long i = 0; i < (long)n; i++) {
            const uint8_t *ip = in + in_off[i];
            size_t il = (size_t)(in_off[i + 1] - in_off[i]);
            uint8_t *op = out + out_off[i];
            size_t ol = (size_t)(out_off[i + 1] - out_off[i]);
            size_t produced = 0;
            int32_t st = OZ_CONTINUE;
            for (;;) {
                size_t ci = il > 0x40000000u ? 0x40000000u : il, co = ol > 0x40000000u ? 0x40000000u : ol;
                int fin = ci == il;
                oz_result r = oz_encode(e, ip, ci, op, co, fin ? OZ_OP_FINISH : OZ_OP_PROCESS);
                size_t used = ci - r.input_remain, made = co - r.output_remain;
                ip += used; il -= used; op += made; ol -= made; produced += made;
                st = r.status;
                if (st == OZ_ENC_FINISHED || st == OZ_ENC_ERROR) break;
                if (used == 0 && made == 0) break;
                if (st == OZ_ENC_NEED_OUTPUT && ol == 0) break;
            }
            out_lens[i] = produced;
            statuses[i] = st;
            if (st != OZ_ENC_FINISHED) bad++;
            e = oz_encoder_reset(e);
        }
        oz_encoder_free(e);
    }
    return bad;
}

/* Worst-case compressed size for the harness' output slots (deflateBound of a fresh stream). */
uint64_t oz_deflate_bound(uint64_t len, int level, int window_bits) {
    z_stream s;
    memset(&s, 0, sizeof s);
    if (deflateInit2(&s, level, Z_DEFLATED, window_bits, 8, 0) != Z_OK) return len + (len >> 3) + 64;
    uint64_t b = 0, rem = len;
    while (rem > 0x40000000u) { b += deflateBound(&s, 0x40000000u); rem -= 0x40000000u; }
    b += deflateBound(&s, (uLong)rem);
    deflateEnd(&s);
    return b;
}

const char *oz_zlib_version(void) { return zlibVersion(); }
