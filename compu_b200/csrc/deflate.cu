// deflate.cu — encoder side of the C ABI. PLACEHOLDER until the deflate kernel lands: every entry point fails loudly.
#include "host_common.h"
using namespace czh;

extern "C" void *cz_encoder_new(int, int, int, int) { set_error("encoder not built yet"); return nullptr; }
extern "C" cz_result cz_encode(void *, const uint8_t *, size_t in_len, uint8_t *, size_t out_len, int) {
    cz_result r; r.input_remain = in_len; r.output_remain = out_len; r.status = CZ_ENCODE_ERROR; return r;
}
extern "C" void *cz_encoder_reset(void *) { return nullptr; }
extern "C" void cz_encoder_free(void *) {}
extern "C" uint64_t cz_deflate_bound(uint64_t len, int, uint64_t) { return len + len / 8 + 1024; }
extern "C" int cz_deflate_batch(size_t, const uint8_t *, const uint64_t *, uint8_t *, const uint64_t *, uint64_t *, int32_t *, int, int,
                                int, uint64_t, uint32_t) { set_error("deflate not built yet"); return CZ_E_STREAM; }
extern "C" int cz_deflate_segmented(const uint8_t *, uint64_t, uint8_t *, uint64_t, uint64_t *, int, int, int, uint64_t, uint32_t,
                                    uint64_t *, uint64_t, uint64_t *) { set_error("deflate not built yet"); return CZ_E_STREAM; }
extern "C" int cz_inflate_segmented(const uint8_t *, uint64_t, uint8_t *, uint64_t, uint64_t *, int, uint64_t, const uint64_t *,
                                    uint64_t, uint32_t) { set_error("not built yet"); return CZ_E_STREAM; }
extern "C" uint64_t cz_deflate_max_segment(void) { return 1u << 20; }
extern "C" uint64_t cz_deflate_segment_bound(uint64_t n) { return n + n / 8 + 1024; }
extern "C" uint64_t cz_deflate_workspace_bytes(size_t) { return 256; }
extern "C" int cz_deflate_segments_device(void *, size_t, const uint8_t *, const uint64_t *, uint8_t *, const uint64_t *, uint64_t *,
                                          int32_t *, uint32_t *, int, int, void *, uint64_t) { set_error("deflate not built yet"); return CZ_E_STREAM; }
