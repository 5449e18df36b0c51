// TEMPORARY: entry points not implemented yet (removed as the real ones land).
#include "rpx_common.cuh"
#define STUB(name) rpx::set_error(#name ": not implemented yet"); return RPX_ERR_UNSUPPORTED
extern "C" {
size_t rpx_encoder_packed_bytes(const rpx_t5_config*) { return 0; }
int rpx_encoder_create(const rpx_t5_config*, const rpx_t5_weights*, void*, size_t, void*, rpx_encoder**) { STUB(rpx_encoder_create); }
int rpx_encoder_destroy(rpx_encoder*) { STUB(rpx_encoder_destroy); }
size_t rpx_encoder_workspace_bytes(const rpx_encoder*, int64_t, int64_t) { return 0; }
int rpx_encode_bytes(rpx_encoder*, const uint8_t*, const int64_t*, int32_t, int32_t, void*, int32_t, void*, size_t, void*) { STUB(rpx_encode_bytes); }
int rpx_encode_ids(rpx_encoder*, const int64_t*, const int64_t*, int32_t, int32_t, void*, int32_t, void*, size_t, void*) { STUB(rpx_encode_ids); }
int rpx_encoder_set_debug_hidden(rpx_encoder*, float*) { STUB(rpx_encoder_set_debug_hidden); }
int rpx_encoder_set_profiling(rpx_encoder*, int32_t) { STUB(rpx_encoder_set_profiling); }
int rpx_encoder_read_profile(rpx_encoder*, float*, int64_t*) { STUB(rpx_encoder_read_profile); }
size_t rpx_sim_topk_workspace_bytes(int32_t, int32_t) { return 0; }
int rpx_sim_topk(const void*, int32_t, const void*, int64_t, int32_t, int32_t, const uint32_t*, int64_t, float*, double*, int64_t*, int32_t*, int64_t, void*, size_t, void*) { STUB(rpx_sim_topk); }
int rpx_topk_merge(const double*, const int64_t*, int32_t, int32_t, int32_t, float*, double*, int64_t*, int32_t*, void*) { STUB(rpx_topk_merge); }
}
