// TEMPORARY: entry points not implemented yet (removed as the real ones land).
#include "rpx_common.cuh"
#define STUB(name) rpx::set_error(#name ": not implemented yet"); return RPX_ERR_UNSUPPORTED
extern "C" {
size_t rpx_sim_topk_workspace_bytes(int32_t, int32_t) { return 0; }
int rpx_sim_topk(const void*, int32_t, const void*, int64_t, int32_t, int32_t, const uint32_t*, int64_t, float*, double*, int64_t*, int32_t*, int64_t, void*, size_t, void*) { STUB(rpx_sim_topk); }
int rpx_topk_merge(const double*, const int64_t*, int32_t, int32_t, int32_t, float*, double*, int64_t*, int32_t*, void*) { STUB(rpx_topk_merge); }
}
