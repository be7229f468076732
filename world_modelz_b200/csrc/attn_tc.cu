// placeholder until the tcgen05 kernels land
#include "wm_common.cuh"
namespace wm {
bool attn_tc_supported(const AttnShape&) { return false; }
int attn_fwd_tc(const void*, const void*, const void*, void*, float*, const AttnShape&, cudaStream_t) {
    return fail(WM_EUNSUPPORTED, "tcgen05 attention forward not built");
}
int attn_bwd_tc(const void*, const void*, const void*, const void*, const float*, const void*, void*, void*, void*,
                float*, const AttnShape&, cudaStream_t) {
    return fail(WM_EUNSUPPORTED, "tcgen05 attention backward not built");
}
}  // namespace wm
