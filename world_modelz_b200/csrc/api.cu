// C-ABI entry points of libwm_b200 (declared in include/wm_b200.h): argument checks and
// dispatch to the kernel families.  No torch headers, no allocation, no synchronisation.
#include "wm_common.cuh"

#include <stdarg.h>

namespace wm {

std::string& last_error() {
    static thread_local std::string msg;
    return msg;
}

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        cached[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
    }
    return cached[dev];
}

namespace {

int check_attn(const AttnShape& s, int dtype, const void* const* ptrs, int nptrs) {
    if (dtype != WM_DTYPE_BF16 && dtype != WM_DTYPE_FP32) return fail(WM_EINVAL, "unknown dtype %d", dtype);
    if (s.B <= 0 || s.S <= 0 || s.H <= 0 || s.W <= 0 || s.heads <= 0 || s.d <= 0)
        return fail(WM_EINVAL, "bad shape B=%d S=%d H=%d W=%d heads=%d dim_head=%d", s.B, s.S, s.H, s.W, s.heads, s.d);
    if (s.eS < 0 || s.eH < 0 || s.eW < 0) return fail(WM_EINVAL, "negative extents (%d,%d,%d)", s.eS, s.eH, s.eW);
    if ((double)s.tokens() * s.inner() >= 9.0e18) return fail(WM_EINVAL, "tensor too large");
    for (int i = 0; i < nptrs; ++i) {
        if (ptrs[i] == nullptr) return fail(WM_EINVAL, "null pointer (argument %d)", i);
        if (!aligned16(ptrs[i])) return fail(WM_EINVAL, "pointer argument %d is not 16-byte aligned", i);
    }
    return WM_OK;
}

}  // namespace
}  // namespace wm

using namespace wm;

extern "C" int wm_version(void) { return WM_B200_VERSION; }

extern "C" const char* wm_last_error(void) { return last_error().c_str(); }

extern "C" int wm_l3d_attn_uses_tensor_cores(int S, int H, int W, int heads, int dim_head, int eS, int eH, int eW,
                                             int dtype) {
    if (dtype != WM_DTYPE_BF16) return 0;
    AttnShape s{1, S, H, W, heads, dim_head, eS, eH, eW, 1.f};
    return attn_tc_supported(s) ? 1 : 0;
}

namespace {
int check_ld(AttnShape& s, long ld_q, long ld_kv, int dtype) {
    const long esz = dtype == WM_DTYPE_BF16 ? 2 : 4;
    if ((ld_q != 0 && ld_q < s.inner()) || (ld_kv != 0 && ld_kv < s.inner()) || ld_q > 0x7fffffffL || ld_kv > 0x7fffffffL)
        return fail(WM_EINVAL, "token strides ld_q=%ld, ld_kv=%ld must be 0 or >= heads*dim_head=%d", ld_q, ld_kv, s.inner());
    if ((ld_q * esz) % 16 != 0 || (ld_kv * esz) % 16 != 0)
        return fail(WM_EINVAL, "token strides must be multiples of 16 bytes (ld_q=%ld, ld_kv=%ld elements)", ld_q, ld_kv);
    s.ldq = (int)ld_q;
    s.ldkv = (int)ld_kv;
    return WM_OK;
}
}  // namespace

extern "C" int wm_l3d_attn_fwd_ld(const void* q, const void* k, const void* v, void* o, float* lse, long ld_q, long ld_kv,
                                  int B, int S, int H, int W, int heads, int dim_head, int eS, int eH, int eW,
                                  float scale, int dtype, int flags, void* stream) {
    AttnShape s{B, S, H, W, heads, dim_head, eS, eH, eW, scale};
    const void* ptrs[] = {q, k, v, o, lse};
    if (int rc = check_attn(s, dtype, ptrs, 5)) return rc;
    if (int rc = check_ld(s, ld_q, ld_kv, dtype)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == WM_DTYPE_BF16 && !(flags & WM_FLAG_SIMT) && attn_tc_supported(s))
        return attn_fwd_tc(q, k, v, o, lse, s, st);
    return attn_fwd_simt(q, k, v, o, lse, s, dtype, st);
}

extern "C" int wm_l3d_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int S, int H,
                               int W, int heads, int dim_head, int eS, int eH, int eW, float scale, int dtype,
                               int flags, void* stream) {
    return wm_l3d_attn_fwd_ld(q, k, v, o, lse, 0, 0, B, S, H, W, heads, dim_head, eS, eH, eW, scale, dtype, flags, stream);
}

extern "C" int wm_l3d_attn_bwd_ld(const void* q, const void* k, const void* v, const void* o, const float* lse,
                                  const void* dout, void* dq, void* dk, void* dv, float* delta, long ld_q, long ld_kv,
                                  int B, int S, int H, int W, int heads, int dim_head, int eS, int eH, int eW,
                                  float scale, int dtype, int flags, void* stream) {
    AttnShape s{B, S, H, W, heads, dim_head, eS, eH, eW, scale};
    const void* ptrs[] = {q, k, v, o, lse, dout, dq, dk, dv, delta};
    if (int rc = check_attn(s, dtype, ptrs, 10)) return rc;
    if (int rc = check_ld(s, ld_q, ld_kv, dtype)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == WM_DTYPE_BF16 && !(flags & WM_FLAG_SIMT) && attn_tc_supported(s))
        return attn_bwd_tc(q, k, v, o, lse, dout, dq, dk, dv, delta, s, st);
    return attn_bwd_simt(q, k, v, o, lse, dout, dq, dk, dv, delta, s, dtype, st);
}

extern "C" int wm_l3d_attn_bwd(const void* q, const void* k, const void* v, const void* o, const float* lse,
                               const void* dout, void* dq, void* dk, void* dv, float* delta, int B, int S, int H,
                               int W, int heads, int dim_head, int eS, int eH, int eW, float scale, int dtype,
                               int flags, void* stream) {
    return wm_l3d_attn_bwd_ld(q, k, v, o, lse, dout, dq, dk, dv, delta, 0, 0, B, S, H, W, heads, dim_head, eS, eH, eW,
                              scale, dtype, flags, stream);
}
