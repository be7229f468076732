// Shared host/device helpers for libwm_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/wm_b200.h"

namespace wm {

// ---- error plumbing: nothing throws across the C ABI ---------------------------------
std::string& last_error();                       // thread-local
int fail(int code, const char* fmt, ...);

#define WM_CUDA_CHECK(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess)                                                           \
            return ::wm::fail(WM_ECUDA, "%s failed: %s (%s:%d)", #expr,                  \
                              cudaGetErrorString(_e), __FILE__, __LINE__);               \
    } while (0)

// ---- problem description shared by every attention kernel ----------------------------
struct AttnShape {
    int B, S, H, W, heads, d;       // d = dim_head
    int eS, eH, eW;                 // window extents; window = (2e+1) per axis
    float scale;
    int ldq = 0, ldkv = 0;          // elements between consecutive tokens of q (and dq) / of k, v, dk, dv; 0 = heads*d.
                                    // A merged projection writes q | k | v side by side: its slices are addressed in place.
    __host__ __device__ int inner() const { return heads * d; }
    __host__ __device__ long q_ld() const { return ldq ? ldq : heads * d; }
    __host__ __device__ long kv_ld() const { return ldkv ? ldkv : heads * d; }
    __host__ __device__ long tokens() const { return (long)B * S * H * W; }
    __host__ __device__ int wS() const { return 2 * eS + 1; }
    __host__ __device__ int wH() const { return 2 * eH + 1; }
    __host__ __device__ int wW() const { return 2 * eW + 1; }
    __host__ __device__ int window() const { return wS() * wH() * wW(); }
};

int sm_count();                                  // multiprocessors of the current device (cached per device)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- kernel families (defined in attn_simt.cu / attn_tc.cu) ----------------------------
int attn_fwd_simt(const void* q, const void* k, const void* v, void* o, float* lse,
                  const AttnShape& s, int dtype, cudaStream_t st);
int attn_bwd_simt(const void* q, const void* k, const void* v, const void* o, const float* lse,
                  const void* dout, void* dq, void* dk, void* dv, float* delta,
                  const AttnShape& s, int dtype, cudaStream_t st);

// bf16: recompute (exactly, like attn_fwd_simt) every (token, head) whose lse is NaN -- the tensor-core forward's
// marker for rows outside the range of its max-free softmax
int attn_fwd_fixup(const void* q, const void* k, const void* v, void* o, float* lse, const AttnShape& s, cudaStream_t st);

bool attn_tc_supported(const AttnShape& s);
int attn_fwd_tc(const void* q, const void* k, const void* v, void* o, float* lse,
                const AttnShape& s, cudaStream_t st);
int attn_bwd_tc(const void* q, const void* k, const void* v, const void* o, const float* lse,
                const void* dout, void* dq, void* dk, void* dv, float* delta,
                const AttnShape& s, cudaStream_t st);

bool vq_tc_supported(long N, int L, int K, int D);
int vq_nearest_tc(const void* x, const void* cb, int64_t* idx, void* quantized, float* sq_err, long N, int L, int K,
                  int D, cudaStream_t st);

}  // namespace wm
