// Local windowed 3D attention, exact SIMT kernels (fp32 math, fp32 or bf16 I/O).
//
// These are the WM_DTYPE_FP32 path (parity bar 1e-5 against the reference's fp32
// modules) and the device-side cross-check of the tcgen05 kernels (WM_FLAG_SIMT).
// One warp owns one (token, head) pair:
//   scores : lanes run over the window's keys, each lane does whole d-long dot products
//   softmax: warp shuffles
//   PV     : lanes run over channels, keys are walked sequentially (deterministic)
// Backward is split query-stationary (dQ) / key-stationary (dK, dV); the key-stationary
// pass uses the symmetry of the window (the queries that see key j are the keys that
// query j sees), so there are no atomics and results are bitwise reproducible.
//
// Replaces: Local3dAttention.local_attention, vq-video-diffusion/local_3d_attention.py:78-99.
#include "wm_common.cuh"

#include <math.h>

namespace wm {
namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kMaxChanPerLane = 8;   // d <= 256

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void from_f(float* p, float v) { *p = v; }
__device__ __forceinline__ void from_f(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// dot product of a shared-memory fp32 vector with a global row (16-byte vector loads)
__device__ __forceinline__ float dot_row(const float* __restrict__ a, const float* __restrict__ row, int d) {
    float acc = 0.f;
    for (int c = 0; c < d; c += 4) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(row + c));
        acc = fmaf(a[c], r.x, acc);
        acc = fmaf(a[c + 1], r.y, acc);
        acc = fmaf(a[c + 2], r.z, acc);
        acc = fmaf(a[c + 3], r.w, acc);
    }
    return acc;
}
__device__ __forceinline__ float dot_row(const float* __restrict__ a, const __nv_bfloat16* __restrict__ row, int d) {
    float acc = 0.f;
    for (int c = 0; c < d; c += 8) {
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(row + c));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(h[i]);
            acc = fmaf(a[c + 2 * i], f.x, acc);
            acc = fmaf(a[c + 2 * i + 1], f.y, acc);
        }
    }
    return acc;
}

struct Item {
    long tok;
    int b, s, h, w, head;
};

__device__ __forceinline__ bool decode_item(const AttnShape& sh, long item, Item& it) {
    const long total = sh.tokens() * sh.heads;
    if (item >= total) return false;
    it.head = (int)(item % sh.heads);
    long t = item / sh.heads;
    it.tok = t;
    it.w = (int)(t % sh.W); t /= sh.W;
    it.h = (int)(t % sh.H); t /= sh.H;
    it.s = (int)(t % sh.S);
    it.b = (int)(t / sh.S);
    return true;
}

// window slot j -> neighbour token (or -1).  Row-major (i j k) order like the reference.
__device__ __forceinline__ long neighbour(const AttnShape& sh, const Item& it, int j) {
    const int wW = sh.wW(), wH = sh.wH();
    const int dk = j % wW;
    const int r = j / wW;
    const int dj = r % wH;
    const int di = r / wH;
    const int ks = it.s + di - sh.eS, kh = it.h + dj - sh.eH, kw = it.w + dk - sh.eW;
    if (ks < 0 || ks >= sh.S || kh < 0 || kh >= sh.H || kw < 0 || kw >= sh.W) return -1;
    return (((long)it.b * sh.S + ks) * sh.H + kh) * sh.W + kw;
}

// ------------------------------------------------------------------------------ forward
// One (token, head) by one warp; qs / ps: this warp's d + Wn floats of shared memory.
template <typename T>
__device__ __forceinline__ void fwd_item(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                                         T* __restrict__ o, float* __restrict__ lse, const AttnShape& sh, const Item& it,
                                         float* qs, float* ps, int lane) {
    const int Wn = sh.window(), d = sh.d;
    const long inner = sh.inner(), ldq = sh.q_ld(), ldkv = sh.kv_ld();
    const long hoff = (long)it.head * d;
    for (int c = lane; c < d; c += 32) qs[c] = to_f(q[it.tok * ldq + hoff + c]);
    __syncwarp();

    float mx = -INFINITY;
    for (int j = lane; j < Wn; j += 32) {
        const long kt = neighbour(sh, it, j);
        float sc = -INFINITY;
        if (kt >= 0) sc = dot_row(qs, k + kt * ldkv + hoff, d) * sh.scale;
        ps[j] = sc;
        mx = fmaxf(mx, sc);
    }
    mx = warp_max(mx);       // the centre key is always inside the grid, so mx is finite
    float sum = 0.f;
    for (int j = lane; j < Wn; j += 32) {
        const float sc = ps[j];
        const float p = (sc == -INFINITY) ? 0.f : expf(sc - mx);
        ps[j] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.f / sum;

    float acc[kMaxChanPerLane];
#pragma unroll
    for (int i = 0; i < kMaxChanPerLane; ++i) acc[i] = 0.f;
    for (int j = 0; j < Wn; ++j) {
        const float p = ps[j];
        if (p == 0.f) continue;                          // warp-uniform
        const long kt = neighbour(sh, it, j);
        const T* vrow = v + kt * ldkv + hoff;
#pragma unroll
        for (int i = 0; i < kMaxChanPerLane; ++i) {
            const int c = lane + 32 * i;
            if (c < d) acc[i] = fmaf(p, to_f(vrow[c]), acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < kMaxChanPerLane; ++i) {
        const int c = lane + 32 * i;
        if (c < d) from_f(o + it.tok * inner + hoff + c, acc[i] * inv);
    }
    if (lane == 0) lse[it.tok * sh.heads + it.head] = mx + logf(sum);
    __syncwarp();
}

template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
l3d_fwd_simt_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                    T* __restrict__ o, float* __restrict__ lse, const AttnShape sh) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* qs = smem + warp * (sh.d + sh.window());
    Item it;
    if (!decode_item(sh, (long)blockIdx.x * kWarpsPerBlock + warp, it)) return;
    fwd_item(q, k, v, o, lse, sh, it, qs, qs + sh.d, lane);
}

// Fix-up pass behind the tensor-core forward: every lane looks at one (token, head); the warp then recomputes, one
// after the other, those whose LSE the tensor-core kernel marked NaN (row sum outside the range of its max-free
// softmax).  Normally nothing is marked and the kernel is one coalesced read of the LSE tensor.
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
l3d_fwd_fixup_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                     const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ o, float* __restrict__ lse,
                     const AttnShape sh) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* qs = smem + warp * (sh.d + sh.window());
    const long total = sh.tokens() * sh.heads;
    const long base = ((long)blockIdx.x * kWarpsPerBlock + warp) * 32;
    const long mine = base + lane;
    const float l = mine < total ? lse[mine] : 0.f;
    unsigned todo = __ballot_sync(0xffffffffu, l != l);
    while (todo) {
        const int j = __ffs(todo) - 1;
        todo &= todo - 1;
        Item it;
        decode_item(sh, base + j, it);
        fwd_item(q, k, v, o, lse, sh, it, qs, qs + sh.d, lane);
    }
}

// --------------------------------------------------------------------------- backward dQ
// Also produces delta = rowsum(dO * O), consumed by the dK/dV pass.
template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
l3d_bwd_dq_simt_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                       const T* __restrict__ o, const float* __restrict__ lse, const T* __restrict__ dout,
                       T* __restrict__ dq, float* __restrict__ delta, const AttnShape sh) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Wn = sh.window(), d = sh.d;
    float* qs = smem + warp * (2 * d + Wn);
    float* dos = qs + d;
    float* dss = dos + d;
    Item it;
    if (!decode_item(sh, (long)blockIdx.x * kWarpsPerBlock + warp, it)) return;
    const long inner = sh.inner(), ldq = sh.q_ld(), ldkv = sh.kv_ld();
    const long hoff = (long)it.head * d;
    const long row = it.tok * inner + hoff, qrow_off = it.tok * ldq + hoff;
    float dl = 0.f;
    for (int c = lane; c < d; c += 32) {
        qs[c] = to_f(q[qrow_off + c]);
        const float g = to_f(dout[row + c]);
        dos[c] = g;
        dl = fmaf(g, to_f(o[row + c]), dl);
    }
    dl = warp_sum(dl);
    const float L = lse[it.tok * sh.heads + it.head];
    if (lane == 0) delta[it.tok * sh.heads + it.head] = dl;
    __syncwarp();

    for (int j = lane; j < Wn; j += 32) {
        const long kt = neighbour(sh, it, j);
        float ds = 0.f;
        if (kt >= 0) {
            const float sc = dot_row(qs, k + kt * ldkv + hoff, d) * sh.scale;
            const float p = expf(sc - L);
            const float dp = dot_row(dos, v + kt * ldkv + hoff, d);
            ds = p * (dp - dl) * sh.scale;
        }
        dss[j] = ds;
    }
    __syncwarp();

    float acc[kMaxChanPerLane];
#pragma unroll
    for (int i = 0; i < kMaxChanPerLane; ++i) acc[i] = 0.f;
    for (int j = 0; j < Wn; ++j) {
        const long kt = neighbour(sh, it, j);
        if (kt < 0) continue;                            // warp-uniform
        const float ds = dss[j];
        const T* krow = k + kt * ldkv + hoff;
#pragma unroll
        for (int i = 0; i < kMaxChanPerLane; ++i) {
            const int c = lane + 32 * i;
            if (c < d) acc[i] = fmaf(ds, to_f(krow[c]), acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < kMaxChanPerLane; ++i) {
        const int c = lane + 32 * i;
        if (c < d) from_f(dq + qrow_off + c, acc[i]);
    }
}

// ------------------------------------------------------------------------ backward dK, dV
template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
l3d_bwd_dkv_simt_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                        const float* __restrict__ lse, const float* __restrict__ delta,
                        const T* __restrict__ dout, T* __restrict__ dk, T* __restrict__ dv,
                        const AttnShape sh) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Wn = sh.window(), d = sh.d;
    float* ks = smem + warp * (2 * d + 2 * Wn);
    float* vs = ks + d;
    float* ps = vs + d;
    float* dss = ps + Wn;
    Item it;                                            // here the item is a KEY token
    if (!decode_item(sh, (long)blockIdx.x * kWarpsPerBlock + warp, it)) return;
    const long inner = sh.inner(), ldq = sh.q_ld(), ldkv = sh.kv_ld();
    const long hoff = (long)it.head * d;
    const long row = it.tok * ldkv + hoff;
    for (int c = lane; c < d; c += 32) {
        ks[c] = to_f(k[row + c]);
        vs[c] = to_f(v[row + c]);
    }
    __syncwarp();

    // window symmetry: the queries that attend to this key are its own window neighbours
    for (int j = lane; j < Wn; j += 32) {
        const long qt = neighbour(sh, it, j);
        float p = 0.f, ds = 0.f;
        if (qt >= 0) {
            const float sc = dot_row(ks, q + qt * ldq + hoff, d) * sh.scale;
            p = expf(sc - lse[qt * sh.heads + it.head]);
            const float dp = dot_row(vs, dout + qt * inner + hoff, d);
            ds = p * (dp - delta[qt * sh.heads + it.head]) * sh.scale;
        }
        ps[j] = p;
        dss[j] = ds;
    }
    __syncwarp();

    float accv[kMaxChanPerLane], acck[kMaxChanPerLane];
#pragma unroll
    for (int i = 0; i < kMaxChanPerLane; ++i) accv[i] = acck[i] = 0.f;
    for (int j = 0; j < Wn; ++j) {
        const long qt = neighbour(sh, it, j);
        if (qt < 0) continue;                            // warp-uniform
        const float p = ps[j], ds = dss[j];
        const T* dorow = dout + qt * inner + hoff;
        const T* qrow = q + qt * ldq + hoff;
#pragma unroll
        for (int i = 0; i < kMaxChanPerLane; ++i) {
            const int c = lane + 32 * i;
            if (c < d) {
                accv[i] = fmaf(p, to_f(dorow[c]), accv[i]);
                acck[i] = fmaf(ds, to_f(qrow[c]), acck[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kMaxChanPerLane; ++i) {
        const int c = lane + 32 * i;
        if (c < d) {
            from_f(dv + row + c, accv[i]);
            from_f(dk + row + c, acck[i]);
        }
    }
}

int check_simt_shape(const AttnShape& s, int dtype) {
    const int vec = (dtype == WM_DTYPE_FP32) ? 4 : 8;
    if (s.d % vec != 0 || s.d > 32 * kMaxChanPerLane)
        return fail(WM_EUNSUPPORTED, "dim_head=%d: SIMT kernels need a multiple of %d, at most %d", s.d, vec,
                    32 * kMaxChanPerLane);
    if (s.window() > 4096) return fail(WM_EUNSUPPORTED, "window of %d keys is too large", s.window());
    return WM_OK;
}

template <typename T>
int launch_fwd(const void* q, const void* k, const void* v, void* o, float* lse, const AttnShape& s, cudaStream_t st) {
    const long items = s.tokens() * s.heads;
    const unsigned grid = (unsigned)((items + kWarpsPerBlock - 1) / kWarpsPerBlock);
    const size_t smem = (size_t)kWarpsPerBlock * (s.d + s.window()) * sizeof(float);
    WM_CUDA_CHECK(cudaFuncSetAttribute(l3d_fwd_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    l3d_fwd_simt_kernel<T><<<grid, kWarpsPerBlock * 32, smem, st>>>(
        static_cast<const T*>(q), static_cast<const T*>(k), static_cast<const T*>(v), static_cast<T*>(o), lse, s);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}

template <typename T>
int launch_bwd(const void* q, const void* k, const void* v, const void* o, const float* lse, const void* dout,
               void* dq, void* dk, void* dv, float* delta, const AttnShape& s, cudaStream_t st) {
    const long items = s.tokens() * s.heads;
    const unsigned grid = (unsigned)((items + kWarpsPerBlock - 1) / kWarpsPerBlock);
    const size_t smem_q = (size_t)kWarpsPerBlock * (2 * s.d + s.window()) * sizeof(float);
    const size_t smem_kv = (size_t)kWarpsPerBlock * (2 * s.d + 2 * s.window()) * sizeof(float);
    WM_CUDA_CHECK(cudaFuncSetAttribute(l3d_bwd_dq_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q));
    WM_CUDA_CHECK(cudaFuncSetAttribute(l3d_bwd_dkv_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv));
    l3d_bwd_dq_simt_kernel<T><<<grid, kWarpsPerBlock * 32, smem_q, st>>>(
        static_cast<const T*>(q), static_cast<const T*>(k), static_cast<const T*>(v), static_cast<const T*>(o), lse,
        static_cast<const T*>(dout), static_cast<T*>(dq), delta, s);
    WM_CUDA_CHECK(cudaGetLastError());
    l3d_bwd_dkv_simt_kernel<T><<<grid, kWarpsPerBlock * 32, smem_kv, st>>>(
        static_cast<const T*>(q), static_cast<const T*>(k), static_cast<const T*>(v), lse, delta,
        static_cast<const T*>(dout), static_cast<T*>(dk), static_cast<T*>(dv), s);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}

}  // namespace

int attn_fwd_simt(const void* q, const void* k, const void* v, void* o, float* lse, const AttnShape& s, int dtype,
                  cudaStream_t st) {
    if (int rc = check_simt_shape(s, dtype)) return rc;
    return dtype == WM_DTYPE_FP32 ? launch_fwd<float>(q, k, v, o, lse, s, st)
                                  : launch_fwd<__nv_bfloat16>(q, k, v, o, lse, s, st);
}

int attn_fwd_fixup(const void* q, const void* k, const void* v, void* o, float* lse, const AttnShape& s, cudaStream_t st) {
    if (int rc = check_simt_shape(s, WM_DTYPE_BF16)) return rc;
    const long items = s.tokens() * s.heads;
    const long blocks = (items + kWarpsPerBlock * 32 - 1) / (kWarpsPerBlock * 32);
    const size_t smem = (size_t)kWarpsPerBlock * (s.d + s.window()) * sizeof(float);
    WM_CUDA_CHECK(cudaFuncSetAttribute(l3d_fwd_fixup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    l3d_fwd_fixup_kernel<<<(unsigned)blocks, kWarpsPerBlock * 32, smem, st>>>(
        static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k), static_cast<const __nv_bfloat16*>(v),
        static_cast<__nv_bfloat16*>(o), lse, s);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}

int attn_bwd_simt(const void* q, const void* k, const void* v, const void* o, const float* lse, const void* dout,
                  void* dq, void* dk, void* dv, float* delta, const AttnShape& s, int dtype, cudaStream_t st) {
    if (int rc = check_simt_shape(s, dtype)) return rc;
    return dtype == WM_DTYPE_FP32 ? launch_bwd<float>(q, k, v, o, lse, dout, dq, dk, dv, delta, s, st)
                                  : launch_bwd<__nv_bfloat16>(q, k, v, o, lse, dout, dq, dk, dv, delta, s, st);
}

}  // namespace wm
