// Layer-level kernels around the attention core: fused residual-add + LayerNorm (forward and
// backward) and column sums (bias gradients).  One warp per token row, 16-byte vector
// accesses, fp32 statistics.
//
// Reference: PreNorm / the residual adds of Local3dAttentionTransformer.forward
// (vq-video-diffusion/local_3d_attention.py:11-17,159-161) and the bias gradients that autograd
// derives for its nn.Linear layers (:24-27,46-53).
#include "wm_common.cuh"

#include <math.h>

namespace wm {
namespace {

constexpr int kMaxChunks = 8;          // 8 elements per lane per chunk of 256 -> dim <= 2048

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 8 consecutive elements <-> 8 floats
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
}
__device__ __forceinline__ void store8(float* p, const float (&f)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
// 8 consecutive elements as loaded (no conversion): lets a prefetch sit in few registers
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> {
    uint4 v;
    __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void unpack(float (&f)[8]) const {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 t = __bfloat1622float2(h[i]);
            f[2 * i] = t.x;
            f[2 * i + 1] = t.y;
        }
    }
};
template <> struct Raw8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* p) {
        a = *reinterpret_cast<const float4*>(p);
        b = *reinterpret_cast<const float4*>(p + 4);
    }
    __device__ __forceinline__ void unpack(float (&f)[8]) const {
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
};
__device__ __forceinline__ float round_to(float v, const __nv_bfloat16*) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float round_to(float v, const float*) { return v; }

// sum = res (+ delta); y = LayerNorm(sum) * gamma + beta; mean / rstd kept for backward
template <typename T, int CHUNKS>
__global__ void __launch_bounds__(256)
add_layernorm_fwd_kernel(const T* __restrict__ res, const T* __restrict__ delta, const T* __restrict__ delta_bias,
                         const T* __restrict__ gamma, const T* __restrict__ beta, T* __restrict__ sum_out, T* __restrict__ y,
                         float* __restrict__ mean_out, float* __restrict__ rstd_out, long rows, int dim, float eps) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    float x[CHUNKS][8];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int col = c * 256 + lane * 8;
        if (col < dim) {
            load8(res + row * dim + col, x[c]);
            if (delta != nullptr) {
                float d[8];
                load8(delta + row * dim + col, d);
                if (delta_bias != nullptr) {             // bias of the linear layer that produced delta, applied here
                    float bb[8];
                    load8(delta_bias + col, bb);
#pragma unroll
                    for (int i = 0; i < 8; ++i) d[i] += bb[i];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) x[c][i] = round_to(x[c][i] + d[i], res);   // the sum as the next layer reads it
                if (sum_out != nullptr) store8(sum_out + row * dim + col, x[c]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) s += x[c][i];
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[c][i] = 0.f;
        }
    }
    const float mean = warp_sum(s) / dim;
    float v = 0.f;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c)
        if (c * 256 + lane * 8 < dim)
#pragma unroll
            for (int i = 0; i < 8; ++i) v = fmaf(x[c][i] - mean, x[c][i] - mean, v);
    const float rstd = rsqrtf(warp_sum(v) / dim + eps);
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int col = c * 256 + lane * 8;
        if (col < dim) {
            float g[8], b[8], o[8];
            load8(gamma + col, g);
            load8(beta + col, b);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = fmaf((x[c][i] - mean) * rstd, g[i], b[i]);
            store8(y + row * dim + col, o);
        }
    }
    if (lane == 0) {
        mean_out[row] = mean;
        rstd_out[row] = rstd;
    }
}

// dx = (dres +) LayerNorm backward(dy); per-block partial sums of dgamma / dbeta
// rows per warp iteration / resident CTAs per SM at dim <= 256, measured at 131 072 x 256 bf16 (same box, L2 flushed):
// (2, 2) 77.9 us with register spills, (1, 2) 74.7 us without, (1, 3) 97.7, (1, 4) 103.6, (2, 1) 82.9
#ifndef WM_LN_MINB
#define WM_LN_MINB 2
#endif
#ifndef WM_LN_R
#define WM_LN_R 1
#endif
template <typename T, int CHUNKS, bool DSUM>
__global__ void __launch_bounds__(256, CHUNKS == 1 ? WM_LN_MINB : 1)
add_layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ dres, const T* __restrict__ x,
                         const float* __restrict__ mean, const float* __restrict__ rstd, const T* __restrict__ gamma,
                         T* __restrict__ dx, float* __restrict__ part, long rows, int dim, int nwhich) {
    extern __shared__ float red[];          // [8 warps][nwhich][dim]; nwhich = 3 adds the column sums of dx
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float g[CHUNKS][8], dg[CHUNKS][8], db[CHUNKS][8], ds[CHUNKS][8];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int col = c * 256 + lane * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) { dg[c][i] = 0.f; db[c][i] = 0.f; ds[c][i] = 0.f; g[c][i] = 0.f; }
        if (col < dim) load8(gamma + col, g[c]);
    }
    const float inv_dim = 1.f / dim;
    // The kernel is a pure stream, so what matters is the number of bytes each warp keeps in flight: the (raw,
    // unconverted) loads of the NEXT row(s) are issued before the current ones are reduced.
    constexpr int R = CHUNKS == 1 ? WM_LN_R : 1;      // wider rows already carry enough bytes per warp
    const long stride = (long)gridDim.x * 8;
    Raw8<T> n_dy[R][CHUNKS], n_x[R][CHUNKS], n_r[R][CHUNKS];
    float n_mu[R], n_rs[R];
    auto prefetch = [&](long row0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long row = row0 + r * stride;
            if (row < rows) {
                n_mu[r] = mean[row];
                n_rs[r] = rstd[row];
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {
                    const int col = c * 256 + lane * 8;
                    if (col < dim) {
                        n_dy[r][c].load(dy + row * dim + col);
                        n_x[r][c].load(x + row * dim + col);
                        if (dres != nullptr) n_r[r][c].load(dres + row * dim + col);
                    }
                }
            }
        }
    };
    long row0 = (long)blockIdx.x * 8 + warp;
    prefetch(row0);
    for (; row0 < rows; row0 += R * stride) {
        float xh[R][CHUNKS][8], gy[R][CHUNKS][8], rr[R][CHUNKS][8];
        float mu[R], rs[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            mu[r] = n_mu[r];
            rs[r] = n_rs[r];
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
                n_dy[r][c].unpack(gy[r][c]);
                n_x[r][c].unpack(xh[r][c]);
                if (dres != nullptr) n_r[r][c].unpack(rr[r][c]);
            }
        }
        prefetch(row0 + R * stride);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long row = row0 + r * stride;
            if (row >= rows) break;
            float c1 = 0.f, c2 = 0.f;
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
                const int col = c * 256 + lane * 8;
                if (col < dim) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float d = gy[r][c][i];
                        xh[r][c][i] = (xh[r][c][i] - mu[r]) * rs[r];
                        gy[r][c][i] = d * g[c][i];
                        c1 = fmaf(gy[r][c][i], xh[r][c][i], c1);
                        c2 += gy[r][c][i];
                        dg[c][i] = fmaf(d, xh[r][c][i], dg[c][i]);
                        db[c][i] += d;
                    }
                }
            }
            c1 = warp_sum(c1) * inv_dim;
            c2 = warp_sum(c2) * inv_dim;
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
                const int col = c * 256 + lane * 8;
                if (col < dim) {
                    float o[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        o[i] = rs[r] * (gy[r][c][i] - c2 - xh[r][c][i] * c1);
                        if (dres != nullptr) o[i] += rr[r][c][i];
                        if constexpr (DSUM) ds[c][i] += o[i];
                    }
                    store8(dx + row * dim + col, o);
                }
            }
        }
    }
    // block reduction of the parameter gradients
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int col = c * 256 + lane * 8;
        if (col < dim)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                red[(warp * nwhich + 0) * dim + col + i] = dg[c][i];
                red[(warp * nwhich + 1) * dim + col + i] = db[c][i];
                if constexpr (DSUM) red[(warp * nwhich + 2) * dim + col + i] = ds[c][i];
            }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nwhich * dim; e += blockDim.x) {
        const int which = e / dim, col = e - which * dim;
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) a += red[(w * nwhich + which) * dim + col];
        part[((long)blockIdx.x * nwhich + which) * dim + col] = a;
    }
}

// out[which][col] = sum_b part[b][which][col]   (final stage of the parameter-gradient reduction)
// block = 32 columns x 32 partial-lanes; the partial index runs across warps (four independent loads in flight per
// thread), then shared memory.  Fixed summation order: the result is deterministic.
template <typename T>
__global__ void __launch_bounds__(1024)
reduce_partials_kernel(const float* __restrict__ part, T* __restrict__ out0, T* __restrict__ out1, T* __restrict__ out2,
                       int nblocks, int dim, int nwhich) {
    __shared__ float red[32][33];
    const int c = threadIdx.x & 31, l = threadIdx.x >> 5;
    const int e = blockIdx.x * 32 + c;                  // flattened (which, col)
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (e < nwhich * dim) {
        const int which = e / dim, col = e - which * dim;
        const float* src = part + (long)which * dim + col;
        const long bs = (long)nwhich * dim;
        int b = l;
        for (; b + 96 < nblocks; b += 128) {
            a0 += src[(long)b * bs];
            a1 += src[(long)(b + 32) * bs];
            a2 += src[(long)(b + 64) * bs];
            a3 += src[(long)(b + 96) * bs];
        }
        for (; b < nblocks; b += 32) a0 += src[(long)b * bs];
    }
    red[l][c] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (l == 0 && e < nwhich * dim) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) s += red[i][c];
        const int which = e / dim, col = e - which * dim;
        T* out = which == 0 ? out0 : (which == 1 ? out1 : out2);
        if constexpr (sizeof(T) == 2) out[col] = __float2bfloat16_rn(s);
        else out[col] = s;
    }
}

// per-block column sums of a [rows, C] matrix: thread = (column group of 8, row lane)
template <typename T>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const T* __restrict__ a, float* __restrict__ part, long rows, int C) {
    extern __shared__ float red[];          // [row lanes][C]
    const int groups = C / 8;
    const int lanes = blockDim.x / groups;                 // row lanes per block
    const int grp = threadIdx.x % groups, rl = threadIdx.x / groups;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    if (rl < lanes) {
        const long stride = (long)gridDim.x * lanes;
        long row = (long)blockIdx.x * lanes + rl;
        for (; row + 3 * stride < rows; row += 4 * stride) {        // four independent 16-byte loads in flight per thread
            float v[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) load8(a + (row + u * stride) * C + grp * 8, v[u]);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += (v[0][i] + v[1][i]) + (v[2][i] + v[3][i]);
        }
        for (; row < rows; row += stride) {
            float v[8];
            load8(a + row * C + grp * 8, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += v[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) red[rl * C + grp * 8 + i] = acc[i];
    }
    __syncthreads();
    for (int col = threadIdx.x; col < C; col += blockDim.x) {
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += red[l * C + col];
        part[(long)blockIdx.x * C + col] = s;
    }
}

// ---- bias + exact (erf) GELU around the first MLP GEMM (FeedForward, local_3d_attention.py:24-27) ------------------
// forward:  y = gelu(h + b)            h = x W1^T without bias, [rows, C]
// backward: dh = dy * gelu'(h + b), and per-block column sums of dh (= gradient of b) in the same pass
// erf by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7): one exponential, one reciprocal, five FMAs.  The exponential
// exp(-v^2/2) is also the Gaussian of gelu', so the backward needs no second one.  (erff + __expf made these kernels
// instruction-bound: ~45 instructions per element against 6 bytes of traffic.)
// Two elements per instruction: sm_100 issues fp32 add / mul / fma on 64-bit register pairs (FADD2 / FMUL2 / FFMA2).  The scalar
// form cost ~25 issue slots per element against 4-6 bytes of traffic, i.e. these kernels were issue-bound, not HBM-bound.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 pk2c(float c) { return pk2(c, c); }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
struct GeluTerms2 { f32x2 cdf, pdf_v; };      // Phi(v) and v * phi(v) for a pair of values
template <bool kWantPdf>
__device__ __forceinline__ GeluTerms2 gelu_terms2(float v0, float v1) {
    const f32x2 x = pk2(fabsf(v0) * 0.70710678118654752f, fabsf(v1) * 0.70710678118654752f);
    float d0, d1;
    upk2(ffma2(pk2c(0.3275911f), x, pk2c(1.f)), d0, d1);
    const f32x2 t = pk2(rcp_approx(d0), rcp_approx(d1));
    float a0, a1;
    upk2(fmul2(x, fmul2(x, pk2c(-1.4426950408889634f))), a0, a1);      // -x^2 log2(e)
    const f32x2 e = pk2(ex2_approx(a0), ex2_approx(a1));                // exp(-x^2) = exp(-v^2 / 2)
    f32x2 p = ffma2(pk2c(1.061405429f), t, pk2c(-1.453152027f));
    p = ffma2(p, t, pk2c(1.421413741f));
    p = ffma2(p, t, pk2c(-0.284496736f));
    p = ffma2(p, t, pk2c(0.254829592f));
    float e0, e1;
    upk2(ffma2(fmul2(fmul2(p, t), e), pk2c(-1.f), pk2c(1.f)), e0, e1);  // erf(|x|)
    GeluTerms2 r;
    r.cdf = ffma2(pk2(copysignf(e0, v0), copysignf(e1, v1)), pk2c(0.5f), pk2c(0.5f));
    r.pdf_v = kWantPdf ? fmul2(fmul2(pk2(v0, v1), e), pk2c(0.3989422804014327f)) : 0ull;
    return r;
}
// in place on 8 values: v <- gelu(v + b)
__device__ __forceinline__ void gelu8(float (&v)[8], const float (&b)[8]) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        const float v0 = v[i] + b[i], v1 = v[i + 1] + b[i + 1];
        upk2(fmul2(pk2(v0, v1), gelu_terms2<false>(v0, v1).cdf), v[i], v[i + 1]);
    }
}
// g <- g * gelu'(v + b)
__device__ __forceinline__ void gelu_grad8(float (&g)[8], const float (&v)[8], const float (&b)[8]) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        const float v0 = v[i] + b[i], v1 = v[i + 1] + b[i + 1];
        const GeluTerms2 t = gelu_terms2<true>(v0, v1);
        upk2(fmul2(pk2(g[i], g[i + 1]), ffma2(t.pdf_v, pk2c(1.f), t.cdf)), g[i], g[i + 1]);
    }
}

// thread = (column group of 8, row lane); four rows in flight per thread
template <typename T>
__global__ void __launch_bounds__(256)
bias_gelu_fwd_kernel(const T* __restrict__ h, const T* __restrict__ bias, T* __restrict__ y, long rows, int C) {
    const int groups = C / 8;
    const int lanes = blockDim.x / groups;
    const int grp = threadIdx.x % groups, rl = threadIdx.x / groups;
    if (rl >= lanes) return;
    float b[8];
    load8(bias + grp * 8, b);
    const long stride = (long)gridDim.x * lanes;
    long row = (long)blockIdx.x * lanes + rl;
    for (; row + 3 * stride < rows; row += 4 * stride) {
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) load8(h + (row + u * stride) * C + grp * 8, v[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            gelu8(v[u], b);
            store8(y + (row + u * stride) * C + grp * 8, v[u]);
        }
    }
    for (; row < rows; row += stride) {
        float v[8];
        load8(h + row * C + grp * 8, v);
        gelu8(v, b);
        store8(y + row * C + grp * 8, v);
    }
}

// same thread layout as colsum_partial_kernel; per-block column sums of dh go to `part`
template <typename T>
__global__ void __launch_bounds__(256, 4)       // 64 registers: the packed-math form compiled to 73 and lost a resident block per SM
bias_gelu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ h, const T* __restrict__ bias, T* __restrict__ dh,
                     float* __restrict__ part, long rows, int C) {
    extern __shared__ float red[];          // [row lanes][C]
    const int groups = C / 8;
    const int lanes = blockDim.x / groups;
    const int grp = threadIdx.x % groups, rl = threadIdx.x / groups;
    float acc[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    if (rl < lanes) {
        load8(bias + grp * 8, b);
        const long stride = (long)gridDim.x * lanes;
        long row = (long)blockIdx.x * lanes + rl;
        for (; row + 3 * stride < rows; row += 4 * stride) {
            float g[4][8], v[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                load8(dy + (row + u * stride) * C + grp * 8, g[u]);
                load8(h + (row + u * stride) * C + grp * 8, v[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                gelu_grad8(g[u], v[u], b);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    g[u][i] = round_to(g[u][i], dh);       // the value the GEMMs and the column sum both see
                    acc[i] += g[u][i];
                }
                store8(dh + (row + u * stride) * C + grp * 8, g[u]);
            }
        }
        for (; row < rows; row += stride) {
            float g[8], v[8];
            load8(dy + row * C + grp * 8, g);
            load8(h + row * C + grp * 8, v);
            gelu_grad8(g, v, b);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                g[i] = round_to(g[i], dh);
                acc[i] += g[i];
            }
            store8(dh + row * C + grp * 8, g);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) red[rl * C + grp * 8 + i] = acc[i];
    }
    __syncthreads();
    for (int col = threadIdx.x; col < C; col += blockDim.x) {
        float sacc = 0.f;
        for (int l = 0; l < lanes; ++l) sacc += red[l * C + col];
        part[(long)blockIdx.x * C + col] = sacc;
    }
}

template <typename T>
int launch_ln_fwd(const void* res, const void* delta, const void* delta_bias, const void* gamma, const void* beta, void* sum_out, void* y,
                  float* mean, float* rstd, long rows, int dim, float eps, cudaStream_t st) {
    const int chunks = (dim + 255) / 256;
    const unsigned grid = (unsigned)((rows + 7) / 8);
#define WM_LN_FWD(CH)                                                                                           \
    add_layernorm_fwd_kernel<T, CH><<<grid, 256, 0, st>>>(static_cast<const T*>(res), static_cast<const T*>(delta), \
                                                         static_cast<const T*>(delta_bias),                        \
                                                         static_cast<const T*>(gamma), static_cast<const T*>(beta), \
                                                         static_cast<T*>(sum_out), static_cast<T*>(y), mean, rstd, \
                                                         rows, dim, eps)
    switch (chunks) {
        case 1: WM_LN_FWD(1); break;
        case 2: WM_LN_FWD(2); break;
        case 3: case 4: WM_LN_FWD(4); break;
        default: WM_LN_FWD(8); break;
    }
#undef WM_LN_FWD
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}

template <typename T>
int launch_ln_bwd(const void* dy, const void* dres, const void* x, const float* mean, const float* rstd,
                  const void* gamma, void* dx, void* dgamma, void* dbeta, void* dsum, float* part, int nblocks, long rows,
                  int dim, cudaStream_t st) {
    const int chunks = (dim + 255) / 256;
    const int nwhich = dsum != nullptr ? 3 : 2;
    const size_t smem = (size_t)8 * nwhich * dim * sizeof(float);
#define WM_LN_BWD_(CH, DS)                                                                                         \
    do {                                                                                                           \
        WM_CUDA_CHECK(cudaFuncSetAttribute(add_layernorm_bwd_kernel<T, CH, DS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        add_layernorm_bwd_kernel<T, CH, DS><<<nblocks, 256, smem, st>>>(static_cast<const T*>(dy), static_cast<const T*>(dres), \
                                                                      static_cast<const T*>(x), mean, rstd,          \
                                                                      static_cast<const T*>(gamma), static_cast<T*>(dx), part, rows, dim, nwhich); \
    } while (0)
#define WM_LN_BWD(CH) do { if (dsum != nullptr) WM_LN_BWD_(CH, true); else WM_LN_BWD_(CH, false); } while (0)
    switch (chunks) {
        case 1: WM_LN_BWD(1); break;
        case 2: WM_LN_BWD(2); break;
        case 3: case 4: WM_LN_BWD(4); break;
        default: WM_LN_BWD(8); break;
    }
#undef WM_LN_BWD_
#undef WM_LN_BWD
    WM_CUDA_CHECK(cudaGetLastError());
    reduce_partials_kernel<T><<<(nwhich * dim + 31) / 32, 1024, 0, st>>>(part, static_cast<T*>(dgamma), static_cast<T*>(dbeta),
                                                                        static_cast<T*>(dsum), nblocks, dim, nwhich);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}

template <typename T>
int launch_colsum(const void* a, void* out, float* part, int nblocks, long rows, int C, cudaStream_t st) {
    const size_t smem = (size_t)(256 / (C / 8) > 0 ? 256 / (C / 8) : 1) * C * sizeof(float);
    WM_CUDA_CHECK(cudaFuncSetAttribute(colsum_partial_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    colsum_partial_kernel<T><<<nblocks, 256, smem, st>>>(static_cast<const T*>(a), part, rows, C);
    WM_CUDA_CHECK(cudaGetLastError());
    reduce_partials_kernel<T><<<(C + 31) / 32, 1024, 0, st>>>(part, static_cast<T*>(out), static_cast<T*>(out), static_cast<T*>(out), nblocks, C, 1);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}

int check_rows(const char* what, long rows, int dim, int dtype) {
    if (dtype != WM_DTYPE_BF16 && dtype != WM_DTYPE_FP32) return fail(WM_EINVAL, "%s: unknown dtype %d", what, dtype);
    if (rows < 0 || dim <= 0) return fail(WM_EINVAL, "%s: bad sizes rows=%ld dim=%d", what, rows, dim);
    if (dim % 8 != 0 || dim > 2048) return fail(WM_EUNSUPPORTED, "%s: dim=%d must be a multiple of 8, at most 2048", what, dim);
    return WM_OK;
}

}  // namespace
}  // namespace wm

using namespace wm;

extern "C" int wm_reduce_blocks(long rows) {
    long b = (rows + 63) / 64;
    if (b > (long)sm_count() * 4) b = (long)sm_count() * 4;
    if (b < 1) b = 1;
    return (int)b;
}

extern "C" int wm_add_layernorm_fwd(const void* res, const void* delta, const void* delta_bias, const void* gamma,
                                    const void* beta, void* sum_out, void* y, float* mean, float* rstd, long rows, int dim,
                                    float eps, int dtype, void* stream) {
    if (int rc = check_rows("wm_add_layernorm_fwd", rows, dim, dtype)) return rc;
    if (rows == 0) return WM_OK;
    if (!res || !gamma || !beta || !y || !mean || !rstd) return fail(WM_EINVAL, "wm_add_layernorm_fwd: null pointer");
    if (!aligned16(res) || !aligned16(y) || (delta && !aligned16(delta)) || (sum_out && !aligned16(sum_out)) ||
        !aligned16(gamma) || !aligned16(beta) || (delta_bias && !aligned16(delta_bias)))
        return fail(WM_EINVAL, "wm_add_layernorm_fwd: pointers must be 16-byte aligned");
    if (delta_bias && !delta) return fail(WM_EINVAL, "wm_add_layernorm_fwd: delta_bias needs delta");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return dtype == WM_DTYPE_BF16
               ? launch_ln_fwd<__nv_bfloat16>(res, delta, delta_bias, gamma, beta, sum_out, y, mean, rstd, rows, dim, eps, st)
               : launch_ln_fwd<float>(res, delta, delta_bias, gamma, beta, sum_out, y, mean, rstd, rows, dim, eps, st);
}

extern "C" int wm_add_layernorm_bwd(const void* dy, const void* dres, const void* x, const float* mean,
                                    const float* rstd, const void* gamma, void* dx, void* dgamma, void* dbeta,
                                    void* dsum, float* workspace, long rows, int dim, int dtype, void* stream) {
    if (int rc = check_rows("wm_add_layernorm_bwd", rows, dim, dtype)) return rc;
    if (rows == 0) return WM_OK;
    if (!dy || !x || !mean || !rstd || !gamma || !dx || !dgamma || !dbeta || !workspace)
        return fail(WM_EINVAL, "wm_add_layernorm_bwd: null pointer");
    if (!aligned16(dy) || !aligned16(x) || !aligned16(dx) || (dres && !aligned16(dres)) || !aligned16(gamma))
        return fail(WM_EINVAL, "wm_add_layernorm_bwd: pointers must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nb = wm_reduce_blocks(rows);
    return dtype == WM_DTYPE_BF16
               ? launch_ln_bwd<__nv_bfloat16>(dy, dres, x, mean, rstd, gamma, dx, dgamma, dbeta, dsum, workspace, nb, rows, dim, st)
               : launch_ln_bwd<float>(dy, dres, x, mean, rstd, gamma, dx, dgamma, dbeta, dsum, workspace, nb, rows, dim, st);
}

extern "C" int wm_colsum(const void* a, void* out, float* workspace, long rows, int cols, int dtype, void* stream) {
    if (int rc = check_rows("wm_colsum", rows, cols, dtype)) return rc;
    if (!a || !out || !workspace) return fail(WM_EINVAL, "wm_colsum: null pointer");
    if (!aligned16(a)) return fail(WM_EINVAL, "wm_colsum: input must be 16-byte aligned");
    if (cols / 8 > 256) return fail(WM_EUNSUPPORTED, "wm_colsum: at most 2048 columns");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nb = wm_reduce_blocks(rows);
    return dtype == WM_DTYPE_BF16 ? launch_colsum<__nv_bfloat16>(a, out, workspace, nb, rows, cols, st)
                                  : launch_colsum<float>(a, out, workspace, nb, rows, cols, st);
}

namespace wm {
namespace {
template <typename T>
int launch_bias_gelu_fwd(const void* h, const void* bias, void* y, long rows, int C, cudaStream_t st) {
    const int lanes = 256 / (C / 8) > 0 ? 256 / (C / 8) : 1;
    long blocks = (rows + 4L * lanes - 1) / (4L * lanes);          // four rows per thread
    if (blocks > (long)sm_count() * 8) blocks = (long)sm_count() * 8;
    if (blocks < 1) blocks = 1;
    bias_gelu_fwd_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const T*>(h), static_cast<const T*>(bias),
                                                              static_cast<T*>(y), rows, C);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}
template <typename T>
int launch_bias_gelu_bwd(const void* dy, const void* h, const void* bias, void* dh, void* dbias, float* part, int nblocks,
                         long rows, int C, cudaStream_t st) {
    const size_t smem = (size_t)(256 / (C / 8) > 0 ? 256 / (C / 8) : 1) * C * sizeof(float);
    WM_CUDA_CHECK(cudaFuncSetAttribute(bias_gelu_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bias_gelu_bwd_kernel<T><<<nblocks, 256, smem, st>>>(static_cast<const T*>(dy), static_cast<const T*>(h),
                                                        static_cast<const T*>(bias), static_cast<T*>(dh), part, rows, C);
    WM_CUDA_CHECK(cudaGetLastError());
    reduce_partials_kernel<T><<<(C + 31) / 32, 1024, 0, st>>>(part, static_cast<T*>(dbias), static_cast<T*>(dbias),
                                                              static_cast<T*>(dbias), nblocks, C, 1);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}
}  // namespace
}  // namespace wm

extern "C" int wm_bias_gelu_fwd(const void* h, const void* bias, void* y, long rows, int cols, int dtype, void* stream) {
    if (int rc = check_rows("wm_bias_gelu_fwd", rows, cols, dtype)) return rc;
    if (rows == 0) return WM_OK;
    if (!h || !bias || !y) return fail(WM_EINVAL, "wm_bias_gelu_fwd: null pointer");
    if (!aligned16(h) || !aligned16(bias) || !aligned16(y)) return fail(WM_EINVAL, "wm_bias_gelu_fwd: pointers must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return dtype == WM_DTYPE_BF16 ? launch_bias_gelu_fwd<__nv_bfloat16>(h, bias, y, rows, cols, st)
                                  : launch_bias_gelu_fwd<float>(h, bias, y, rows, cols, st);
}

extern "C" int wm_bias_gelu_bwd(const void* dy, const void* h, const void* bias, void* dh, void* dbias, float* workspace,
                                long rows, int cols, int dtype, void* stream) {
    if (int rc = check_rows("wm_bias_gelu_bwd", rows, cols, dtype)) return rc;
    if (!dy || !h || !bias || !dh || !dbias || !workspace) return fail(WM_EINVAL, "wm_bias_gelu_bwd: null pointer");
    if (!aligned16(dy) || !aligned16(h) || !aligned16(bias) || !aligned16(dh))
        return fail(WM_EINVAL, "wm_bias_gelu_bwd: pointers must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nb = wm_reduce_blocks(rows);
    return dtype == WM_DTYPE_BF16 ? launch_bias_gelu_bwd<__nv_bfloat16>(dy, h, bias, dh, dbias, workspace, nb, rows, cols, st)
                                  : launch_bias_gelu_bwd<float>(dy, h, bias, dh, dbias, workspace, nb, rows, cols, st);
}
