// VQ nearest-codebook search, exact path.
//
// Replaces the distance / argmin / gather / per-latent-error part of
// VectorQuantizerEMA.forward and .encode (vq-video-diffusion/vq.py:30-36, 84-87), which
// in the reference materialises an [N,L,D,K] fp32 temporary (512x the input).
//
// Stage 1 (all codes): register-tiled direct-form distances sum_d (x_d - e_d)^2 in fp32,
//   64 latents x 64 codes per step out of shared memory, keeping the best and the
//   runner-up (value, index) per latent.  All terms are non-negative, so the fp32 value
//   is within (D+2)*2^-24 relative of the exact distance.
// Stage 2 (rare): a latent whose runner-up lies inside that error band -- including exact
//   duplicates in the codebook -- is re-scanned over all K codes with fp64 accumulation,
//   lowest index winning ties.  The returned index is therefore the exact-arithmetic
//   argmin with ATen's first-minimum tie rule (vq.py:33).
// Epilogue: quantized = x + (e - x) (the straight-through forward value, vq.py:70) and
//   sq_err = sum_d (e - x)^2 (vq.py:35).
#include "wm_common.cuh"

#include <math.h>

namespace wm {
namespace {

constexpr int kTileN = 64;      // latents per block
constexpr int kTileK = 64;      // codes per shared-memory step
constexpr int kThreads = 128;
constexpr int kPad = 4;         // floats of row padding: conflict-free LDS.128

struct Best2 {
    float d0, d1;               // best, runner-up
    int i0, i1;
};

__device__ __forceinline__ void best2_insert(Best2& b, float d, int i) {
    // strict ordering by (distance, index): equal distances keep the lower index first
    if (d < b.d0 || (d == b.d0 && i < b.i0)) {
        b.d1 = b.d0; b.i1 = b.i0;
        b.d0 = d; b.i0 = i;
    } else if (d < b.d1 || (d == b.d1 && i < b.i1)) {
        b.d1 = d; b.i1 = i;
    }
}

__device__ __forceinline__ Best2 best2_merge(Best2 a, const Best2& o) {
    best2_insert(a, o.d0, o.i0);
    best2_insert(a, o.d1, o.i1);
    return a;
}

__global__ void __launch_bounds__(kThreads)
vq_nearest_kernel(const float* __restrict__ x, const float* __restrict__ cb, int64_t* __restrict__ idx,
                  float* __restrict__ quantized, float* __restrict__ sq_err, long N, int L, int K, int D) {
    extern __shared__ __align__(16) float smem[];
    const int ld = D + kPad;
    float* xs = smem;                       // [kTileN][ld]
    float* es = xs + kTileN * ld;           // [kTileK][ld]
    __shared__ int s_idx[kTileN];
    __shared__ int s_ambiguous[kTileN];
    __shared__ double s_red_d[kThreads];
    __shared__ int s_red_i[kThreads];

    const int tid = threadIdx.x;
    const int l = blockIdx.y;
    const long n0 = (long)blockIdx.x * kTileN;
    const float* cbl = cb + (long)l * K * D;
    const int tx = tid & 7;                 // codes  tx + 8*j
    const int ty = tid >> 3;                // latents ty + 16*i

    // stage the latent tile (rows beyond N are zero-filled and never written back)
    const int vecs_per_row = D / 4;
    for (int e = tid; e < kTileN * vecs_per_row; e += kThreads) {
        const int r = e / vecs_per_row, c = (e % vecs_per_row) * 4;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n0 + r < N) val = __ldg(reinterpret_cast<const float4*>(x + ((n0 + r) * L + l) * (long)D + c));
        *reinterpret_cast<float4*>(xs + r * ld + c) = val;
    }

    Best2 best[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) best[i] = Best2{INFINITY, INFINITY, 0x7fffffff, 0x7fffffff};

    for (int k0 = 0; k0 < K; k0 += kTileK) {
        __syncthreads();                    // previous step done with es (and xs staged)
        for (int e = tid; e < kTileK * vecs_per_row; e += kThreads) {
            const int r = e / vecs_per_row, c = (e % vecs_per_row) * 4;
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + r < K) val = __ldg(reinterpret_cast<const float4*>(cbl + (long)(k0 + r) * D + c));
            *reinterpret_cast<float4*>(es + r * ld + c) = val;
        }
        __syncthreads();

        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        for (int c = 0; c < D; c += 4) {
            float4 xv[4], ev[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(xs + (ty + 16 * i) * ld + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) ev[j] = *reinterpret_cast<const float4*>(es + (tx + 8 * j) * ld + c);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float t;
                    t = xv[i].x - ev[j].x; acc[i][j] = fmaf(t, t, acc[i][j]);
                    t = xv[i].y - ev[j].y; acc[i][j] = fmaf(t, t, acc[i][j]);
                    t = xv[i].z - ev[j].z; acc[i][j] = fmaf(t, t, acc[i][j]);
                    t = xv[i].w - ev[j].w; acc[i][j] = fmaf(t, t, acc[i][j]);
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int code = k0 + tx + 8 * j;
                if (code < K) best2_insert(best[i], acc[i][j], code);
            }
    }

    // merge the 8 threads (tx) that share a latent: they are 8 consecutive lanes
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            Best2 other;
            other.d0 = __shfl_xor_sync(0xffffffffu, best[i].d0, o);
            other.d1 = __shfl_xor_sync(0xffffffffu, best[i].d1, o);
            other.i0 = __shfl_xor_sync(0xffffffffu, best[i].i0, o);
            other.i1 = __shfl_xor_sync(0xffffffffu, best[i].i1, o);
            best[i] = best2_merge(best[i], other);
        }
        if (tx == 0) {
            const int r = ty + 16 * i;
            // fp32 direct form: |d32 - d| <= (D+2) 2^-24 d for both codes -> 4x safety band
            const float band = 4.f * (float)(D + 2) * 5.9604645e-8f;
            s_idx[r] = best[i].i0;
            s_ambiguous[r] = (K > 1 && best[i].d1 <= best[i].d0 * (1.f + band) + 1e-37f) ? 1 : 0;
        }
    }
    __syncthreads();

    // stage 2: fp64 re-scan of ambiguous latents (block-uniform loop)
    for (int r = 0; r < kTileN; ++r) {
        if (!s_ambiguous[r] || n0 + r >= N) continue;
        double bd = INFINITY;
        int bi = 0x7fffffff;
        for (int code = tid; code < K; code += kThreads) {
            const float* e = cbl + (long)code * D;
            double a = 0.0;
            for (int c = 0; c < D; ++c) {
                const double t = (double)xs[r * ld + c] - (double)__ldg(e + c);
                a = fma(t, t, a);
            }
            if (a < bd) { bd = a; bi = code; }          // ascending codes: first minimum kept
        }
        s_red_d[tid] = bd;
        s_red_i[tid] = bi;
        __syncthreads();
        for (int o = kThreads / 2; o > 0; o >>= 1) {
            if (tid < o) {
                const double od = s_red_d[tid + o];
                const int oi = s_red_i[tid + o];
                if (od < s_red_d[tid] || (od == s_red_d[tid] && oi < s_red_i[tid])) {
                    s_red_d[tid] = od;
                    s_red_i[tid] = oi;
                }
            }
            __syncthreads();
        }
        if (tid == 0) s_idx[r] = s_red_i[0];
        __syncthreads();
    }

    // epilogue: indices, straight-through value, per-latent squared error
    for (int r = tid; r < kTileN; r += kThreads)
        if (n0 + r < N) idx[(n0 + r) * L + l] = (int64_t)s_idx[r];
    if (quantized != nullptr || sq_err != nullptr) {
        const int lane = tid & 31, warp = tid >> 5;
        for (int r = warp; r < kTileN; r += kThreads / 32) {
            if (n0 + r >= N) continue;                  // warp-uniform
            const float* e = cbl + (long)s_idx[r] * D;
            float err = 0.f;
            for (int c = lane; c < D; c += 32) {
                const float xv = xs[r * ld + c];
                const float diff = __ldg(e + c) - xv;
                err = fmaf(diff, diff, err);
                if (quantized != nullptr) quantized[((n0 + r) * L + l) * (long)D + c] = xv + diff;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
            if (lane == 0 && sq_err != nullptr) sq_err[(n0 + r) * L + l] = err;
        }
    }
}

__global__ void vq_distance_kernel(const float* __restrict__ x, const float* __restrict__ cb,
                                   float* __restrict__ dist, long N, int L, int K, int D, float mul) {
    const long total = N * L * (long)K;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int k = (int)(e % K);
        const long nl = e / K;
        const int l = (int)(nl % L);
        const float* xr = x + nl * D;
        const float* er = cb + ((long)l * K + k) * D;
        float a = 0.f;
        for (int c = 0; c < D; ++c) {
            const float t = __ldg(xr + c) - __ldg(er + c);
            a = fmaf(t, t, a);
        }
        dist[e] = a * mul;
    }
}

}  // namespace
}  // namespace wm

using namespace wm;

extern "C" int wm_vq_nearest(const void* x, const void* codebook, int64_t* idx, void* quantized, float* sq_err,
                             long N, int L, int K, int D, int dtype, int flags, void* stream) {
    if (dtype != WM_DTYPE_FP32) return fail(WM_EUNSUPPORTED, "wm_vq_nearest: only fp32 latents are supported");
    if (N < 0 || L <= 0 || K <= 0 || D <= 0) return fail(WM_EINVAL, "wm_vq_nearest: bad sizes N=%ld L=%d K=%d D=%d", N, L, K, D);
    if (N == 0) return WM_OK;
    if (!x || !codebook || !idx) return fail(WM_EINVAL, "wm_vq_nearest: null pointer");
    if (D % 4 != 0 || D > 1024) return fail(WM_EUNSUPPORTED, "wm_vq_nearest: D=%d must be a multiple of 4, at most 1024", D);
    if (!aligned16(x) || !aligned16(codebook)) return fail(WM_EINVAL, "wm_vq_nearest: x / codebook must be 16-byte aligned");
    if (L > 65535) return fail(WM_EUNSUPPORTED, "wm_vq_nearest: L=%d too large", L);
    if (!(flags & WM_FLAG_SIMT) && vq_tc_supported(N, L, K, D))          // tensor-core filter + exact re-check
        return vq_nearest_tc(x, codebook, idx, quantized, sq_err, N, L, K, D, (cudaStream_t)stream);
    const size_t smem = (size_t)(kTileN + kTileK) * (D + kPad) * sizeof(float);
    if (smem > 200 * 1024) return fail(WM_EUNSUPPORTED, "wm_vq_nearest: D=%d needs %zu bytes of shared memory", D, smem);
    WM_CUDA_CHECK(cudaFuncSetAttribute(vq_nearest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((N + kTileN - 1) / kTileN), (unsigned)L);
    vq_nearest_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(
        static_cast<const float*>(x), static_cast<const float*>(codebook), idx, static_cast<float*>(quantized), sq_err,
        N, L, K, D);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}

extern "C" int wm_vq_distance(const void* x, const void* codebook, float* dist, long N, int L, int K, int D,
                              int normalize, void* stream) {
    if (N < 0 || L <= 0 || K <= 0 || D <= 0) return fail(WM_EINVAL, "wm_vq_distance: bad sizes");
    if (N == 0) return WM_OK;
    if (!x || !codebook || !dist) return fail(WM_EINVAL, "wm_vq_distance: null pointer");
    const long total = N * L * (long)K;
    const long cap = (long)sm_count() * 32;
    const unsigned grid = (unsigned)((total + 255) / 256 < cap ? (total + 255) / 256 : cap);
    vq_distance_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<const float*>(x),
                                                               static_cast<const float*>(codebook), dist, N, L, K, D,
                                                               normalize ? 1.f / (float)D : 1.f);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}
