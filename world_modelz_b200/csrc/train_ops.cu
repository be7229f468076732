// Device-side pieces of the training / sampling loops around the denoiser, so that a step never leaves the GPU:
//   wm_vq_stats          codebook statistics of VectorQuantizerEMA.forward in training mode (vq.py:35-46) without the
//                        dense one-hot: counts, per-code sums of the latents (dw) and accumulated error in one pass
//   wm_vq_onehot         the float one-hot `encodings` tensor the reference returns (vq.py:39), written directly
//   wm_sample_step       one draw of the mask/replace sampler (main.py:80-109): optional top-k filter, softmax +
//                        multinomial as a Gumbel-max, re-masking -- one warp per position, no [P,K] temporaries
//   wm_loss_hist_update  LossAwareSamplerEma.update_with_losses (importance_sampling.py:35-41) on the device
//   wm_colsq             (see optim.cu) squared gradient norm for main.py:189-193
#include "wm_common.cuh"

#include <math.h>

namespace wm {
namespace {

// ------------------------------------------------------------------------------------------ VQ statistics
// One CTA accumulates its slice of the latents into shared memory (counts, error, and the [K, D] sums when they
// fit), then adds its non-zero partials to global memory.  fp32 atomics: sums are exact per code up to ordering.
__global__ void __launch_bounds__(256)
vq_stats_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, const float* __restrict__ err,
                float* __restrict__ counts, float* __restrict__ dw, float* __restrict__ acc_err, long N, int L, int K,
                int D, int dw_in_smem) {
    extern __shared__ float sm[];
    float* s_cnt = sm;                  // [K]
    float* s_err = sm + K;              // [K]
    float* s_dw = sm + 2 * K;           // [K * D] (if dw_in_smem)
    const int l = blockIdx.y;
    const int tid = threadIdx.x;
    for (int i = tid; i < 2 * K + (dw_in_smem ? K * D : 0); i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const long per = (N + gridDim.x - 1) / gridDim.x;
    const long n0 = (long)blockIdx.x * per, n1 = min(N, n0 + per);
    // a warp walks latents; lanes run over channels
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    for (long n = n0 + warp; n < n1; n += nwarps) {
        const int k = (int)idx[n * L + l];
        if (lane == 0) {
            atomicAdd(&s_cnt[k], 1.f);
            if (err != nullptr) atomicAdd(&s_err[k], err[n * L + l]);
        }
        if (dw != nullptr) {
            const float* xr = x + (n * L + l) * (long)D;
            float* dst = dw_in_smem ? (s_dw + (long)k * D) : (dw + ((long)l * K + k) * D);
            for (int c = lane; c < D; c += 32) atomicAdd(&dst[c], xr[c]);
        }
    }
    __syncthreads();
    for (int k = tid; k < K; k += blockDim.x) {
        if (s_cnt[k] != 0.f) {
            atomicAdd(&counts[(long)l * K + k], s_cnt[k]);
            if (acc_err != nullptr) atomicAdd(&acc_err[(long)l * K + k], s_err[k]);
        }
    }
    if (dw != nullptr && dw_in_smem)
        for (int i = tid; i < K * D; i += blockDim.x)
            if (s_cnt[i / D] != 0.f) atomicAdd(&dw[(long)l * K * D + i], s_dw[i]);
}

__global__ void __launch_bounds__(256)
vq_onehot_kernel(const int64_t* __restrict__ idx, float* __restrict__ out, long rows, int K) {
    // rows = N*L; each row K floats; 4 floats per thread-iteration
    const long total4 = rows * (long)(K / 4);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
        const long r = i / (K / 4);
        const int c = (int)(i - r * (K / 4)) * 4;
        const int k = (int)idx[r];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k >= c && k < c + 4) (&v.x)[k - c] = 1.f;
        reinterpret_cast<float4*>(out)[i] = v;
    }
}

// ------------------------------------------------------------------------------------------ sampler step
// Philox-4x32-10 (Salmon et al., SC'11): counter-based, so that any (call, position, code) has its own stream.
__device__ __forceinline__ uint4 philox(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
__device__ __forceinline__ float u01(uint32_t r) { return ((r >> 8) + 0.5f) * (1.f / 16777216.f); }   // (0, 1)

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// one warp per position: logits row [K] -> sample (Gumbel-max over the top-k-filtered row == multinomial(softmax)),
// frame = mask token where a uniform draw exceeds alpha (main.py:86-109).  dyn = {alpha, call counter} on the device.
template <typename T>
__global__ void __launch_bounds__(128)
sample_step_kernel(const T* __restrict__ logits, int64_t* __restrict__ sample, int64_t* __restrict__ frame, long frame_stride,
                   long per_clip, long P, int K, int topk, int mask_token, const float* __restrict__ dyn, uint64_t seed) {
    const int lane = threadIdx.x & 31;
    const long p = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= P) return;
    const float alpha = dyn[0];
    const uint32_t call = (uint32_t)dyn[1];
    const T* row = logits + p * (long)K;
    float kth = -INFINITY;
    if (topk > 0 && topk < K) {
        // k-th largest value by bisection on the order-preserving integer image of the floats
        auto key_of = [](float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); };
        uint32_t lo = 0u, hi = 0xffffffffu;              // invariant: count(key >= lo) >= topk, count(key > hi) < topk
        while (lo < hi) {
            const uint32_t mid = lo + (uint32_t)(((uint64_t)hi - lo + 1) >> 1);
            int c = 0;
            for (int j = lane; j < K; j += 32) c += key_of(ldf(row + j)) >= mid;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (c >= topk) lo = mid; else hi = mid - 1;
        }
        const uint32_t u = (lo & 0x80000000u) ? (lo & 0x7fffffffu) : ~lo;
        kth = __uint_as_float(u);
    }
    float best = -INFINITY;
    int best_j = 0x7fffffff;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    for (int j0 = lane * 4; j0 < K; j0 += 128) {
        const uint4 r = philox(make_uint4((uint32_t)p, (uint32_t)(p >> 32), call, (uint32_t)(j0 >> 2)), key);
        const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = j0 + e;
            if (j < K) {
                const float lg = ldf(row + j);
                if (lg >= kth) {                                   // main.py:42: logits < kth -> -inf
                    const float g = lg - __logf(-__logf(u01(rr[e])));
                    if (g > best || (g == best && j < best_j)) { best = g; best_j = j; }
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
        if (ob > best || (ob == best && oj < best_j)) { best = ob; best_j = oj; }
    }
    if (lane == 0) {
        const uint4 r = philox(make_uint4((uint32_t)p, (uint32_t)(p >> 32), call, 0xffffffffu), key);
        const bool remask = u01(r.x) > alpha;
        sample[p] = best_j;
        if (frame != nullptr) {
            const long clip = p / per_clip, pos = p - clip * per_clip;
            frame[clip * frame_stride + pos] = remask ? mask_token : best_j;
        }
    }
}

// weights[b] = weights[b] * alpha + loss * (1 - alpha), in sample order within a bucket (importance_sampling.py:40-41)
__global__ void loss_hist_update_kernel(const float* __restrict__ ts, const float* __restrict__ losses, float* __restrict__ weights,
                                        int64_t* __restrict__ counts, int B, int nb, float alpha) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    float w = weights[b];
    long c = counts[b];
    for (int i = 0; i < B; ++i) {
        int bi = (int)(ts[i] * nb);
        bi = bi < 0 ? 0 : (bi > nb - 1 ? nb - 1 : bi);
        if (bi == b) { w = w * alpha + losses[i] * (1.f - alpha); ++c; }
    }
    weights[b] = w;
    counts[b] = c;
}

// ------------------------------------------------------------------------------ token + position embedding
// x[b,s,h,w,:] = E[token] + ((Ps[s] + Ph[h]) + Pw[w])  -- Local3dAttentionTransformer.forward's first line
// (local_3d_attention.py:149-157) in one pass: no gathered [B,S,H,W,dim] temporary, no broadcast add.  The three
// sums round to the storage type one after the other, exactly like the chain of stock ops they replace.
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ float round_to(float v, const __nv_bfloat16*) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float round_to(float v, const float*) { return v; }
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = v;
}
__device__ __forceinline__ void store8(float* p, const float (&f)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

// A group of dim/8 consecutive threads writes one token row (16 bytes per thread); token coordinates are decoded once per
// row in 32-bit arithmetic, position rows come from L1/L2.
template <typename T>
__global__ void __launch_bounds__(256)
embed_pos_kernel(const int64_t* __restrict__ tokens, const T* __restrict__ table, const T* __restrict__ ps,
                 const T* __restrict__ ph, const T* __restrict__ pw, T* __restrict__ out, long ntok, int S, int H, int W,
                 int dim, int num_rows) {
    const int vecs = dim >> 3;
    const int rows_per_block = 256 / vecs;                 // host guarantees vecs <= 256
    const int r = threadIdx.x / vecs, c = (threadIdx.x - r * vecs) * 8;
    if (r >= rows_per_block) return;
    const T* tag = nullptr;
    for (long tok = (long)blockIdx.x * rows_per_block + r; tok < ntok; tok += (long)gridDim.x * rows_per_block) {
        const unsigned in_clip = (unsigned)(tok % ((long)S * H * W));
        const unsigned w = in_clip % (unsigned)W, hs = in_clip / (unsigned)W;
        const unsigned h = hs % (unsigned)H, s = hs / (unsigned)H;
        long row = tokens[tok];
        row = row < 0 ? 0 : (row >= num_rows ? num_rows - 1 : row);
        float e[8], a[8], b2[8], c2[8], o[8];
        load8(table + row * dim + c, e);
        load8(ps + (long)s * dim + c, a);
        load8(ph + (long)h * dim + c, b2);
        load8(pw + (long)w * dim + c, c2);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = e[k] + round_to(round_to(a[k] + b2[k], tag) + c2[k], tag);
        store8(out + tok * dim + c, o);
    }
}

}  // namespace
}  // namespace wm

using namespace wm;

extern "C" int wm_vq_stats(const void* x, const int64_t* idx, const float* sq_err, float* counts, float* dw,
                           float* acc_err, long N, int L, int K, int D, void* stream) {
    if (N < 0 || L <= 0 || K <= 0 || D <= 0) return fail(WM_EINVAL, "wm_vq_stats: bad shape N=%ld L=%d K=%d D=%d", N, L, K, D);
    if (N == 0) return WM_OK;
    if (!idx || !counts || (dw && !x)) return fail(WM_EINVAL, "wm_vq_stats: null pointer");
    if (L > 65535) return fail(WM_EUNSUPPORTED, "wm_vq_stats: L=%d", L);
    const int sms = sm_count();
    const size_t small = (size_t)2 * K * sizeof(float), full = small + (size_t)K * D * sizeof(float);
    const int dw_in_smem = (dw != nullptr && full <= 200 * 1024) ? 1 : 0;
    const size_t smem = dw_in_smem ? full : small;
    if (smem > 200 * 1024) return fail(WM_EUNSUPPORTED, "wm_vq_stats: K=%d too large", K);
    WM_CUDA_CHECK(cudaFuncSetAttribute(vq_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long ctas = (N + 255) / 256;
    const long cap = sms / L > 0 ? sms / L : 1;
    if (ctas > cap) ctas = cap;
    vq_stats_kernel<<<dim3((unsigned)ctas, (unsigned)L), 256, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float*>(x), idx, sq_err, counts, dw, acc_err, N, L, K, D, dw_in_smem);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}

extern "C" int wm_vq_onehot(const int64_t* idx, float* encodings, long rows, int K, void* stream) {
    if (rows < 0 || K <= 0 || K % 4 != 0) return fail(WM_EINVAL, "wm_vq_onehot: rows=%ld K=%d (K must be a multiple of 4)", rows, K);
    if (rows == 0) return WM_OK;
    if (!idx || !encodings || !aligned16(encodings)) return fail(WM_EINVAL, "wm_vq_onehot: null or misaligned pointer");
    long blocks = (rows * (K / 4) + 255) / 256;
    if (blocks > (long)sm_count() * 32) blocks = (long)sm_count() * 32;
    vq_onehot_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(idx, encodings, rows, K);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}

extern "C" int wm_sample_step(const void* logits, int64_t* sample, int64_t* frame, long frame_stride, long per_clip,
                              long P, int K, int topk, int mask_token, const float* dyn, uint64_t seed, int dtype,
                              void* stream) {
    if (P < 0 || K <= 0 || per_clip <= 0) return fail(WM_EINVAL, "wm_sample_step: P=%ld K=%d per_clip=%ld", P, K, per_clip);
    if (P == 0) return WM_OK;
    if (!logits || !sample || !dyn) return fail(WM_EINVAL, "wm_sample_step: null pointer");
    if (dtype != WM_DTYPE_BF16 && dtype != WM_DTYPE_FP32) return fail(WM_EINVAL, "wm_sample_step: dtype %d", dtype);
    const unsigned blocks = (unsigned)((P + 3) / 4);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == WM_DTYPE_FP32)
        sample_step_kernel<float><<<blocks, 128, 0, st>>>(static_cast<const float*>(logits), sample, frame, frame_stride, per_clip,
                                                         P, K, topk, mask_token, dyn, seed);
    else
        sample_step_kernel<__nv_bfloat16><<<blocks, 128, 0, st>>>(static_cast<const __nv_bfloat16*>(logits), sample, frame,
                                                                 frame_stride, per_clip, P, K, topk, mask_token, dyn, seed);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}

extern "C" int wm_loss_hist_update(const float* ts, const float* losses, float* weights, int64_t* counts, int B, int buckets,
                                   float alpha, void* stream) {
    if (B < 0 || buckets <= 0) return fail(WM_EINVAL, "wm_loss_hist_update: B=%d buckets=%d", B, buckets);
    if (!ts || !losses || !weights || !counts) return fail(WM_EINVAL, "wm_loss_hist_update: null pointer");
    loss_hist_update_kernel<<<(buckets + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(ts, losses, weights, counts, B,
                                                                                                 buckets, alpha);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}

extern "C" int wm_embed_pos_fwd(const int64_t* tokens, const void* table, const void* pos_s, const void* pos_h,
                                const void* pos_w, void* out, long B, int S, int H, int W, int dim, int num_rows, int dtype,
                                void* stream) {
    if (B < 0 || S <= 0 || H <= 0 || W <= 0 || dim <= 0 || num_rows <= 0)
        return fail(WM_EINVAL, "wm_embed_pos_fwd: bad shape B=%ld S=%d H=%d W=%d dim=%d rows=%d", B, S, H, W, dim, num_rows);
    if (B == 0) return WM_OK;
    if (!tokens || !table || !pos_s || !pos_h || !pos_w || !out) return fail(WM_EINVAL, "wm_embed_pos_fwd: null pointer");
    if (dim % 8 != 0) return fail(WM_EUNSUPPORTED, "wm_embed_pos_fwd: dim=%d must be a multiple of 8", dim);
    if (dtype != WM_DTYPE_BF16 && dtype != WM_DTYPE_FP32) return fail(WM_EINVAL, "wm_embed_pos_fwd: dtype %d", dtype);
    const void* ptrs[] = {table, pos_s, pos_h, pos_w, out};
    for (const void* q : ptrs)
        if (!aligned16(q)) return fail(WM_EINVAL, "wm_embed_pos_fwd: pointers must be 16-byte aligned");
    if (dim > 2048) return fail(WM_EUNSUPPORTED, "wm_embed_pos_fwd: dim=%d", dim);
    const long ntok = B * S * H * W;
    const int rows_per_block = 256 / (dim / 8);
    long want = (ntok + rows_per_block - 1) / rows_per_block;
    const long cap = (long)sm_count() * 16;
    const unsigned blocks = (unsigned)(want < cap ? want : cap);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == WM_DTYPE_BF16)
        embed_pos_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(tokens, static_cast<const __nv_bfloat16*>(table),
            static_cast<const __nv_bfloat16*>(pos_s), static_cast<const __nv_bfloat16*>(pos_h),
            static_cast<const __nv_bfloat16*>(pos_w), static_cast<__nv_bfloat16*>(out), ntok, S, H, W, dim, num_rows);
    else
        embed_pos_kernel<float><<<blocks, 256, 0, st>>>(tokens, static_cast<const float*>(table), static_cast<const float*>(pos_s),
            static_cast<const float*>(pos_h), static_cast<const float*>(pos_w), static_cast<float*>(out), ntok, S, H, W, dim,
            num_rows);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}
