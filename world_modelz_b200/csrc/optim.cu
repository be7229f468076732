// Fused AdamW over a flat parameter buffer (one launch per optimiser step).
// Reference: torch.optim.AdamW as constructed at vq-video-diffusion/main.py:433.
#include "wm_common.cuh"

#include <math.h>

namespace wm {
namespace {

template <typename G>
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ master, __nv_bfloat16* __restrict__ shadow, const G* __restrict__ grad,
             float* __restrict__ m, float* __restrict__ v, long n, const float* __restrict__ dyn, float b1, float b2,
             float eps, float wd, float gscale, float* __restrict__ grad_sq) {
    const float step = dyn[0], lr = dyn[1];
    float sq = 0.f;
    const float bc1 = 1.f - powf(b1, step);
    const float rsqrt_bc2 = rsqrtf(1.f - powf(b2, step));
    const float step_size = lr / bc1;
    const float decay = 1.f - lr * wd;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float g;
        if constexpr (sizeof(G) == 2) g = __bfloat162float(grad[i]) * gscale;
        else g = grad[i] * gscale;
        sq = fmaf(g, g, sq);
        float w = master[i] * decay;
        const float mi = b1 * m[i] + (1.f - b1) * g;
        const float vi = b2 * v[i] + (1.f - b2) * g * g;
        m[i] = mi;
        v[i] = vi;
        w -= step_size * mi / (sqrtf(vi) * rsqrt_bc2 + eps);
        master[i] = w;
        if (shadow != nullptr) shadow[i] = __float2bfloat16_rn(w);
    }
    if (grad_sq != nullptr) {
        // squared gradient norm (main.py:189-193) for free: slot [step & 1] collects this step's sum, the other slot
        // is cleared for the next step (nobody adds to it during this launch)
        __shared__ float red[8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tot = 0.f;
            for (int w = 0; w < 8; ++w) tot += red[w];
            const int slot = ((int)step) & 1;
            atomicAdd(&grad_sq[slot], tot);
            if (blockIdx.x == 0) grad_sq[slot ^ 1] = 0.f;
        }
    }
}

}  // namespace
}  // namespace wm

using namespace wm;

extern "C" int wm_adamw_step_norm(float* master, void* shadow, const void* grad, float* exp_avg, float* exp_avg_sq, long n,
                                  const float* dyn, float beta1, float beta2, float eps, float weight_decay,
                                  float grad_scale, int grad_dtype, float* grad_sq, void* stream);

extern "C" int wm_adamw_step(float* master, void* shadow, const void* grad, float* exp_avg, float* exp_avg_sq, long n,
                             const float* dyn, float beta1, float beta2, float eps, float weight_decay,
                             float grad_scale, int grad_dtype, void* stream) {
    return wm_adamw_step_norm(master, shadow, grad, exp_avg, exp_avg_sq, n, dyn, beta1, beta2, eps, weight_decay, grad_scale,
                              grad_dtype, nullptr, stream);
}

extern "C" int wm_adamw_step_norm(float* master, void* shadow, const void* grad, float* exp_avg, float* exp_avg_sq, long n,
                                  const float* dyn, float beta1, float beta2, float eps, float weight_decay,
                                  float grad_scale, int grad_dtype, float* grad_sq, void* stream) {
    if (n < 0) return fail(WM_EINVAL, "wm_adamw_step: n=%ld", n);
    if (n == 0) return WM_OK;
    if (!master || !grad || !exp_avg || !exp_avg_sq || !dyn) return fail(WM_EINVAL, "wm_adamw_step: null pointer");
    if (grad_dtype != WM_DTYPE_BF16 && grad_dtype != WM_DTYPE_FP32) return fail(WM_EINVAL, "wm_adamw_step: grad dtype %d", grad_dtype);
    long blocks = (n + 255) / 256;
    if (blocks > (long)sm_count() * 16) blocks = (long)sm_count() * 16;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (grad_dtype == WM_DTYPE_BF16)
        adamw_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(master, static_cast<__nv_bfloat16*>(shadow),
                                                                     static_cast<const __nv_bfloat16*>(grad), exp_avg,
                                                                     exp_avg_sq, n, dyn, beta1, beta2, eps,
                                                                     weight_decay, grad_scale, grad_sq);
    else
        adamw_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(master, static_cast<__nv_bfloat16*>(shadow),
                                                             static_cast<const float*>(grad), exp_avg, exp_avg_sq, n,
                                                             dyn, beta1, beta2, eps, weight_decay, grad_scale, grad_sq);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}
