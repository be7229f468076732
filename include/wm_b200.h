/*
 * wm_b200.h -- C ABI of the B200-native local-3D-attention / VQ hot path.
 *
 * This is the drop-in boundary.  Every entry point takes plain
 * device pointers, sizes and a CUDA stream handle -- no torch types -- and replaces a
 * piece of the reference that today runs as a chain of ATen ops (or, for QK^T only,
 * as the Triton-1 prototype):
 *
 *   wm_l3d_attn_fwd / wm_l3d_attn_bwd
 *       replace Local3dAttention.local_attention and its autograd graph
 *       (vq-video-diffusion/local_3d_attention.py:78-99, checkpointed at :110-111),
 *       and the prototype launcher blub()
 *       (vq-video-diffusion/triton_prototpye/local_3d_attention_triton1.py:74-104).
 *   wm_vq_nearest / wm_vq_distance
 *       replace the distance + argmin + gather + per-latent error of
 *       VectorQuantizerEMA.forward / .encode / .codebook_distance
 *       (vq-video-diffusion/vq.py:30-36, 77-87).
 *   wm_add_layernorm_fwd / _bwd, wm_bias_gelu_fwd / _bwd, wm_colsum
 *       replace PreNorm, the residual adds and the bias / GELU element-wise chains of the
 *       transformer block around the attention core, with the bias gradients of its
 *       nn.Linear layers (vq-video-diffusion/local_3d_attention.py:11-31, 159-161).
 *   wm_adamw_step
 *       replaces optim.AdamW.step() of the training loop (vq-video-diffusion/main.py:283, 433).
 *
 * Conventions
 *   - All pointers are device pointers owned by the caller (PyTorch's caching
 *     allocator in the Python host); the library allocates no device memory.
 *   - q/k/v/o/dout/dq/dk/dv are the contiguous outputs of nn.Linear:
 *     [B, S, H, W, heads*dim_head], channels split head-major (local_3d_attention.py:85-87).
 *     lse / delta are [B, S, H, W, heads] fp32 (natural-log LSE of the scaled scores).
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Calls
 *     only enqueue work; they never synchronise and are CUDA-graph capturable.
 *   - Return value: 0 on success, a negative WM_E* code otherwise; wm_last_error()
 *     returns a thread-local message.  Nothing throws across the boundary and there
 *     is no CPU fallback.
 */
#ifndef WM_B200_H_
#define WM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define WM_API __attribute__((visibility("default")))
#else
#define WM_API
#endif

#define WM_B200_VERSION 100          /* major*100 + minor */

/* dtype of q/k/v/o and their gradients */
#define WM_DTYPE_BF16 0              /* tensor-core path (tcgen05), fp32 accumulate   */
#define WM_DTYPE_FP32 1              /* exact path, fp32 throughout                   */

/* flags */
#define WM_FLAG_SIMT 1               /* force the SIMT kernels (bf16 attention, VQ): the
                                        on-device cross-check of the tcgen05 kernels   */

/* error codes */
#define WM_OK            0
#define WM_EINVAL      (-1)          /* bad shape / null pointer / misaligned pointer */
#define WM_EUNSUPPORTED (-2)         /* shape outside what the kernels implement      */
#define WM_ECUDA       (-3)          /* a CUDA runtime / driver call failed           */

WM_API int wm_version(void);
WM_API const char* wm_last_error(void);

/* Which kernel family a bf16 call with these shapes would run: 1 = tcgen05/TMA,
 * 0 = SIMT (shape not covered by the tensor-core tiling).  Host-side query, no launch. */
WM_API int wm_l3d_attn_uses_tensor_cores(int S, int H, int W, int heads, int dim_head,
                                  int eS, int eH, int eW, int dtype);

/* out[i] = sum_j softmax_j(scale * q_i . k_j) v_j over the (2e+1)^3 window of token i,
 * neighbours outside the grid excluded (reference: zero pad + masked_fill(-1e9),
 * local_3d_attention.py:82-97).  lse may not be NULL (it is what backward consumes). */
WM_API int wm_l3d_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse,
                    int B, int S, int H, int W, int heads, int dim_head,
                    int eS, int eH, int eW, float scale, int dtype, int flags, void* stream);

/* Gradients of wm_l3d_attn_fwd.  `delta` is a caller-provided fp32 workspace of
 * B*S*H*W*heads elements (row sums of dO*O).  dq/dk/dv are fully overwritten;
 * deterministic (no atomics). */
WM_API int wm_l3d_attn_bwd(const void* q, const void* k, const void* v, const void* o, const float* lse,
                    const void* dout, void* dq, void* dk, void* dv, float* delta,
                    int B, int S, int H, int W, int heads, int dim_head,
                    int eS, int eH, int eW, float scale, int dtype, int flags, void* stream);

/* Nearest codebook entry per latent (vq.py:30-36).
 *   x         [N, L, D] fp32        codebook [L, K, D] fp32
 *   idx       [N, L]    int64       argmin_k sum_d (x_d - e_kd)^2, lowest index on ties
 *   quantized [N, L, D] fp32 or NULL  forward value of the straight-through output,
 *                                     x + (e_idx - x) in fp32 (vq.py:70)
 *   sq_err    [N, L]    fp32 or NULL  sum_d (e_idx - x)^2 (vq.py:35)
 * The winner is decided on distances accumulated in fp64 from the fp32 inputs, so the
 * index equals the exact-arithmetic argmin.  For D in {32,64,96,128} and K <= 512 (multiple
 * of 32) a tf32 tcgen05 distance GEMM selects the candidates that are re-checked exactly;
 * other shapes (or WM_FLAG_SIMT) use the fp32 SIMT filter.  Same indices either way. */
WM_API int wm_vq_nearest(const void* x, const void* codebook, int64_t* idx, void* quantized, float* sq_err,
                  long N, int L, int K, int D, int dtype, int flags, void* stream);

/* distances [N, L, K] fp32 = sum_d (x_d - e_kd)^2, divided by D when normalize != 0
 * (vq.py:77-82). */
WM_API int wm_vq_distance(const void* x, const void* codebook, float* dist,
                   long N, int L, int K, int D, int normalize, void* stream);

/* Fused AdamW over one flat parameter buffer (reference: optim.AdamW, main.py:433, with
 * torch's update rule).  master/exp_avg/exp_avg_sq are fp32 [n]; grad is bf16 or fp32 [n]
 * (grad_dtype) and is multiplied by grad_scale first (1/world_size after a SUM
 * all-reduce); shadow, if not NULL, receives the updated weights rounded to bf16 (the
 * copy the bf16 kernels read).  dyn is a DEVICE array {step (1-based), lr} so that a
 * captured CUDA graph can be replayed while step and learning rate advance. */
WM_API int wm_adamw_step(float* master, void* shadow, const void* grad, float* exp_avg, float* exp_avg_sq,
                         long n, const float* dyn, float beta1, float beta2, float eps, float weight_decay,
                         float grad_scale, int grad_dtype, void* stream);

/* ---- layer-level kernels around the attention core -------------------------------------
 * Replace the ATen chains of PreNorm + the residual adds of the transformer block
 * (local_3d_attention.py:11-17,159-161) and the bias gradients of its nn.Linear layers.
 * Rows are tokens, `dim` channels contiguous (multiple of 8, <= 2048); dtype bf16 or fp32.
 *
 * wm_add_layernorm_fwd: sum = res (+ delta (+ delta_bias[dim]), when != NULL; written to sum_out
 *   when that is != NULL); y = LayerNorm(sum) * gamma + beta; mean / rstd [rows] fp32 kept for
 *   backward.  delta_bias is the bias of the nn.Linear that produced delta (to_out.0 / net.3,
 *   local_3d_attention.py:24-27,52), applied here so that its gradient falls out of the backward.
 * wm_add_layernorm_bwd: dx = (dres, when != NULL, +) dLayerNorm(dy; x, mean, rstd, gamma);
 *   dgamma / dbeta [dim]; dsum [dim] (when != NULL) = column sums of dx = gradient of delta_bias.
 *   workspace: wm_reduce_blocks(rows) * 3 * dim floats.
 * wm_colsum: out[c] = sum_r a[r, c] (bias gradient).  workspace: wm_reduce_blocks(rows) * cols floats. */
WM_API int wm_reduce_blocks(long rows);
WM_API int wm_add_layernorm_fwd(const void* res, const void* delta, const void* delta_bias, const void* gamma,
                                const void* beta, void* sum_out, void* y, float* mean, float* rstd, long rows,
                                int dim, float eps, int dtype, void* stream);
WM_API int wm_add_layernorm_bwd(const void* dy, const void* dres, const void* x, const float* mean,
                                const float* rstd, const void* gamma, void* dx, void* dgamma, void* dbeta,
                                void* dsum, float* workspace, long rows, int dim, int dtype, void* stream);
WM_API int wm_colsum(const void* a, void* out, float* workspace, long rows, int cols, int dtype, void* stream);

/* Bias + exact (erf) GELU around the first MLP GEMM (FeedForward net.0 / net.1, local_3d_attention.py:24-27).
 * h = x W1^T WITHOUT bias, [rows, cols]; cols a multiple of 8, <= 2048.
 * wm_bias_gelu_fwd: y = gelu(h + bias).
 * wm_bias_gelu_bwd: dh = dy * gelu'(h + bias); dbias [cols] = column sums of dh (the gradient of net.0.bias),
 *   reduced in the same pass.  workspace: wm_reduce_blocks(rows) * cols floats. */
WM_API int wm_bias_gelu_fwd(const void* h, const void* bias, void* y, long rows, int cols, int dtype, void* stream);
WM_API int wm_bias_gelu_bwd(const void* dy, const void* h, const void* bias, void* dh, void* dbias, float* workspace,
                            long rows, int cols, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WM_B200_H_ */
