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
 *   wm_adamw_step / wm_adamw_step_norm
 *       replace optim.AdamW.step() of the training loop (vq-video-diffusion/main.py:283, 433) and, with
 *       the squared gradient norm reduced in the same pass, grad_norm() (main.py:189-193).
 *   wm_vq_stats / wm_vq_onehot
 *       replace the one-hot based codebook statistics of VectorQuantizerEMA.forward in training mode
 *       (vq.py:35-46: accumulated_error, embeding_onehot_sum, dw) and the dense float one-hot it returns (vq.py:39).
 *   wm_sample_step
 *       replaces one draw of the mask/replace sampler: top_k_logits + softmax + multinomial + re-masking
 *       (vq-video-diffusion/main.py:39-43, 80-109).
 *   wm_embed_pos_fwd
 *       replaces embedding(img_z) + get_pos_embedding(...) (vq-video-diffusion/local_3d_attention.py:143-157).
 *   wm_loss_hist_update
 *       replaces LossAwareSamplerEma.update_with_losses (vq-video-diffusion/importance_sampling.py:35-41).
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

#define WM_B200_VERSION 120          /* major*100 + minor */

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

/* The same two calls for operands that live side by side in the output of a MERGED projection (to_k and to_v
 * of local_3d_attention.py:106-107 evaluated as one GEMM over the normalised input, N = 2*heads*dim_head): k and v
 * (and dk, dv) are channel slices whose consecutive tokens are ld_kv elements apart, q (and dq) ld_q elements;
 * 0 means heads*dim_head (contiguous).  o, dout, lse and delta are always contiguous.  Strides must be multiples
 * of 16 bytes.  The slices are read by TMA / written by the kernels in place: no split or concatenation copies. */
WM_API int wm_l3d_attn_fwd_ld(const void* q, const void* k, const void* v, void* o, float* lse,
                       long ld_q, long ld_kv, int B, int S, int H, int W, int heads, int dim_head,
                       int eS, int eH, int eW, float scale, int dtype, int flags, void* stream);
WM_API int wm_l3d_attn_bwd_ld(const void* q, const void* k, const void* v, const void* o, const float* lse,
                       const void* dout, void* dq, void* dk, void* dv, float* delta,
                       long ld_q, long ld_kv, int B, int S, int H, int W, int heads, int dim_head,
                       int eS, int eH, int eW, float scale, int dtype, int flags, void* stream);

/* Nearest codebook entry per latent (vq.py:30-36).
 *   x         [N, L, D] fp32        codebook [L, K, D] fp32
 *   idx       [N, L]    int64       argmin_k sum_d (x_d - e_kd)^2, lowest index on ties
 *   quantized [N, L, D] fp32 or NULL  forward value of the straight-through output,
 *                                     x + (e_idx - x) in fp32 (vq.py:70)
 *   sq_err    [N, L]    fp32 or NULL  sum_d (e_idx - x)^2 (vq.py:35)
 * The winner is decided on distances accumulated in fp64 from the fp32 inputs, so the
 * index equals the exact-arithmetic argmin.  For D = 64 and K <= 512 (multiple of 64) a
 * split-bf16 (hi + lo, three products) tcgen05 distance GEMM is the filter and only codes
 * inside its error window are re-checked exactly; other shapes (or WM_FLAG_SIMT) use the
 * fp32 SIMT filter.  Same indices either way. */
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

/* As wm_adamw_step; additionally (grad_sq != NULL) the squared L2 norm of the scaled gradient is reduced in the same
 * pass: grad_sq is a DEVICE array of two floats, both zero before the first step; step s adds its sum to
 * grad_sq[s & 1] and clears the other slot, so after step s the value is grad_sq[s & 1] (main.py:189-193). */
WM_API int wm_adamw_step_norm(float* master, void* shadow, const void* grad, float* exp_avg, float* exp_avg_sq,
                              long n, const float* dyn, float beta1, float beta2, float eps, float weight_decay,
                              float grad_scale, int grad_dtype, float* grad_sq, void* stream);

/* Codebook statistics of one quantizer forward (vq.py:35-46) from the indices, without a one-hot:
 *   counts  [L, K] += number of latents assigned to each code          (embeding_onehot_sum, activation_count)
 *   dw      [L, K, D] += sum of the latents x assigned to each code     (may be NULL; then x may be NULL)
 *   acc_err [L, K] += sum of sq_err over the latents of each code       (accumulated_error; NULL with sq_err NULL)
 * x [N, L, D] fp32, idx [N, L] int64, sq_err [N, L] fp32.  Outputs are ACCUMULATED into (zero them for per-call
 * statistics).  fp32 atomics: each sum is exact up to the order of its terms. */
WM_API int wm_vq_stats(const void* x, const int64_t* idx, const float* sq_err, float* counts, float* dw,
                       float* acc_err, long N, int L, int K, int D, void* stream);

/* encodings [rows, K] fp32 = one_hot(idx[rows]) (vq.py:39), every element written once; K a multiple of 4. */
WM_API int wm_vq_onehot(const int64_t* idx, float* encodings, long rows, int K, void* stream);

/* One draw of the iterative sampler (main.py:80-109) for P = clips * per_clip positions:
 *   logits [P, K] (dtype bf16 / fp32); topk > 0 keeps every logit >= the topk-th largest of its row (main.py:39-43);
 *   sample [P] int64 ~ multinomial(softmax(filtered logits)) (drawn as a Gumbel-max, Philox-4x32-10 keyed by
 *   (seed, call counter, position, code)); frame, when != NULL, receives the draw, or mask_token where an independent
 *   uniform exceeds alpha, at frame[clip * frame_stride + pos] (the last frame of the token tensor).
 *   dyn: DEVICE array {alpha, call counter} so that a captured CUDA graph can be replayed. */
WM_API int wm_sample_step(const void* logits, int64_t* sample, int64_t* frame, long frame_stride, long per_clip,
                          long P, int K, int topk, int mask_token, const float* dyn, uint64_t seed, int dtype,
                          void* stream);

/* Loss-aware diffusion-time histogram (importance_sampling.py:35-41): for i in order,
 * b = clamp(int(ts[i] * buckets)); weights[b] = weights[b] * alpha + losses[i] * (1 - alpha); counts[b] += 1.
 * All device arrays: ts, losses [B] fp32, weights [buckets] fp32, counts [buckets] int64. */
WM_API int wm_loss_hist_update(const float* ts, const float* losses, float* weights, int64_t* counts, int B,
                               int buckets, float alpha, void* stream);

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

/* x[b,s,h,w,:] = table[tokens[b,s,h,w]] + ((pos_s[s] + pos_h[h]) + pos_w[w]): the token embedding plus the three
 * axis position embeddings of Local3dAttentionTransformer.forward (vq-video-diffusion/local_3d_attention.py:149-157)
 * in one pass -- no gathered temporary, no broadcast adds; the sums round to the storage type in the order of the
 * stock ops they replace.  tokens int64 [B,S,H,W] (clamped to [0, num_rows)), table [num_rows, dim], pos_* [>= S|H|W, dim],
 * out [B,S,H,W,dim]; dim a multiple of 8. */
WM_API int wm_embed_pos_fwd(const int64_t* tokens, const void* table, const void* pos_s, const void* pos_h,
                     const void* pos_w, void* out, long B, int S, int H, int W, int dim, int num_rows, int dtype,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WM_B200_H_ */
