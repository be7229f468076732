// VQ nearest-codebook search on the tensor cores: tf32 distance GEMM as a candidate filter,
// exact fp64 re-check of the candidates (bit-exact indices, lowest index on ties).
//
//   d(x, e_k) = |x|^2 + |e_k|^2 - 2 x.e_k        (vq.py:30 computes the direct form in fp32)
//
// A persistent CTA keeps one codebook [K, D] in shared memory (TMA, 128B swizzle, K-major = the
// B operand), streams 128-latent tiles of x through a 2-stage TMA ring (the A operand) and
// computes all 128 x K dot products with tcgen05.mma kind::tf32 into TMEM (K <= 512 columns).
// Epilogue, two threads per latent row:
//   pass 1: t_k = |e_k|^2 - 2 dot_k, row minimum
//   pass 2: every code with t_k <= min + 2*eps is a candidate, eps = 2^-8 |x| max|e| bounding the
//           tf32 rounding of both operands (Cauchy-Schwarz); the candidate set provably contains the
//           exact nearest code; candidates are re-evaluated as sum_d (x_d - e_d)^2 in fp64
//   then indices, the straight-through value x + (e - x) and sum (e - x)^2 (vq.py:34-36,70).
// Replaces VectorQuantizerEMA's [N,L,D,K] distance temporary + argmin + gather (vq.py:30-36,84-87).
#include "tc_common.cuh"
#include "wm_common.cuh"

#include <math.h>

namespace wm {
namespace vq {

using namespace wm::tc;

constexpr int kThreads = 256;
constexpr int kTileM = 128;

struct Params {
    const float* x;            // [N, L, D]
    const float* cb;           // [L, K, D]
    int64_t* idx;              // [N, L]
    float* quantized;          // [N, L, D] or null
    float* sq_err;             // [N, L] or null
    long N;
    int L, K, D;
    int tiles;                 // ceil(N / 128)
    int cb_box;                // codebook rows per TMA box
};

// element (row, channel) of a [rows x 32-float slabs] tile stored with the 128B TMA swizzle
__device__ __forceinline__ const float* sw_elem(const uint8_t* tile, int slab_bytes, int row, int ch) {
    const int slab = ch >> 5, c = ch & 31;
    return reinterpret_cast<const float*>(tile + slab * slab_bytes + row * 128 + ((((c >> 2) ^ (row & 7)) << 4) | ((c & 3) << 2)));
}
__device__ __forceinline__ float4 sw_vec4(const uint8_t* tile, int slab_bytes, int row, int ch4) {   // channels [4*ch4, 4*ch4+4)
    const int slab = ch4 >> 3, c = ch4 & 7;
    return *reinterpret_cast<const float4*>(tile + slab * slab_bytes + row * 128 + ((c ^ (row & 7)) << 4));
}

__host__ __device__ constexpr uint32_t make_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
vq_nearest_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_cb, const Params prm) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int K = prm.K, D = prm.D, L = prm.L;
    const int slabs = D / 32;
    const int cb_slab_bytes = K * 128, x_slab_bytes = kTileM * 128;
    uint8_t* sCB = smem;                                         // [slabs][K rows][128 B]
    uint8_t* sX = sCB + slabs * cb_slab_bytes;                   // [2 stages][slabs][128 rows][128 B]
    float* sNorm = reinterpret_cast<float*>(sX + 2 * slabs * x_slab_bytes);     // [K] |e_k|^2
    float* sXch = sNorm + K;                                     // [2 halves][128] exchange (min / best)
    double* sXd = reinterpret_cast<double*>(sXch + 2 * 128);     // [2 uses][best|second][2 halves][128]
    int* sXi = reinterpret_cast<int*>(sXd + 2 * 512);            // [2 uses][2 halves][128]
    float* sRed = reinterpret_cast<float*>(sXi + 2 * 256);       // [8] block reduction
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + 8);
    uint64_t* bar_cb = bars;
    uint64_t* bar_x = bars + 1;      // [2]
    uint64_t* bar_mma = bars + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane;
    const bool leader = (lane == 0);
    const int l = blockIdx.y;                                    // codebook / latent slot
    const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;

    if (tid == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_cb);
        mbar_init(bar_cb, 1);
        mbar_init(&bar_x[0], 1);
        mbar_init(&bar_x[1], 1);
        mbar_init(bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto issue_x_load = [&](int it, int tile) {                  // warp 0
        if (leader) {
            const int st = it & 1;
            mbar_expect_tx(&bar_x[st], (uint32_t)(slabs * x_slab_bytes));
            for (int sl = 0; sl < slabs; ++sl)
                tma_load_2d(sX + (st * slabs + sl) * x_slab_bytes, &map_x, &bar_x[st], l * D + sl * 32, tile * kTileM);
        }
    };
    const int first_tile = blockIdx.x, stride = gridDim.x;
    if (warp == 0) {
        if (leader) {
            mbar_expect_tx(bar_cb, (uint32_t)(slabs * cb_slab_bytes));
            for (int sl = 0; sl < slabs; ++sl)
                for (int r0 = 0; r0 < K; r0 += prm.cb_box)
                    tma_load_2d(sCB + sl * cb_slab_bytes + r0 * 128, &map_cb, bar_cb, sl * 32, l * K + r0);
        }
        if (first_tile < prm.tiles) issue_x_load(0, first_tile);
        if (first_tile + stride < prm.tiles) issue_x_load(1, first_tile + stride);
    }
    mbar_wait(bar_cb, 0);
    // |e_k|^2 (fp64 accumulate) and max |e_k| of this codebook
    float emax2 = 0.f;
    for (int k = tid; k < K; k += kThreads) {
        double a = 0.0;
        for (int c4 = 0; c4 < D / 4; ++c4) {
            const float4 e = sw_vec4(sCB, cb_slab_bytes, k, c4);
            a += (double)e.x * e.x + (double)e.y * e.y + (double)e.z * e.z + (double)e.w * e.w;
        }
        sNorm[k] = (float)a;
        emax2 = fmaxf(emax2, (float)a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) emax2 = fmaxf(emax2, __shfl_xor_sync(0xffffffffu, emax2, o));
    if (lane == 0) sRed[warp] = emax2;
    __syncthreads();
    emax2 = sRed[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) emax2 = fmaxf(emax2, sRed[w]);
    const float emax = sqrtf(emax2) * 1.0000002f;

    const uint32_t idesc = make_idesc_tf32(K > 256 ? 256 : K);
    const uint32_t idesc_hi = make_idesc_tf32(K > 256 ? K - 256 : 16);
    const uint64_t dx0 = make_smem_desc(smem_u32(sX), 16, 1024, 2u);
    const uint64_t dc0 = make_smem_desc(smem_u32(sCB), 16, 1024, 2u);

    int it = 0;
    for (int tile = first_tile; tile < prm.tiles; tile += stride, ++it) {
        const int st = it & 1;
        const uint8_t* xt = sX + st * slabs * x_slab_bytes;
        // ---- dot products of the tile against the whole codebook -----------------------------------
        if (warp == 0) {
            mbar_wait(&bar_x[st], (it >> 1) & 1);
            tc_fence_after();
            for (int nh = 0; nh * 256 < K; ++nh) {
                const uint32_t id = nh == 0 ? idesc : idesc_hi;
                for (int ks = 0; ks < D / 8; ++ks) {             // 8 tf32 (32 bytes) per MMA
                    const uint32_t sl = (uint32_t)(ks >> 2), ko = (uint32_t)((ks & 3) * 2);
                    const uint64_t da = dx0 + (uint32_t)(((st * slabs + sl) * x_slab_bytes) >> 4) + ko;
                    const uint64_t db = dc0 + (uint32_t)((sl * cb_slab_bytes + nh * 256 * 128) >> 4) + ko;
                    if (leader) umma_tf32_ss(tmem_base + nh * 256, da, db, id, ks > 0);
                }
            }
            if (leader) umma_commit(bar_mma);
        }
        // |x|^2 of this row while the MMAs run (each half takes half of the channels)
        mbar_wait(&bar_x[st], (it >> 1) & 1);
        float xn2 = 0.f;
        for (int c4 = half; c4 < D / 4; c4 += 2) {
            const float4 v = sw_vec4(xt, x_slab_bytes, row, c4);
            xn2 += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        sXch[half * 128 + row] = xn2;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        xn2 += sXch[(half ^ 1) * 128 + row];
        // tf32 keeps 10 mantissa bits of each operand: |dot~ - dot| <= 2^-9 |x||e| (+ fp32 accumulation);
        // eps bounds the error of t_k = |e_k|^2 - 2 dot_k with a 2x safety factor
        const float eps = 0.0078125f * sqrtf(xn2) * emax + 1e-30f;
        mbar_wait(bar_mma, it & 1);
        tc_fence_after();
        const int c_lo = half * (K / 2);                        // this thread's codes: [c_lo, c_lo + K/2), K % 32 == 0
        const int nchunks = K / 32;                             // 16-column chunks per thread (<= 16)
        // pass 1: t_k = |e_k|^2 - 2 dot_k; minimum per 16-column chunk and per row
        float cmin[16];
        float tmin = INFINITY;
#pragma unroll
        for (int ci = 0; ci < 16; ++ci) {
            cmin[ci] = INFINITY;
            if (ci < nchunks) {
                const int c = c_lo + ci * 16;
                uint32_t r[16];
                tmem_ld16(tmem_base + lane_sel + c, r);
                float nk[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 v = *reinterpret_cast<const float4*>(sNorm + c + 4 * i);
                    nk[4 * i] = v.x; nk[4 * i + 1] = v.y; nk[4 * i + 2] = v.z; nk[4 * i + 3] = v.w;
                }
                tmem_wait_ld();
                float m0 = INFINITY, m1 = INFINITY;
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    m0 = fminf(m0, fmaf(-2.f, __uint_as_float(r[i]), nk[i]));
                    m1 = fminf(m1, fmaf(-2.f, __uint_as_float(r[i + 1]), nk[i + 1]));
                }
                cmin[ci] = fminf(m0, m1);
                tmin = fminf(tmin, cmin[ci]);
            }
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");      // xn2 exchange reads are done
        sXch[half * 128 + row] = tmin;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        tmin = fminf(tmin, sXch[(half ^ 1) * 128 + row]);
        const float thr = tmin + 2.f * eps;
        // pass 2: only chunks that hold a candidate (t_k <= thr) are revisited.  Candidates are first
        // ranked by their fp32 direct-form distance; fp64 is needed only for (near-)ties.
        double best = INFINITY, second = INFINITY;
        int best_k = 0x7fffffff;
        // tcgen05.ld is warp-collective: whether a chunk is revisited is decided by a warp vote, and a lane
        // that has nothing to do in it simply ends up with an empty candidate mask
        auto scan = [&](auto zero, bool active) {                 // zero: 0.f -> fp32 arithmetic, 0.0 -> fp64
            using acc_t = decltype(zero);
            if (active) { best = INFINITY; second = INFINITY; best_k = 0x7fffffff; }
#pragma unroll
            for (int ci = 0; ci < 16; ++ci) {
                if (ci < nchunks && __any_sync(0xffffffffu, active && cmin[ci] <= thr)) {
                    const int c = c_lo + ci * 16;
                    uint32_t r[16];
                    tmem_ld16(tmem_base + lane_sel + c, r);
                    tmem_wait_ld();
                    uint32_t cand = 0;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (fmaf(-2.f, __uint_as_float(r[i]), sNorm[c + i]) <= thr) cand |= 1u << i;
                    if (!active) cand = 0;
                    while (cand) {
                        const int i = __ffs(cand) - 1;
                        cand &= cand - 1;
                        const int k = c + i;
                        acc_t a = zero;
                        for (int c4 = 0; c4 < D / 4; ++c4) {
                            const float4 xv = sw_vec4(xt, x_slab_bytes, row, c4);
                            const float4 ev = sw_vec4(sCB, cb_slab_bytes, k, c4);
                            const acc_t d0 = (acc_t)xv.x - (acc_t)ev.x, d1 = (acc_t)xv.y - (acc_t)ev.y;
                            const acc_t d2 = (acc_t)xv.z - (acc_t)ev.z, d3 = (acc_t)xv.w - (acc_t)ev.w;
                            a = fma(d0, d0, a); a = fma(d1, d1, a); a = fma(d2, d2, a); a = fma(d3, d3, a);
                        }
                        const double ad = (double)a;
                        if (ad < best) { second = best; best = ad; best_k = k; }     // ascending codes: first minimum kept
                        else if (ad < second) second = ad;
                    }
                }
            }
        };
        auto merge_halves = [&](int slot_parity) {                 // combine with the other half of the row
            double* xd = sXd + slot_parity * 512;
            int* xi = sXi + slot_parity * 256;
            xd[half * 128 + row] = best;
            xd[256 + half * 128 + row] = second;
            xi[half * 128 + row] = best_k;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
            const double ob = xd[(half ^ 1) * 128 + row], os = xd[256 + (half ^ 1) * 128 + row];
            const int ok = xi[(half ^ 1) * 128 + row];
            if (ob < best || (ob == best && ok < best_k)) { second = fmin(best, os); best = ob; best_k = ok; }
            else second = fmin(second, ob);
        };
        scan(0.f, true);
        merge_halves(0);
        // fp32 direct form is within (D+2) 2^-24 relative of the exact distance: 4x safety band
        const double band = 4.0 * (double)(D + 2) * 5.9604645e-8;
        const bool ambiguous = second <= best * (1.0 + band) + 1e-37;
        if (__any_sync(0xffffffffu, ambiguous)) {                  // both warps of the quadrant see the same rows
            scan(0.0, ambiguous);
            merge_halves(1);
        }
        tc_fence_before();
        __syncthreads();                                           // every TMEM read of this tile is done
        // ---- epilogue: index, straight-through value, squared error ---------------------------------
        const long n = (long)tile * kTileM + row;
        if (n < prm.N) {
            if (half == 0) prm.idx[n * L + l] = (int64_t)best_k;
            float err = 0.f;
            for (int c4 = half; c4 < D / 4; c4 += 2) {
                const float4 xv = sw_vec4(xt, x_slab_bytes, row, c4);
                const float4 ev = sw_vec4(sCB, cb_slab_bytes, best_k, c4);
                const float d0 = ev.x - xv.x, d1 = ev.y - xv.y, d2 = ev.z - xv.z, d3 = ev.w - xv.w;
                err = fmaf(d0, d0, err); err = fmaf(d1, d1, err); err = fmaf(d2, d2, err); err = fmaf(d3, d3, err);
                if (prm.quantized != nullptr)
                    *reinterpret_cast<float4*>(prm.quantized + (n * L + l) * (long)D + c4 * 4) =
                        make_float4(xv.x + d0, xv.y + d1, xv.z + d2, xv.w + d3);
            }
            if (prm.sq_err != nullptr) {
                sXch[half * 128 + row] = err;
            }
        }
        __syncthreads();                                           // x stage fully consumed
        if (prm.sq_err != nullptr && half == 0 && n < prm.N)
            prm.sq_err[n * L + l] = sXch[row] + sXch[128 + row];
        if (warp == 0 && tile + 2 * stride < prm.tiles) issue_x_load(it + 2, tile + 2 * stride);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

static int cb_box_rows(int K) { return K % 256 == 0 ? 256 : K % 128 == 0 ? 128 : K % 64 == 0 ? 64 : 32; }

}  // namespace vq

bool vq_tc_supported(long N, int L, int K, int D) {
    if (D % 32 != 0 || D > 128 || K % 32 != 0 || K > 512 || K < 32) return false;
    const size_t smem = 1024 + (size_t)K * D * 4 + 2ul * 128 * D * 4 + (size_t)K * 4 + 16384;
    return smem <= 227ul * 1024 && N >= 1 && L <= 65535;
}

int vq_nearest_tc(const void* x, const void* cb, int64_t* idx, void* quantized, float* sq_err, long N, int L, int K,
                  int D, cudaStream_t st) {
    using namespace vq;
    CUtensorMap mx, mc;
    // x viewed as [N rows, L*D cols] (row stride L*D floats); the kernel offsets columns by l*D
    const int box = cb_box_rows(K);
    if (int rc = make_tensor_map_2d_f32(&mx, x, (uint64_t)L * D, (uint64_t)N, (uint64_t)L * D * 4, 128)) return rc;
    if (int rc = make_tensor_map_2d_f32(&mc, cb, (uint64_t)D, (uint64_t)L * K, (uint64_t)D * 4, (uint32_t)box)) return rc;
    Params prm{static_cast<const float*>(x), static_cast<const float*>(cb), idx, static_cast<float*>(quantized), sq_err,
               N, L, K, D, (int)((N + kTileM - 1) / kTileM), box};
    const size_t smem = 1024 + (size_t)K * D * 4 + 2ul * 128 * D * 4 + (size_t)K * 4 + 16384;
    WM_CUDA_CHECK(cudaFuncSetAttribute(vq_nearest_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ctas = prm.tiles < 148 ? prm.tiles : 148;
    if (L > 1 && ctas > 148 / L) ctas = 148 / L > 0 ? 148 / L : 1;
    vq_nearest_tc_kernel<<<dim3((unsigned)ctas, (unsigned)L), kThreads, smem, st>>>(mx, mc, prm);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}

}  // namespace wm
