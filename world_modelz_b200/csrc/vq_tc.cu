// VQ nearest-codebook search on the tensor cores: a split-bf16 distance GEMM as the filter, exact fp64 re-check of
// the (rare) near-ties -- bit-exact indices, lowest index on ties.
//
//   d(x, e_k) = |x|^2 + |e_k|^2 - 2 x.e_k            (vq.py:30 computes the direct form in fp32)
//
// Precision of the filter.  x = x_hi + x_lo + r_x with x_hi = bf16(x), x_lo = bf16(x - x_hi), |r_x| <= 2^-17 |x| (same
// for e).  Three bf16 MMAs accumulate  x_hi.e_hi + x_lo.e_hi + x_hi.e_lo  in fp32 (bf16 products are exact in
// fp32), so -2 x.e is off by at most 6 * 2^-18 |x||e| plus the fp32 accumulation of 3D terms (<= 3D * 2^-24 * 2|x||e|);
// eps = 2^-15 |x| max|e| bounds the error of the filtered distance (Cauchy-Schwarz; the split error alone is
// 2.3e-5 |x||e|, the accumulation of the 12 K-steps at D = 64 adds ~1e-6).  Two codes closer than the window 2 eps + (key truncation) are
// re-evaluated exactly; everything else is decided by the filter.
//
// Pipeline (one persistent CTA per SM keeps the whole codebook, pre-scaled by -2, as bf16 hi / lo in shared memory):
//   warps  8-15  loader / writer, 16 lanes per latent row (coalesced): fp32 rows from global memory two tiles
//                ahead (registers), split into the bf16 hi / lo operand tiles (128B-swizzled, K-major) + |x|^2; later,
//                for the finished tile: merge the two code halves, gather the winner, write idx, x + (e - x), sum (e - x)^2
//   warps  0-3   scan code half A (codes [0, K/2), TMEM columns [0, 256)), warps 4-7 half B ([256, 512)); the first warp
//                of each group also issues its half's tcgen05.mma chain for the next tile the moment the group has
//                drained the scores -- while one half is being scanned the tensor pipe works on the other.
//                One thread per row: distance keys (fp32 bits with the code's column
//                index in the low 8 bits) reduced with 3-input integer min / max to the two smallest keys
// Replaces VectorQuantizerEMA's [N,L,D,K] distance temporary + argmin + gather (vq.py:30-36,84-87).
#include "tc_common.cuh"
#include "wm_common.cuh"

#include <math.h>
#include <type_traits>

#ifndef WM_VQ_EXP
#define WM_VQ_EXP 0      // timing experiments: 1 no ambiguity re-scan, 2 no MMAs, 4 no scan, 8 no output, 16 no cross-half check
#endif

namespace wm {
namespace vq {

using namespace wm::tc;

#if WM_VQ_EXP & 32
__device__ unsigned long long g_vq_dbg[4];
#endif

constexpr int kTileM = 128;
constexpr int kThreads = 16 * 32;      // 4 warps per SM sub-partition: 128 registers per thread

struct Params {
    const float* x;            // [N, L, D]
    const float* cb;           // [L, K, D]
    int64_t* idx;              // [N, L]
    float* quantized;          // [N, L, D] or null
    float* sq_err;             // [N, L] or null
    long N;
    int L, K, D;
    int tiles;                 // ceil(N / 128)
};

// byte offset of 16-byte chunk c16 of row r in a [rows x 64 bf16] slab stored K-major with the 128-byte swizzle
__device__ __forceinline__ uint32_t sw128(int r, int c16) { return (uint32_t)r * 128u + (uint32_t)((c16 ^ (r & 7)) << 4); }

__device__ __forceinline__ void split_bf16(float v, uint16_t& hi, uint16_t& lo) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    hi = *reinterpret_cast<const uint16_t*>(&h);
    lo = *reinterpret_cast<const uint16_t*>(&l);
}

// exact squared distance (fp64 accumulation of fp32 differences) between a latent row and a code, both in global memory;
// all loads are issued before the first use (one memory round trip)
template <int D>
__device__ __noinline__ double exact_dist(const float* __restrict__ xr, const float* __restrict__ er) {
    double a = 0.0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {                          // two halves: 16 + 16 loads in flight at D = 64
        float4 xv[D / 8], ev[D / 8];
#pragma unroll
        for (int c = 0; c < D / 8; ++c) {
            xv[c] = __ldg(reinterpret_cast<const float4*>(xr) + h * (D / 8) + c);
            ev[c] = __ldg(reinterpret_cast<const float4*>(er) + h * (D / 8) + c);
        }
#pragma unroll
        for (int c = 0; c < D / 8; ++c) {
            const double d0 = (double)xv[c].x - (double)ev[c].x, d1 = (double)xv[c].y - (double)ev[c].y;
            const double d2 = (double)xv[c].z - (double)ev[c].z, d3 = (double)xv[c].w - (double)ev[c].w;
            a = fma(d0, d0, a); a = fma(d1, d1, a); a = fma(d2, d2, a); a = fma(d3, d3, a);
        }
    }
    return a;
}

// Exact (fp64) winner among the candidates the two code halves reported, by a whole warp: r.y = best code, r.z =
// runner-up (a candidate when r.w == 1), r.w == 2: more than two codes of that half are inside the window -> the whole
// half.  Lanes take candidates round-robin; lowest index wins ties.  Rare path, kept out of line.
template <int D>
__device__ __noinline__ int settle_exact_warp(const float* __restrict__ xr, const float* __restrict__ cbl, uint4 ra, uint4 rb,
                                              bool a_in, bool b_in, int KH, int lane) {
    const int nA = a_in ? (ra.w == 2u ? KH : 1 + (ra.w == 1u)) : 0;
    const int nB = b_in ? (rb.w == 2u ? KH : 1 + (rb.w == 1u)) : 0;
    double bd = INFINITY;
    int bk = 0x7fffffff;
    for (int c = lane; c < nA + nB; c += 32) {
        int k;
        if (c < nA) k = ra.w == 2u ? c : (c == 0 ? (int)ra.y : (int)ra.z);
        else k = rb.w == 2u ? KH + (c - nA) : (c == nA ? (int)rb.y : (int)rb.z);
        const double dd = exact_dist<D>(xr, cbl + (long)k * D);
        if (dd < bd || (dd == bd && k < bk)) { bd = dd; bk = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, bd, o);
        const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
        if (od < bd || (od == bd && ok < bk)) { bd = od; bk = ok; }
    }
    return bk;
}

template <int D>
__global__ void __launch_bounds__(kThreads, 1)
vq_nearest_tc_kernel(const Params prm) {
    constexpr int kSlabs = D / 64;                         // 64-channel (128-byte) slabs
    constexpr int kVec = D / 4;                            // float4 per row
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int K = prm.K, L = prm.L;
    const int KH = K >> 1;                                 // codes per half
    const int e_slab = K * 128, x_slab = kTileM * 128;
    uint8_t* sEhi = smem;                                  // [slabs][K rows][128 B]   -2 * e, bf16 hi
    uint8_t* sElo = sEhi + kSlabs * e_slab;                //                          -2 * e, bf16 lo
    uint8_t* sXhi = sElo + kSlabs * e_slab;                // [2 buffers][slabs][128 rows][128 B]
    uint8_t* sXlo = sXhi + 2 * kSlabs * x_slab;
    float* sNorm = reinterpret_cast<float*>(sXlo + 2 * kSlabs * x_slab);     // [K]  |e_k|^2 + 1
    float* sXn2 = sNorm + K;                               // [2][128] |x|^2
    uint32_t* sRes = reinterpret_cast<uint32_t*>(sXn2 + 2 * kTileM);          // [2][2 halves][128] {t1, code 1, code 2, extra}
    float* sRed = reinterpret_cast<float*>(sRes + 2 * 2 * kTileM * 4);       // [32] block reduction
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + 32);
    uint64_t* bar_xready = bars;         // [2] operand tiles of a latent tile written          (8 loader warps)
    uint64_t* bar_xfree = bars + 2;      // [2] every MMA reading that buffer has retired       (tcgen05.commit of both halves)
    uint64_t* bar_full = bars + 4;       // [2] scores of code half A / B computed              (tcgen05.commit)
    uint64_t* bar_free = bars + 6;       // [2] scores of half A / B drained                    (4 scanner warps)
    uint64_t* bar_res = bars + 8;        // [2] both halves' results of a tile are in sRes      (8 scanner warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int l = blockIdx.y;
    const float* cbl = prm.cb + (long)l * K * D;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_xready[i], 8);
            mbar_init(&bar_xfree[i], 2);
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_free[i], 4);
            mbar_init(&bar_res[i], 8);
        }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);

    // ---- codebook: -2 e as bf16 hi / lo operand tiles, |e|^2 + 1 (fp64 accumulate), max |e| ---------------------------
    float emax2 = 0.f;
    for (int k = tid; k < K; k += kThreads) {
        double a = 0.0;
        const float* er = cbl + (long)k * D;
#pragma unroll 2
        for (int c8 = 0; c8 < D / 8; ++c8) {               // 8 channels = one 16-byte chunk of the bf16 row
            const float4 e0 = __ldg(reinterpret_cast<const float4*>(er + 8 * c8));
            const float4 e1 = __ldg(reinterpret_cast<const float4*>(er + 8 * c8 + 4));
            const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint16_t h0, l0, h1, l1;
                split_bf16(-2.f * ev[2 * i], h0, l0);
                split_bf16(-2.f * ev[2 * i + 1], h1, l1);
                hi[i] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                lo[i] = (uint32_t)l0 | ((uint32_t)l1 << 16);
                a += (double)ev[2 * i] * ev[2 * i] + (double)ev[2 * i + 1] * ev[2 * i + 1];
            }
            const uint32_t off = (uint32_t)(c8 >> 3) * e_slab + sw128(k, c8 & 7);
            *reinterpret_cast<uint4*>(sEhi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(sElo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        sNorm[k] = (float)a + 1.f;        // +1 keeps every filtered distance positive: its fp32 bits then order like integers
        emax2 = fmaxf(emax2, (float)a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) emax2 = fmaxf(emax2, __shfl_xor_sync(0xffffffffu, emax2, o));
    if (lane == 0) sRed[warp] = emax2;
    fence_proxy_async();                  // operand tiles (generic proxy) -> visible to tcgen05.mma
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    emax2 = sRed[0];
#pragma unroll
    for (int w = 1; w < 16; ++w) emax2 = fmaxf(emax2, sRed[w]);
    const float emax = sqrtf(emax2) * 1.0000002f;

    const int first_tile = blockIdx.x, stride = gridDim.x;
    const int my_tiles = first_tile < prm.tiles ? (prm.tiles - first_tile + stride - 1) / stride : 0;

    if (warp < 8) {
        // =============================== scanners (+ MMA issue by the first warp of each half) =====================================================================
        const int quad = warp & 3, half = warp >> 2;
        const int row = quad * 32 + lane;
        const uint32_t taddr = tmem_base + half * 256 + ((uint32_t)(quad * 32) << 16);
        const float* nk = sNorm + half * KH;
        const int nchunk = KH >> 4;
        // ---- the half's MMA chain: x_hi e_hi + x_lo e_hi + x_hi e_lo over all D channels, N = K/2 columns
        const bool issuer = quad == 0;
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc_bf16(KH, false, false);
        const uint64_t dxh = make_smem_desc(smem_u32(sXhi), 16, 1024, 2u), dxl = make_smem_desc(smem_u32(sXlo), 16, 1024, 2u);
        const uint64_t deh = make_smem_desc(smem_u32(sEhi), 16, 1024, 2u), del = make_smem_desc(smem_u32(sElo), 16, 1024, 2u);
        auto issue_tile = [&](int j) {           // called by the issuing warp once its half is free (or was never used)
            const int b = j & 1;
            mbar_wait(&bar_xready[b], (j >> 1) & 1);
            if (j > 0) mbar_wait(&bar_free[half], (j - 1) & 1);           // all four warps of the group have drained tile j-1
            tc_fence_after();
            if (leader) {
                const uint32_t xoff = (uint32_t)((b * kSlabs * x_slab) >> 4);
                const uint32_t eoff = (uint32_t)((half * KH * 128) >> 4);
                bool first = true;
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const uint64_t da = (p == 1 ? dxl : dxh) + xoff;
                    const uint64_t db = (p == 2 ? del : deh) + eoff;
#pragma unroll
                    for (int kk = 0; kk < ((WM_VQ_EXP & 2) ? 0 : D / 16); ++kk) {
                        const uint32_t sl = (uint32_t)(kk >> 2), ko = (uint32_t)((kk & 3) * 2);
                        umma_bf16_ss(tmem_base + half * 256, da + sl * (uint32_t)(x_slab >> 4) + ko,
                                     db + sl * (uint32_t)(e_slab >> 4) + ko, idesc, first ? 0u : 1u);
                        first = false;
                    }
                }
                umma_commit(&bar_full[half]);
                umma_commit(&bar_xfree[b]);                               // (both halves' commits free operand buffer b)
            }
            __syncwarp();
        };
        if (issuer && my_tiles > 0) issue_tile(0);
        for (int j = 0; j < my_tiles; ++j) {
            const int b = j & 1;
            const long n = ((long)(first_tile + j * stride)) * kTileM + row;
            mbar_wait(&bar_full[half], j & 1);
            tc_fence_after();
            const float xn2 = sXn2[b * kTileM + row];
            const uint64_t xx = pk2(xn2, xn2);
            // two smallest keys of this half and the 16-column chunks they came from.  A key is the filtered distance's
            // fp32 bits (positive, so they order like integers) with the column's position inside its chunk in the low 4
            // bits: the comparison network sees distances truncated by 2^-19 relative, nothing more
            uint32_t m1 = 0x7f800000u, m2 = 0x7f800000u;
            int c1 = 0, c2 = 0;
            uint32_t r0[16], r1[16];
            auto scan16 = [&](const uint32_t (&r)[16], int c) {
                uint32_t a1 = 0x7f800000u, a2 = 0x7f800000u;              // the chunk's two smallest
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 nn = *reinterpret_cast<const float2*>(nk + c * 16 + 2 * i);
                    const uint64_t t2 = fadd2(fadd2(pk2u(r[2 * i], r[2 * i + 1]), xx), pk2(nn.x, nn.y));
                    float ta, tb;
                    upk2(t2, ta, tb);
                    const uint32_t ka = (__float_as_uint(ta) & 0xfffffff0u) | (uint32_t)(2 * i);
                    const uint32_t kb = (__float_as_uint(tb) & 0xfffffff0u) | (uint32_t)(2 * i + 1);
                    const uint32_t lo = min(ka, kb), hi = max(ka, kb);
                    a2 = (uint32_t)__vimin3_s32((int)a2, (int)hi, (int)max(a1, lo));
                    a1 = min(a1, lo);
                }
                // merge (a1, a2) of chunk c into (m1, m2): the second smallest overall is min(m2, a2, max(m1, a1))
                const bool first = a1 < m1;
                const uint32_t loser = first ? m1 : a1;                  // max(m1, a1)
                const int loser_c = first ? c1 : c;
                const uint32_t rest = min(m2, a2);
                const int rest_c = (a2 < m2) ? c : c2;
                const bool l2 = loser < rest;
                m2 = l2 ? loser : rest;
                c2 = l2 ? loser_c : rest_c;
                if (first) { m1 = a1; c1 = c; }
            };
            tmem_ld16(taddr, r0);
            for (int c = 0; c < ((WM_VQ_EXP & 4) ? 0 : nchunk); c += 2) {
                tmem_wait_ld();
                tmem_regs_ready(r0);
                if (c + 1 < nchunk) tmem_ld16(taddr + (c + 1) * 16, r1);
                scan16(r0, c);
                if (c + 1 < nchunk) {
                    tmem_wait_ld();
                    tmem_regs_ready(r1);
                    if (c + 2 < nchunk) tmem_ld16(taddr + (c + 2) * 16, r0);
                    scan16(r1, c + 1);
                }
            }
            const float t1 = __uint_as_float(m1 & 0xfffffff0u), t2nd = __uint_as_float(m2 & 0xfffffff0u);
            const int k1 = half * KH + c1 * 16 + (int)(m1 & 15u), k2 = half * KH + c2 * 16 + (int)(m2 & 15u);
            // window: rounding of the filter (2 eps) + the 4 key bits dropped from each of the two distances
            const float win = 6.103515625e-5f * sqrtf(xn2) * emax + 3.9e-6f * t2nd + 1e-30f;
            const bool close2 = !(WM_VQ_EXP & 1) && (n < prm.N) && (t2nd - t1 <= win);
            uint32_t extra = 0;                                           // 1: the runner-up is a candidate too; 2: more than two are
#if WM_VQ_EXP & 32
            { const unsigned m = __ballot_sync(0xffffffffu, close2); if (lane == 0) { atomicAdd(&g_vq_dbg[2], (unsigned long long)__popc(m)); atomicAdd(&g_vq_dbg[3], m ? 1ull : 0ull); } }
#endif
            if (__any_sync(0xffffffffu, close2) && !((WM_VQ_EXP & 128) && (prm.N >> 40) == 0)) {
                // rare (about one row in a thousand): count the codes of this half inside the window while the scores are
                // still ours.  Two: the writer settles k1 against k2 exactly.  More: it re-scans the half exactly.
                const float thr = t1 + win;
                int cnt = 0;
                for (int c = 0; c < nchunk; ++c) {
                    uint32_t r[16];
                    tmem_ld16(taddr + c * 16, r);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) cnt += ((__uint_as_float(r[i]) + xn2) + nk[c * 16 + i]) <= thr;
                }
                if (close2) extra = cnt > 2 ? 2u : 1u;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_free[half]);                    // this half may be refilled
            if (issuer && j + 1 < my_tiles) issue_tile(j + 1);
            uint4* res = reinterpret_cast<uint4*>(sRes) + (b * 2 + half) * kTileM + row;
            *res = make_uint4(__float_as_uint(t1), (uint32_t)k1, (uint32_t)k2, extra);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_res[b]);
        }
    } else {
        // =============================== loaders / writers (warps 8-15) ===============================================
        // Coalesced mapping: a warp owns 16 rows; one 16-byte access per lane covers TWO whole rows (16 lanes x 16 B = one
        // 256-byte fp32 row), so instruction i touches rows 2i, 2i+1 of the warp and a lane holds fp32 chunk `ch` (4
        // channels) of 8 rows.  (One thread per row would put every lane on its own 128-byte line: 16x the L1 wavefronts.)
        constexpr int kRowsPerLane = 8;
        static_assert(D == 64, "row = 16 lanes x float4");
        const int wrow0 = (warp - 8) * 16, sub = lane >> 4, ch = lane & 15;
        float4 xv[kRowsPerLane];                           // tile being loaded (two tiles ahead of its output)
        auto tile_base = [&](int j) { return ((long)(first_tile + j * stride)) * kTileM; };
        auto load_rows = [&](int j, float4 (&dst)[kRowsPerLane]) {
            const long n0 = tile_base(j) + wrow0 + sub;
#pragma unroll
            for (int i = 0; i < kRowsPerLane; ++i) {
                const long n = n0 + 2 * i;
                dst[i] = n < prm.N ? __ldg(reinterpret_cast<const float4*>(prm.x + (n * L + l) * (long)D) + ch)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto convert_rows = [&](int j) {                   // registers -> bf16 hi / lo operand tiles of buffer j & 1, |x|^2
            const int b = j & 1;
            if (j >= 2) mbar_wait(&bar_xfree[b], ((j - 2) >> 1) & 1);       // the MMAs of tile j-2 have read this buffer
#pragma unroll
            for (int i = 0; i < kRowsPerLane; ++i) {
                const int row = wrow0 + 2 * i + sub;
                const float v[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
                uint16_t h[4], lo[4];
                float xn2 = 0.f;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    split_bf16(v[e], h[e], lo[e]);
                    xn2 = fmaf(v[e], v[e], xn2);
                }
                // fp32 chunk ch (4 channels) = half `ch & 1` of the 16-byte bf16 chunk ch >> 1
                const uint32_t off = (uint32_t)(b * kSlabs * x_slab) + sw128(row, ch >> 1) + (uint32_t)(ch & 1) * 8u;
                *reinterpret_cast<uint2*>(sXhi + off) = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
                *reinterpret_cast<uint2*>(sXlo + off) = make_uint2((uint32_t)lo[0] | ((uint32_t)lo[1] << 16), (uint32_t)lo[2] | ((uint32_t)lo[3] << 16));
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) xn2 += __shfl_xor_sync(0xffffffffu, xn2, o);     // over the row's 16 lanes
                if (ch == 0) sXn2[b * kTileM + row] = xn2;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_xready[b]);
        };
        auto output_rows = [&](int j) {                    // merge the halves, gather the winner, write the outputs
            const int b = j & 1;
            float4 xo[kRowsPerLane];
            load_rows(j, xo);                              // re-read (L2 hit), in flight while waiting for the scanners
            mbar_wait(&bar_res[b], (j >> 1) & 1);
            const bool want_q = !(WM_VQ_EXP & 8) && (prm.quantized != nullptr || prm.sq_err != nullptr);
            uint32_t redo = 0;                             // rows of this warp (bit = row - wrow0) whose winner needs the exact path
#pragma unroll
            for (int i = 0; i < kRowsPerLane; ++i) {
                const int row = wrow0 + 2 * i + sub;
                const long n = tile_base(j) + row;
                const bool valid = n < prm.N;
                const uint4 ra = reinterpret_cast<const uint4*>(sRes)[(b * 2 + 0) * kTileM + row];
                const uint4 rb = reinterpret_cast<const uint4*>(sRes)[(b * 2 + 1) * kTileM + row];
                const float ta = __uint_as_float(ra.x), tb = __uint_as_float(rb.x);
                const float xn2 = sXn2[b * kTileM + row];
                int best = ta <= tb ? (int)ra.y : (int)rb.y;
                const float win = 6.103515625e-5f * sqrtf(xn2) * emax + 3.9e-6f * fmaxf(ta, tb) + 1e-30f;
                const bool a_in = ta <= tb + win, b_in = tb <= ta + win;   // does the half hold a candidate for the row minimum?
                const bool need = valid && !(WM_VQ_EXP & 16) && ((a_in && b_in) || (a_in && ra.w) || (b_in && rb.w));
                redo |= need ? (1u << (2 * i + sub)) : 0u;  // settled exactly after the loop (rare); the row is rewritten then
                if (valid && ch == 0) prm.idx[n * L + l] = (int64_t)best;
                if (want_q) {
                    const float4 ev = __ldg(reinterpret_cast<const float4*>(cbl + (long)(valid ? best : 0) * D) + ch);
                    const float d0 = ev.x - xo[i].x, d1 = ev.y - xo[i].y, d2 = ev.z - xo[i].z, d3 = ev.w - xo[i].w;
                    float err = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3)));
                    if (prm.quantized != nullptr && valid)
                        reinterpret_cast<float4*>(prm.quantized + (n * L + l) * (long)D)[ch] =
                            make_float4(xo[i].x + d0, xo[i].y + d1, xo[i].z + d2, xo[i].w + d3);
                    if (prm.sq_err != nullptr) {
#pragma unroll
                        for (int o = 8; o > 0; o >>= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
                        if (ch == 0 && valid) prm.sq_err[n * L + l] = err;
                    }
                }
            }
            // ---- rare: rows whose candidates lie inside the filter's window are settled exactly by the whole warp and rewritten
            redo = __reduce_or_sync(0xffffffffu, redo);
#if WM_VQ_EXP & 64
            redo &= (uint32_t)(prm.N >> 40);               // timing experiment: never run the exact path (all code kept)
#endif
#if WM_VQ_EXP & 32
            if (lane == 0) { atomicAdd(&g_vq_dbg[0], (unsigned long long)__popc(redo)); atomicAdd(&g_vq_dbg[1], 16ull); }
#endif
            while (redo) {
                const int rr = __ffs(redo) - 1;
                redo &= redo - 1;
                const int row = wrow0 + rr;
                const long n = tile_base(j) + row;
                const uint4 ra = reinterpret_cast<const uint4*>(sRes)[(b * 2 + 0) * kTileM + row];
                const uint4 rb = reinterpret_cast<const uint4*>(sRes)[(b * 2 + 1) * kTileM + row];
                const float ta = __uint_as_float(ra.x), tb = __uint_as_float(rb.x);
                const float win = 6.103515625e-5f * sqrtf(sXn2[b * kTileM + row]) * emax + 3.9e-6f * fmaxf(ta, tb) + 1e-30f;
                const float* xr = prm.x + (n * L + l) * (long)D;
                const int best = settle_exact_warp<D>(xr, cbl, ra, rb, ta <= tb + win, tb <= ta + win, KH, lane);
                if (lane == 0) prm.idx[n * L + l] = (int64_t)best;
                if (want_q && lane < 16) {
                    const float4 xq = __ldg(reinterpret_cast<const float4*>(xr) + lane);
                    const float4 ev = __ldg(reinterpret_cast<const float4*>(cbl + (long)best * D) + lane);
                    const float d0 = ev.x - xq.x, d1 = ev.y - xq.y, d2 = ev.z - xq.z, d3 = ev.w - xq.w;
                    float err = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3)));
                    if (prm.quantized != nullptr)
                        reinterpret_cast<float4*>(prm.quantized + (n * L + l) * (long)D)[lane] = make_float4(xq.x + d0, xq.y + d1, xq.z + d2, xq.w + d3);
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) err += __shfl_xor_sync(0x0000ffffu, err, o);
                    if (prm.sq_err != nullptr && lane == 0) prm.sq_err[n * L + l] = err;
                }
                __syncwarp();
            }
        };
        // software pipeline: load two tiles ahead of the output
        if (my_tiles > 0) { load_rows(0, xv); convert_rows(0); }
        if (my_tiles > 1) { load_rows(1, xv); convert_rows(1); }
        for (int j = 0; j < my_tiles; ++j) {
            if (j + 2 < my_tiles) load_rows(j + 2, xv);    // global loads in flight across the output of tile j
            output_rows(j);
            if (j + 2 < my_tiles) convert_rows(j + 2);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

static size_t smem_bytes(int K, int D) {
    const int slabs = D / 64;
    return 1024 + 2ul * slabs * K * 128 + 4ul * slabs * kTileM * 128 + (size_t)K * 4 + 2 * kTileM * 4 + 2 * 2 * kTileM * 16 + 128 + 256;
}

}  // namespace vq

#if WM_VQ_EXP & 32
extern "C" __attribute__((visibility("default"))) int wm_vq_debug_read(unsigned long long* out) {
    return (int)cudaMemcpyFromSymbol(out, vq::g_vq_dbg, sizeof(unsigned long long) * 4);
}
#endif

bool vq_tc_supported(long N, int L, int K, int D) {
    if (D != 64) return false;                             // one 64-channel (128-byte) operand slab; the codebook's hi / lo
                                                           // tiles of a 128-channel code would not fit next to the latents
    if (K % 32 != 0 || K > 512 || K < 32) return false;    // two halves of <= 256 codes, multiples of 16
    return vq::smem_bytes(K, D) <= 227ul * 1024 && N >= 1 && L <= 65535;
}

int vq_nearest_tc(const void* x, const void* cb, int64_t* idx, void* quantized, float* sq_err, long N, int L, int K,
                  int D, cudaStream_t st) {
    using namespace vq;
    Params prm{static_cast<const float*>(x), static_cast<const float*>(cb), idx, static_cast<float*>(quantized), sq_err,
               N, L, K, D, (int)((N + kTileM - 1) / kTileM)};
    const size_t smem = smem_bytes(K, D);
    const int sms = sm_count();
    int ctas = prm.tiles < sms ? prm.tiles : sms;
    if (L > 1 && ctas > sms / L) ctas = sms / L > 0 ? sms / L : 1;
    const dim3 grid((unsigned)ctas, (unsigned)L);
    WM_CUDA_CHECK(cudaFuncSetAttribute(vq_nearest_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vq_nearest_tc_kernel<64><<<grid, kThreads, smem, st>>>(prm);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}

}  // namespace wm
