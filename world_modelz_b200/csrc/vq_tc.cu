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
// Pipeline (one persistent CTA per SM keeps the whole codebook, pre-scaled by -2, as bf16 hi / lo in shared memory; 17 warps):
//   warps  8-15  loader / writer, 16 lanes per latent row (coalesced).  A tile's fp32 rows arrive by ONE cp.async.bulk three
//                tiles ahead and are split in place into the bf16 hi / lo operand tiles (128B-swizzled, K-major) + |x|^2; later,
//                for the finished tile: merge the two code halves and decide (one row per lane), write idx, gather the
//                winners, write x + (e - x) and sum (e - x)^2.  These warps take 112 registers (setmaxnreg): a tile's
//                re-read rows + gathered codes are 64 of them, and at 96 the row loop reloaded spills at ~300 cycles each
//   warp   16    MMA issuer: per tile four tcgen05.mma chains, one per code QUARTER (N = K/4 columns of tensor memory each),
//                each issued as soon as the scanners have handed that quarter back.  The issuing thread blocks while the
//                tensor pipe's queue is full (3 300 cycles per tile when a scanner warp did this between two scans), and
//                sleeps between barrier probes (spinning probes were a fifth of all issued instructions)
//   warps  0-3   scan code quarters 0, 1 (codes [0, K/2)), warps 4-7 quarters 2, 3; 80 registers.  A quarter goes back to the
//                issuer the moment its last tcgen05.ld has returned, so its next chain runs while the group scans the other
//                quarter: the scanners never wait for the tensor pipe.
//                One thread per row: distance keys (the score's fp32 bits times 16, the code's position inside its
//                16-column chunk below them: one IMAD per score -- with the factor passed as a kernel parameter, or ptxas
//                turns it into an LEA on the ALU pipe that the min / max instructions already saturate).  Smallest key by a
//                3-input min tree per chunk; the runner-up exactly from two partitions of the keys (chunk minima and the
//                minima of the 16 position classes: 1.3 ALU instructions per score instead of 3.25 for a running top-2).  |x|^2 + |e_k|^2 + 2 arrive WITH the score: a
//                thirteenth K = 16 MMA step multiplies [1 1 1 | split3(|x|^2 + 2)] with [split3(|e_k|^2) | 1 1 1] (three bf16
//                terms carry an fp32 value exactly), so the scanners add nothing and read no shared memory.  A runner-up
//                inside the error window marks the whole half for the exact settlement (the scores are gone by then)
// Timeline of the hand-offs: tools/vq_timeline.py (library built with -DWM_VQ_EXP=32).
// Replaces VectorQuantizerEMA's [N,L,D,K] distance temporary + argmin + gather (vq.py:30-36,84-87).
#include "tc_common.cuh"
#include "wm_common.cuh"

#include <math.h>
#include <type_traits>

#ifndef WM_VQ_EXP
#define WM_VQ_EXP 0      // timing experiments: 1 no ambiguity flag, 2 no MMAs, 4 no scan, 8 no output, 16 no cross-half check, 32 timeline
#endif

namespace wm {
namespace vq {

using namespace wm::tc;

#if WM_VQ_EXP & 32
// timeline of CTA (0, 0): clock64 at the hand-offs of the first 64 tiles (tools/vq_timeline.py)
__device__ unsigned long long g_vq_dbg[64 * 24];
#define VQ_TL(tile, ev)                                                                                   \
    do {                                                                                                  \
        if (blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && (tile) < 64)                 \
            g_vq_dbg[(tile) * 24 + (ev)] = (unsigned long long)clock64();                                 \
    } while (0)
#else
#define VQ_TL(tile, ev) do { } while (0)
#endif

#ifndef WM_VQ_TUNE
#define WM_VQ_TUNE 19         // bit 0: the issuer warp sleeps between barrier probes; bit 1: keys by IMAD (FMA pipe) instead of LEA (ALU pipe);
                              // bit 4: runner-up from two partitions of the keys (chunks and position classes) instead of a running top-2;
                              // bit 2: the gather of a tile's winners is in flight across the conversion of the tile two ahead (measured: 0.285 vs
                              // 0.279 ms -- the re-read of x then waits behind the conversion); bit 3: with bit 2, re-read x first (spills)
#endif
#ifndef WM_VQ_REGS
#define WM_VQ_REGS 1          // 1: the scanner warps hand 16 registers per thread to the writer warps (setmaxnreg: 80 / 112 instead of 96 / 96)
#endif
#ifndef WM_VQ_ISSUER_WARP
#define WM_VQ_ISSUER_WARP 1   // 1: a seventeenth warp issues the MMA chains; 0: the first scanner warp of each group does (16 warps, 128 registers)
#endif

constexpr int kTileM = 128;
constexpr int kThreads = (16 + WM_VQ_ISSUER_WARP) * 32;      // 8 scanner, 8 loader / writer warps, 1 MMA issuer: 120 registers per thread

struct Params {
    const float* x;            // [N, L, D]
    const float* cb;           // [L, K, D]
    int64_t* idx;              // [N, L]
    float* quantized;          // [N, L, D] or null
    float* sq_err;             // [N, L] or null
    long N;
    int L, K, D;
    int tiles;                 // ceil(N / 128)
    uint32_t key_mul;          // 16, as a kernel parameter: a multiplier the compiler cannot see keeps the key an IMAD (FMA pipe)
};

// byte offset of 16-byte chunk c16 of row r in a [rows x 64 bf16] slab stored K-major with the 128-byte swizzle
__device__ __forceinline__ uint32_t sw128(int r, int c16) { return (uint32_t)r * 128u + (uint32_t)((c16 ^ (r & 7)) << 4); }

// 1-D bulk copy global -> shared (TMA engine, no tensor map), completion counted on an mbarrier
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr uint32_t kKeyBase = 0x08000000u;      // -(bits(1.0f) << 4) mod 2^32: key = (bits(t) - bits(1.0f)) * 16 + position
constexpr float kHugeNorm = 1.0e9f;             // |x|^2 + max|e|^2 below this keeps every score under 2^32 (28 key bits of exponent + mantissa)

// A wait that gives its issue slots away: the MMA issuer spends most of its life waiting, on a sub-partition it shares with two
// scanner and two writer warps (spinning probes were a fifth of all instructions the kernel issued).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
#if WM_VQ_TUNE & 1
    while (!mbar_test(bar, parity)) __nanosleep(32);
#else
    mbar_wait(bar, parity);
#endif
}

__device__ __forceinline__ uint32_t vmin3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("min.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));      // ptxas fuses the pair into one 3-input VIMNMX3
    asm("min.u32 %0, %1, %2;" : "=r"(d) : "r"(d), "r"(c));
    return d;
}

// Two codes whose FILTERED scores differ by more than this are ordered like their exact distances.  The filter's score is
// |x|^2 + |e|^2 + 2 - 2 x.e with x.e from three bf16 products (split error <= 6 * 2^-18 |x||e| ~ 2.3e-5 |x||e|) and every
// term accumulated in fp32 by the tensor core: 13 K-steps, each adding a rounding of at most 2^-23 of the running sum,
// which is bounded by (|x| + |e|)^2 + 2 <= 2 (|x|^2 + |e|^2) + 2.  Twice the per-score bound, with slack:
__device__ __forceinline__ float filter_window(float xn2, float emax) {
    float r;                                              // one MUFU; its relative error (2^-22) is inside the 2^-10 slack of the factor
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(xn2));
    return 6.11e-5f * r * emax + 8.0e-6f * (xn2 + emax * emax + 1.f) + 1e-30f;
}

// hi / lo split of two values with the packed conversion (F2FP on the ALU; the scalar cvt is a quarter-rate XU instruction)
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// Sum of p[i] over the 16 lanes that share lane bit 4, for 8 values at once: 8 shuffles instead of 32.  The butterfly
// exchanges half of the remaining values per stage, so the additions form the same tree as four xor-shuffles per value;
// afterwards the lanes ch = 2 i and 2 i + 1 (ch = lane & 15) hold the total of value i.
__device__ __forceinline__ float reduce8_over16(const float (&p)[8], int ch) {
    float q[4], r[2];
    const bool b8 = ch & 8, b4 = ch & 4, b2 = ch & 2;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float mine = b8 ? p[k + 4] : p[k], other = b8 ? p[k] : p[k + 4];
        q[k] = mine + __shfl_xor_sync(0xffffffffu, other, 8);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float mine = b4 ? q[k + 2] : q[k], other = b4 ? q[k] : q[k + 2];
        r[k] = mine + __shfl_xor_sync(0xffffffffu, other, 4);
    }
    const float mine = b2 ? r[1] : r[0], other = b2 ? r[0] : r[1];
    float t = mine + __shfl_xor_sync(0xffffffffu, other, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    return t;                                             // value index ((ch >> 1) & 1) + 2 ((ch >> 2) & 1) + 4 ((ch >> 3) & 1) = ch >> 1
}

__device__ __forceinline__ void split_bf16(float v, uint16_t& hi, uint16_t& lo) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    hi = *reinterpret_cast<const uint16_t*>(&h);
    lo = *reinterpret_cast<const uint16_t*>(&l);
}

// v = h + m + l exactly (three bf16 terms hold the 24 significant bits of an fp32 value)
__device__ __forceinline__ void split3_bf16(float v, uint16_t (&out)[3]) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    const __nv_bfloat16 l = __float2bfloat16_rn(r1 - __bfloat162float(m));
    out[0] = *reinterpret_cast<const uint16_t*>(&h);
    out[1] = *reinterpret_cast<const uint16_t*>(&m);
    out[2] = *reinterpret_cast<const uint16_t*>(&l);
}

// exact squared distance (fp64 accumulation of fp32 differences) between a latent row (staged in shared memory: every lane of
// the warp reads the same values, a broadcast) and a code in global memory; all 16 loads of the code are issued before the
// first use -- ONE memory round trip per candidate
template <int D>
__device__ __forceinline__ double exact_dist(const float* __restrict__ xs, const float* __restrict__ er) {
    float4 ev[D / 4];
#pragma unroll
    for (int c = 0; c < D / 4; ++c) ev[c] = __ldg(reinterpret_cast<const float4*>(er) + c);
    double a = 0.0;
#pragma unroll
    for (int c = 0; c < D / 4; ++c) {
        const float4 xv = reinterpret_cast<const float4*>(xs)[c];
        const double d0 = (double)xv.x - (double)ev[c].x, d1 = (double)xv.y - (double)ev[c].y;
        const double d2 = (double)xv.z - (double)ev[c].z, d3 = (double)xv.w - (double)ev[c].w;
        a = fma(d0, d0, a); a = fma(d1, d1, a); a = fma(d2, d2, a); a = fma(d3, d3, a);
    }
    return a;
}

// Undecided rows (two or more codes inside the filter's error window -- about one row in a thousand on Gaussian data)
// leave the filter kernel with a NEGATIVE idx that packs their candidates: per code half the best code, the runner-up
// and whether the runner-up (w = 1) or more than two codes (w = 2: the whole half) lie inside the window.
__device__ __forceinline__ long pack_candidates(uint32_t k1a, uint32_t k2a, uint32_t wa, uint32_t k1b, uint32_t k2b, uint32_t wb,
                                                bool a_in, bool b_in) {
    const uint64_t m = (uint64_t)(k1a & 511u) | ((uint64_t)(k2a & 511u) << 9) | ((uint64_t)(wa & 3u) << 18) |
                       ((uint64_t)(k1b & 511u) << 20) | ((uint64_t)(k2b & 511u) << 29) | ((uint64_t)(wb & 3u) << 38) |
                       ((uint64_t)(a_in ? 1u : 0u) << 40) | ((uint64_t)(b_in ? 1u : 0u) << 41) | (1ull << 63);
    return (long)m;
}

// Second kernel, launched right behind the filter: undecided rows are settled exactly (fp64 distances to their candidates,
// lowest index on ties) and idx / quantized / sq_err rewritten.  Keeping this out of the filter kernel matters: with the
// fp64 path inlined there, its mere presence (registers, code size) cost the filter 17 %.
// One wave of blocks; a block scans its slice of idx in rounds of kSettleRound items (coalesced), queues the marked rows in
// shared memory and then spreads them over its warps -- with a warp per 32 consecutive rows the few warps that met a marked
// row set the time of each of seven waves (38 us for ~2 000 rows; 16 us of it just to get every warp scheduled once).
constexpr int kSettleThreads = 256;
constexpr int kSettleRound = kSettleThreads * 8;

template <int D>
__device__ __forceinline__ void settle_row(const Params& prm, long item, int lane, float* xs) {
    const uint64_t m = (uint64_t)prm.idx[item];
    const int l = (int)(item % prm.L);
    const float* xr = prm.x + item * (long)D;
    __syncwarp();                                         // the previous row's distances are done with xs
    if (lane < D / 4) reinterpret_cast<float4*>(xs)[lane] = __ldg(reinterpret_cast<const float4*>(xr) + lane);
    __syncwarp();
    const float* cbl = prm.cb + (long)l * prm.K * D;
    // candidates: per code half its best code, its runner-up (w == 1) or the whole half (w == 2)
    const int KH = prm.K >> 1;
    const uint32_t k1a = m & 511u, k2a = (m >> 9) & 511u, wa = (m >> 18) & 3u;
    const uint32_t k1b = (m >> 20) & 511u, k2b = (m >> 29) & 511u, wb = (m >> 38) & 3u;
    const int nA = ((m >> 40) & 1u) ? (wa == 2u ? KH : 1 + (wa == 1u)) : 0;
    const int nB = ((m >> 41) & 1u) ? (wb == 2u ? KH : 1 + (wb == 1u)) : 0;
    double bd = INFINITY;
    int bk = 0x7fffffff;
    for (int c = lane; c < nA + nB; c += 32) {
        int k;
        if (c < nA) k = wa == 2u ? c : (c == 0 ? (int)k1a : (int)k2a);
        else k = wb == 2u ? KH + (c - nA) : (c == nA ? (int)k1b : (int)k2b);
        const double dd = exact_dist<D>(xs, cbl + (long)k * D);
        if (dd < bd || (dd == bd && k < bk)) { bd = dd; bk = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, bd, o);
        const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
        if (od < bd || (od == bd && ok < bk)) { bd = od; bk = ok; }
    }
    __syncwarp();                                         // every lane has read the marker before it is overwritten
    if (lane == 0) prm.idx[item] = (int64_t)bk;
    if ((prm.quantized != nullptr || prm.sq_err != nullptr) && lane < D / 4) {
        const float4 xq = __ldg(reinterpret_cast<const float4*>(xr) + lane);
        const float4 ev = __ldg(reinterpret_cast<const float4*>(cbl + (long)bk * D) + lane);
        const float d0 = ev.x - xq.x, d1 = ev.y - xq.y, d2 = ev.z - xq.z, d3 = ev.w - xq.w;
        float err = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3)));
        if (prm.quantized != nullptr)
            reinterpret_cast<float4*>(prm.quantized + item * (long)D)[lane] = make_float4(xq.x + d0, xq.y + d1, xq.z + d2, xq.w + d3);
#pragma unroll
        for (int o = D / 8; o > 0; o >>= 1) err += __shfl_xor_sync((1u << (D / 4)) - 1u, err, o);
        if (prm.sq_err != nullptr && lane == 0) prm.sq_err[item] = err;
    }
    __syncwarp();
}

template <int D>
__global__ void __launch_bounds__(kSettleThreads)
vq_settle_kernel(const Params prm) {
    __shared__ unsigned short queue[kSettleRound];
    __shared__ unsigned queued;
    __shared__ __align__(16) float xrow[kSettleThreads / 32][D];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long items = prm.N * prm.L;
    const long rounds = (items + kSettleRound - 1) / kSettleRound;
    for (long r = blockIdx.x; r < rounds; r += gridDim.x) {
        const long base = r * kSettleRound;
        if (tid == 0) queued = 0;
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kSettleRound / kSettleThreads; ++u) {
            const int off = u * kSettleThreads + tid;
            if (base + off < items && prm.idx[base + off] < 0) queue[atomicAdd(&queued, 1u)] = (unsigned short)off;
        }
        __syncthreads();
        const unsigned n = queued;
        for (unsigned e = warp; e < n; e += kSettleThreads / 32) settle_row<D>(prm, base + queue[e], lane, xrow[warp]);
        __syncthreads();
    }
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
    const int e_slab = K * 128;
    uint8_t* sEhi = smem;                                  // [slabs][K rows][128 B]   -2 * e, bf16 hi
    uint8_t* sElo = sEhi + kSlabs * e_slab;                //                          -2 * e, bf16 lo
    // Latent tiles: two 32 KB buffers.  A tile arrives as fp32 rows (cp.async.bulk, 256 B per row) and is converted IN
    // PLACE into the bf16 hi / lo operand tiles: the 8-row swizzle atoms of hi and lo alternate (hi atom g at g * 2048,
    // lo atom g at g * 2048 + 1024; descriptors with SBO = 2048), so the 16 rows a loader warp owns occupy the same
    // 4 KB as fp32 data and as operands -- the warp reads them all, then overwrites them.
    uint8_t* sX = sElo + kSlabs * e_slab;                  // [2 buffers][32 KB]
    constexpr int kXBuf = kTileM * D * 4;
    static_assert(kXBuf == 2 * kTileM * 128, "fp32 tile and the hi + lo operand tiles must have the same size");
    // operands of the norm step (K = 16 channels, no swizzle: 8-row x 16-byte core matrices, the two 8-channel columns
    // LBO apart): x side [1 1 1 h m l 0 0 | 0..], h + m + l = |x|^2 + 2;  e side [h m l 1 1 1 0 0 | 0..], h + m + l = |e_k|^2
    uint8_t* sXe = sX + 2 * kXBuf;                         // [2 buffers][2 columns][128 rows][16 B]
    uint8_t* sEe = sXe + 2 * 2 * kTileM * 16;              // [2 columns][K rows][16 B]
    float* sXn2 = reinterpret_cast<float*>(sEe + 2 * K * 16);               // [3][128] |x|^2 (tile j in slot j % 3: tile j+2 is
                                                                              // converted while tiles j and j+1 still need theirs)
    uint32_t* sRes = reinterpret_cast<uint32_t*>(sXn2 + 3 * kTileM);          // [2][2 halves][128] {t1, code 1, code 2, extra}
    float* sRed = reinterpret_cast<float*>(sRes + 2 * 2 * kTileM * 4);       // [32] block reduction
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + 32);
    uint64_t* bar_xready = bars;         // [2] operand tiles of a latent tile written          (8 loader warps)
    uint64_t* bar_xfree = bars + 2;      // [2] every MMA reading that buffer has retired       (tcgen05.commit behind the tile's last chain)
    uint64_t* bar_full = bars + 4;       // [4] scores of a code quarter computed               (tcgen05.commit)
    uint64_t* bar_free = bars + 8;       // [4] scores of a code quarter drained                (4 scanner warps)
    uint64_t* bar_res = bars + 12;       // [2] both halves' results of a tile are in sRes      (8 scanner warps)
    uint64_t* bar_xload = bars + 14;     // [2] fp32 rows of a latent tile landed                 (cp.async.bulk complete_tx)
    uint64_t* bar_resfree = bars + 16;   // [2] the writers have consumed a tile's sRes entries     (8 writer warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int l = blockIdx.y;
    const float* cbl = prm.cb + (long)l * K * D;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_xready[i], 8);
            mbar_init(&bar_xfree[i], WM_VQ_ISSUER_WARP ? 1 : 2);
            mbar_init(&bar_xload[i], 1);
            mbar_init(&bar_resfree[i], 8);
            mbar_init(&bar_res[i], 8);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_free[i], 4);
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
        {   // e side of the norm step
            const float nrm = (float)a;
            uint16_t h3[3];
            split3_bf16(nrm, h3);
            const uint32_t off = (uint32_t)(k >> 3) * 128u + (uint32_t)(k & 7) * 16u;
            *reinterpret_cast<uint4*>(sEe + off) = make_uint4((uint32_t)h3[0] | ((uint32_t)h3[1] << 16), (uint32_t)h3[2] | (0x3F80u << 16),
                                                              0x3F803F80u, 0u);
            *reinterpret_cast<uint4*>(sEe + K * 16 + off) = make_uint4(0u, 0u, 0u, 0u);
        }
        emax2 = fmaxf(emax2, (float)a);
    }
    for (int i = tid; i < 2 * kTileM; i += kThreads)
        *reinterpret_cast<uint4*>(sXe + (i / kTileM) * (2 * kTileM * 16) + kTileM * 16 + (i % kTileM) * 16) = make_uint4(0u, 0u, 0u, 0u);
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
    for (int w = 1; w < kThreads / 32; ++w) emax2 = fmaxf(emax2, sRed[w]);
    const float emax = sqrtf(emax2) * 1.0000002f;

    const int first_tile = blockIdx.x, stride = gridDim.x;
    const int my_tiles = first_tile < prm.tiles ? (prm.tiles - first_tile + stride - 1) / stride : 0;

    const int KQ = KH >> 1;                                // codes per quarter = N of one MMA chain (tensor-memory columns [q * 128, q * 128 + KQ))
    // One MMA chain: scores of code quarter q for the latent tile in operand buffer b -- x_hi e_hi + x_lo e_hi + x_hi e_lo over all D
    // channels, then the norm step (+ |x|^2 + 2 + |e_k|^2).  Called by one elected thread.
    auto issue_chain = [&](int b, int q, bool last_of_tile) {
        const uint32_t idesc = make_idesc_bf16(KQ, false, false);
        const uint64_t dxh = make_smem_desc(smem_u32(sX), 16, 2048, 2u), dxl = make_smem_desc(smem_u32(sX + 1024), 16, 2048, 2u);
        const uint64_t deh = make_smem_desc(smem_u32(sEhi), 16, 1024, 2u), del = make_smem_desc(smem_u32(sElo), 16, 1024, 2u);
        const uint64_t dxe = make_smem_desc(smem_u32(sXe), (uint32_t)kTileM * 16u, 128u, 0u);      // norm step: no swizzle
        const uint64_t dee = make_smem_desc(smem_u32(sEe), (uint32_t)K * 16u, 128u, 0u);
        const uint32_t xoff = (uint32_t)((b * kXBuf) >> 4);
        const uint32_t eoff = (uint32_t)((q * KQ * 128) >> 4);
        const uint32_t dst = tmem_base + (uint32_t)q * 128u;
        bool first = true;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const uint64_t da = (p == 1 ? dxl : dxh) + xoff;
            const uint64_t db = (p == 2 ? del : deh) + eoff;
#pragma unroll
            for (int kk = 0; kk < ((WM_VQ_EXP & 2) ? 0 : D / 16); ++kk) {
                const uint32_t sl = (uint32_t)(kk >> 2), ko = (uint32_t)((kk & 3) * 2);
                umma_bf16_ss(dst, da + ko, db + sl * (uint32_t)(e_slab >> 4) + ko, idesc, first ? 0u : 1u);
                first = false;
            }
        }
        if (!(WM_VQ_EXP & 2))
            umma_bf16_ss(dst, dxe + (uint32_t)((b * 2 * kTileM * 16) >> 4), dee + (uint32_t)((q * KQ * 16) >> 4), idesc, 1u);
        umma_commit(&bar_full[q]);
        if (last_of_tile) umma_commit(&bar_xfree[b]);      // every MMA (of this issuer) reading operand buffer b has been issued
    };
#if WM_VQ_ISSUER_WARP
    if (warp == 16) {
        // =============================== MMA issuer (its own warp: the issuing thread blocks while the tensor pipe's queue is full -- 3 300
        // cycles per tile when a scanner warp did this between two scans).  Per tile four chains, one per code quarter, in the
        // order 0, 2, 1, 3: each scanner group gets its first quarter first. ===========================================================
        const bool leader = elect_one();
        for (int j = 0; j < my_tiles; ++j) {
            const int b = j & 1;
            mbar_wait_relaxed(&bar_xready[b], (j >> 1) & 1);
            VQ_TL(j, 9);
#pragma unroll 1
            for (int qi = 0; qi < 4; ++qi) {
                const int q = ((qi & 1) << 1) | (qi >> 1);
                if (j > 0) mbar_wait_relaxed(&bar_free[q], (j - 1) & 1);  // the quarter's four scanner warps have drained tile j-1
                tc_fence_after();
                if (leader) issue_chain(b, q, qi == 3);
                __syncwarp();
            }
            VQ_TL(j, 10);
        }
    } else
#endif
    if (warp < 8) {
        // =============================== scanners ===============================================================================================
        // Group A (warps 0-3) scans code quarters 0 and 1, group B (warps 4-7) quarters 2 and 3, one thread per row.  A quarter is handed
        // back to the issuer the moment its last tcgen05.ld has returned: its next tile's chain runs while the group scans the other quarter.
#if WM_VQ_REGS && WM_VQ_ISSUER_WARP
        asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
#endif
        const int quad = warp & 3, half = warp >> 2;
        const int row = quad * 32 + lane;
        const int nq = KQ >> 4;                            // 16-column chunks per quarter
#if !WM_VQ_ISSUER_WARP
        const bool leader = elect_one();
        if (quad == 0 && my_tiles > 0) {                   // the group's first warp issues its two quarters' chains
            mbar_wait(&bar_xready[0], 0);
            tc_fence_after();
            if (leader) { issue_chain(0, half * 2, false); issue_chain(0, half * 2 + 1, true); }
            __syncwarp();
        }
#endif
        for (int j = 0; j < my_tiles; ++j) {
            const int b = j & 1;
            const long n = ((long)(first_tile + j * stride)) * kTileM + row;
            // Smallest key of this half (with the 16-column chunk it came from) and the second smallest.  A score t = |x - e|^2 + 2
            // (>= 1, so its fp32 bits order like integers) becomes the key (bits(t) - bits(1)) * 16 + position-in-chunk: one
            // IMAD, no bits dropped; t < 2^32 is the writer's business (kHugeNorm).
            // The runner-up costs ONE min per key instead of the 2.5 min / max of a running top-2: the keys are covered by two
            // partitions -- the chunks (min by a 3-input tree, 8 instructions per 16 keys, then a top-2 of the chunk minima) and the
            // 16 position classes (b[i] = smallest key seen at position i of any chunk).  Winner w and runner-up r differ in chunk
            // or in position; every chunk / class minimum other than w is >= r, and r is the minimum of its chunk (if the chunks
            // differ) or of its class (if the positions do).  So min(second chunk minimum, second class minimum) = r exactly.
#if WM_VQ_TUNE & 16
            uint32_t m1 = 0xffffffffu, m2 = 0xffffffffu;
            int c1 = 0;
            uint32_t cls[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) cls[i] = 0xffffffffu;
            uint32_t r0[16], r1[16];
            const uint32_t sixteen = prm.key_mul;          // opaque to the compiler: the key stays an IMAD (FMA pipe), not an LEA (ALU pipe)
            auto scan16 = [&](const uint32_t (&r)[16], int c) {
                uint32_t k[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    k[i] = r[i] * sixteen + (kKeyBase + (uint32_t)i);
                    cls[i] = min(cls[i], k[i]);
                }
                const uint32_t t0 = vmin3(k[0], k[1], k[2]), t1 = vmin3(k[3], k[4], k[5]), t2 = vmin3(k[6], k[7], k[8]);
                const uint32_t t3 = vmin3(k[9], k[10], k[11]), t4 = vmin3(k[12], k[13], k[14]);
                const uint32_t a1 = min(vmin3(t0, t1, t2), vmin3(t3, t4, k[15]));
                m2 = min(m2, max(m1, a1));                 // second smallest chunk minimum (== m1 for an equal key in another chunk)
                if (a1 < m1) c1 = c;
                m1 = min(m1, a1);
            };
#else
            uint32_t m1 = 0xffffffffu, m2 = 0xffffffffu;
            int c1 = 0, c2 = 0;
            uint32_t r0[16], r1[16];
#if WM_VQ_TUNE & 2
            const uint32_t sixteen = prm.key_mul;          // opaque to the compiler: the key becomes an IMAD (FMA pipe), not an LEA on the
                                                           // ALU pipe, which the 3-input min / max instructions already saturate
#else
            const uint32_t sixteen = 16u;
#endif
            auto scan16 = [&](const uint32_t (&r)[16], int c) {
                uint32_t a1 = 0xffffffffu, a2 = 0xffffffffu;              // the chunk's two smallest
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t ka = r[2 * i] * sixteen + (kKeyBase + (uint32_t)(2 * i));
                    const uint32_t kb = r[2 * i + 1] * sixteen + (kKeyBase + (uint32_t)(2 * i + 1));
                    const uint32_t lo = min(ka, kb), hi = max(ka, kb);
                    a2 = vmin3(a2, hi, max(a1, lo));
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
#endif
#pragma unroll 1
            for (int qq = 0; qq < 2; ++qq) {
                const int q = half * 2 + qq;
                const uint32_t taddr = tmem_base + (uint32_t)q * 128u + ((uint32_t)(quad * 32) << 16);
                const int cbase = qq * nq;
                mbar_wait(&bar_full[q], j & 1);
                tc_fence_after();
                if (quad == 0 && qq == 0) VQ_TL(j, half ? 12 : 7);
                auto release = [&]() {                     // every score of the quarter is in registers
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_free[q]);
#if !WM_VQ_ISSUER_WARP
                    if (quad == 0 && j + 1 < my_tiles) {   // next tile's chain for this quarter, before this warp scans its last chunk
                        mbar_wait(&bar_xready[(j + 1) & 1], ((j + 1) >> 1) & 1);
                        mbar_wait(&bar_free[q], j & 1);
                        tc_fence_after();
                        if (leader) issue_chain((j + 1) & 1, q, qq == 1);
                        __syncwarp();
                    }
#endif
                };
                if (!(WM_VQ_EXP & 4)) {
                    tmem_ld16(taddr, r0);
                    for (int c = 0; c < nq; c += 2) {
                        tmem_wait_ld();
                        tmem_regs_ready(r0);
                        if (c + 1 < nq) tmem_ld16(taddr + (c + 1) * 16, r1);
                        else release();
                        scan16(r0, cbase + c);
                        if (c + 1 < nq) {
                            tmem_wait_ld();
                            tmem_regs_ready(r1);
                            if (c + 2 < nq) tmem_ld16(taddr + (c + 2) * 16, r0);
                            else release();
                            scan16(r1, cbase + c + 1);
                        }
                    }
                } else {
                    release();
                }
            }
            if (quad == 0) VQ_TL(j, half ? 13 : 8);
            const float xn2 = sXn2[(j % 3) * kTileM + row];
#if WM_VQ_TUNE & 16
            {   // second smallest class minimum: the winner's own class wraps to 2^32 - 1 under  - (m1 + 1)
                const uint32_t sub = m1 + 1u;
                uint32_t d[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) d[i] = cls[i] - sub;
                const uint32_t t0 = vmin3(d[0], d[1], d[2]), t1b = vmin3(d[3], d[4], d[5]), t2 = vmin3(d[6], d[7], d[8]);
                const uint32_t t3 = vmin3(d[9], d[10], d[11]), t4 = vmin3(d[12], d[13], d[14]);
                m2 = min(m2, min(vmin3(t0, t1b, t2), vmin3(t3, t4, d[15])) + sub);
            }
            const int c2 = c1;
#endif
            const float t1 = __uint_as_float((m1 >> 4) + 0x3f800000u), t2nd = __uint_as_float((m2 >> 4) + 0x3f800000u);
            const int k1 = half * KH + c1 * 16 + (int)(m1 & 15u), k2 = half * KH + c2 * 16 + (int)(m2 & 15u);
            const float win = filter_window(xn2, emax);
            // the runner-up inside the window (about one row in a thousand): the scores are gone by now (the quarters were handed back), so
            // the whole half becomes the candidate set of the exact settlement
            const bool close2 = !(WM_VQ_EXP & 1) && (n < prm.N) && (t2nd - t1 <= win);
            const uint32_t extra = close2 ? 2u : 0u;
            if (j >= 2) mbar_wait(&bar_resfree[b], ((j - 2) >> 1) & 1);    // tile j-2's results have been written out
            uint4* res = reinterpret_cast<uint4*>(sRes) + (b * 2 + half) * kTileM + row;
            *res = make_uint4(__float_as_uint(t1), (uint32_t)k1, (uint32_t)k2, extra);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_res[b]);
            if (quad == 0 && half == 0) VQ_TL(j, 11);
        }
    } else {
        // =============================== loaders / writers (warps 8-15) ===============================================
        // Coalesced mapping: a warp owns 16 rows; one 16-byte access per lane covers TWO whole rows (16 lanes x 16 B = one
        // 256-byte fp32 row), so instruction i touches rows 2i, 2i+1 of the warp and a lane holds fp32 chunk `ch` (4
        // channels) of 8 rows.  (One thread per row would put every lane on its own 128-byte line: 16x the L1 wavefronts.)
#if WM_VQ_REGS && WM_VQ_ISSUER_WARP
        asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");     // xo + ev of a tile are 64 registers; at 96 the row loop reloaded spills
#endif
        constexpr int kRowsPerLane = 8;
        static_assert(D == 64, "row = 16 lanes x float4");
        const int wrow0 = (warp - 8) * 16, sub = lane >> 4, ch = lane & 15;
        auto tile_base = [&](int j) { return ((long)(first_tile + j * stride)) * kTileM; };
        // Rows are addressed as (tile pointer) + 32-bit offset and bounds are a 32-bit row count: eight 64-bit row indices per lane
        // spilled to local memory, and every reload in the output loop was a ~300-cycle stall (the L1 is all but carved away).
        const int LD = L * D;                              // floats between consecutive latent rows of this codebook slot
        auto rows_left = [&](int j) { const long left = prm.N - tile_base(j); return left < kTileM ? (int)left : kTileM; };
        auto load_rows = [&](int j, float4 (&dst)[kRowsPerLane]) {      // output path: re-read of the tile (L2 hit)
            const float* xt = prm.x + (tile_base(j) * L + l) * (long)D + ch * 4;
            const int nv = rows_left(j);
#pragma unroll
            for (int i = 0; i < kRowsPerLane; ++i) {
                const int r = wrow0 + sub + 2 * i;
                dst[i] = r < nv ? __ldg(reinterpret_cast<const float4*>(xt + r * LD)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        // fp32 rows of tile j -> buffer j & 1, one 256-byte bulk copy per row (rows are L * D floats apart), all on one
        // mbarrier: 32 KB in flight per SM without a register or an instruction slot spent on it
        const int ltid = tid - 8 * 32;
        auto issue_tile_load = [&](int j) {
            const int b = j & 1;
            if (j >= 2) mbar_wait(&bar_xfree[b], ((j - 2) >> 1) & 1);       // the MMAs of tile j-2 have read this buffer
            if (warp == 8) VQ_TL(j, 3);
            const long base = tile_base(j);
            const long left = prm.N - base;
            const int valid = left < kTileM ? (int)left : kTileM;
            if (ltid == 0) mbar_expect_tx(&bar_xload[b], (uint32_t)valid * (uint32_t)(D * 4));
            if (L == 1) {                                  // rows back to back: ONE copy (a bulk copy issues per thread, ~60 cycles each)
                if (ltid == 0) bulk_load(sX + b * kXBuf, prm.x + base * (long)D, (uint32_t)valid * (uint32_t)(D * 4), &bar_xload[b]);
            } else if (ltid < valid) {
                bulk_load(sX + b * kXBuf + ltid * (D * 4), prm.x + ((base + ltid) * L + l) * (long)D, D * 4, &bar_xload[b]);
            }
        };
        auto convert_rows = [&](int j) {                   // fp32 rows (shared) -> bf16 hi / lo operand tiles in place, |x|^2
            const int b = j & 1;
            uint8_t* buf = sX + b * kXBuf;
            if (warp == 8) VQ_TL(j, 0);
            mbar_wait(&bar_xload[b], (j >> 1) & 1);
            if (warp == 8) VQ_TL(j, 1);
            float4 xv[kRowsPerLane];
#pragma unroll
            for (int i = 0; i < kRowsPerLane; ++i) {
                const int row = wrow0 + 2 * i + sub;
                xv[i] = (tile_base(j) + row < prm.N) ? *reinterpret_cast<const float4*>(buf + row * (D * 4) + ch * 16)
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncwarp();                                  // every fp32 value of this warp's 16 rows is in registers
            float part[kRowsPerLane];
#pragma unroll
            for (int i = 0; i < kRowsPerLane; ++i) {
                const int row = wrow0 + 2 * i + sub;
                uint32_t h01, l01, h23, l23;
                split_bf16x2(xv[i].x, xv[i].y, h01, l01);
                split_bf16x2(xv[i].z, xv[i].w, h23, l23);
                part[i] = fmaf(xv[i].w, xv[i].w, fmaf(xv[i].z, xv[i].z, fmaf(xv[i].y, xv[i].y, xv[i].x * xv[i].x)));
                // fp32 chunk ch (4 channels) = half `ch & 1` of the 16-byte bf16 chunk ch >> 1; hi atom, lo atom 1 KB behind it
                const uint32_t off = (uint32_t)(row >> 3) * 2048u + sw128(row & 7, ch >> 1) + (uint32_t)(ch & 1) * 8u;
                *reinterpret_cast<uint2*>(buf + off) = make_uint2(h01, h23);
                *reinterpret_cast<uint2*>(buf + off + 1024) = make_uint2(l01, l23);
            }
            {   // |x|^2 of the 8 rows in 8 shuffles; lanes ch = 2 i write row i's norm operands
                const float xn2 = reduce8_over16(part, ch);
                if ((ch & 1) == 0) {
                    const int row = wrow0 + 2 * (ch >> 1) + sub;
                    sXn2[(j % 3) * kTileM + row] = xn2;
                    uint16_t h3[3];
                    split3_bf16(xn2 + 2.f, h3);     // + 2: every score stays >= 1 under the filter's error
                    *reinterpret_cast<uint4*>(sXe + b * (2 * kTileM * 16) + (row >> 3) * 128 + (row & 7) * 16) =
                        make_uint4(0x3F803F80u, 0x3F80u | ((uint32_t)h3[0] << 16), (uint32_t)h3[1] | ((uint32_t)h3[2] << 16), 0u);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_xready[b]);
            if (warp == 8) VQ_TL(j, 2);
        };
        const bool want_q = !(WM_VQ_EXP & 8) && (prm.quantized != nullptr || prm.sq_err != nullptr);
        // Output, first part: wait for the scanners, merge the halves, decide, write idx, start the gather of the winners.
        // ONE row per lane here (lane r and r + 16 both take row r of the warp's 16; all 16 lanes of a row doing it for their
        // 8 rows was 16-fold redundant work), the winners then go to the lanes that own the rows by shuffle.
        auto output_decide = [&](int j, float4 (&ev)[kRowsPerLane]) {
            const int b = j & 1;
            if (warp == 8) VQ_TL(j, 4);
            mbar_wait(&bar_res[b], (j >> 1) & 1);
            if (warp == 8) VQ_TL(j, 5);
            int bk;                                        // winner, or -1: undecided (its packed candidates go to idx)
            {
                const int row = wrow0 + ch;
                const bool valid = row < rows_left(j);
                const uint4 ra = reinterpret_cast<const uint4*>(sRes)[(b * 2 + 0) * kTileM + row];
                const uint4 rb = reinterpret_cast<const uint4*>(sRes)[(b * 2 + 1) * kTileM + row];
                const float ta = __uint_as_float(ra.x), tb = __uint_as_float(rb.x);
                const float xn2 = sXn2[(j % 3) * kTileM + row];
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_resfree[b]);   // this warp's sRes entries of tile j may be overwritten
                bk = valid ? (ta <= tb ? (int)ra.y : (int)rb.y) : 0;
                const float win = filter_window(xn2, emax);
                const bool huge = !(xn2 + emax * emax < kHugeNorm);       // scores may leave the key range (or are not finite): settle exactly
                const bool a_in = ta <= tb + win, b_in = tb <= ta + win;   // does the half hold a candidate for the row minimum?
                const bool need = valid && !(WM_VQ_EXP & 16) && (huge || (a_in && b_in) || (a_in && ra.w) || (b_in && rb.w));
                long out = (long)bk;
                if (need) {                                 // undecided: settled exactly by vq_settle_kernel, which rewrites the row
                    out = pack_candidates(ra.y, ra.z, huge ? 2u : ra.w, rb.y, rb.z, huge ? 2u : rb.w, huge || a_in, huge || b_in);
                    bk = -1;
                }
                if (valid && sub == 0) prm.idx[tile_base(j) * L + l + row * L] = (int64_t)out;
            }
            if (want_q) {
#pragma unroll
                for (int i = 0; i < kRowsPerLane; ++i) {
                    const int best = __shfl_sync(0xffffffffu, bk, 2 * i + sub);
                    ev[i] = __ldg(reinterpret_cast<const float4*>(cbl + (best < 0 ? 0 : best) * D) + ch);
                }
            }
            if (warp == 8) VQ_TL(j, 15);
        };
        // Output, second part: x + (e - x), sum (e - x)^2
        auto output_write = [&](int j, const float4 (&ev)[kRowsPerLane], const float4 (&xo)[kRowsPerLane]) {
            const int nv = rows_left(j);
            const long o0 = tile_base(j) * L + l;          // output slot of the tile's first row
            float* qt = prm.quantized != nullptr ? prm.quantized + o0 * (long)D + ch * 4 : nullptr;
            float errp[kRowsPerLane];
#pragma unroll
            for (int i = 0; i < kRowsPerLane; ++i) {
                const int r = wrow0 + sub + 2 * i;
                const bool valid = r < nv;
                errp[i] = 0.f;
                if (want_q) {
                    const float d0 = ev[i].x - xo[i].x, d1 = ev[i].y - xo[i].y, d2 = ev[i].z - xo[i].z, d3 = ev[i].w - xo[i].w;
                    errp[i] = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3)));
                    if (qt != nullptr && valid)
                        *reinterpret_cast<float4*>(qt + r * LD) = make_float4(xo[i].x + d0, xo[i].y + d1, xo[i].z + d2, xo[i].w + d3);
                }
                if (i == 0 && warp == 8) VQ_TL(j, 14);
            }
            if (warp == 8) VQ_TL(j, 16);
            if (want_q && prm.sq_err != nullptr) {        // the 8 rows' errors in 8 shuffles; lanes ch = 2 i hold row i's
                const float err = reduce8_over16(errp, ch);
                const int r = wrow0 + sub + 2 * (ch >> 1);
                if ((ch & 1) == 0 && r < nv) prm.sq_err[o0 + r * L] = err;
            }
            if (warp == 8) VQ_TL(j, 6);
        };
        // Software pipeline.  Tile j's results arrive, its winners are decided and the gather of their codes is started; the
        // conversion of tile j+2 (its fp32 rows landed during the previous iteration) and the copy of tile j+3 (its buffer is
        // free once the MMAs of tile j+1 have retired) run while those loads are in flight -- the gathered rows come back one
        // every ~400 cycles per warp -- and then tile j is written out.  The conversion waits for the SCAN of tile j only, never
        // for its output (that would close a loop scan -> output -> convert -> MMA -> scan over two tiles).
        if (my_tiles > 0) issue_tile_load(0);
        if (my_tiles > 1) issue_tile_load(1);
        if (my_tiles > 0) convert_rows(0);
        if (my_tiles > 1) convert_rows(1);
        if (my_tiles > 2) issue_tile_load(2);
        for (int j = 0; j < my_tiles; ++j) {
            float4 ev[kRowsPerLane], xo[kRowsPerLane];
#if WM_VQ_TUNE & 4
#if WM_VQ_TUNE & 8
            load_rows(j, xo);                              // re-read of tile j (L2 hit)
#endif
            output_decide(j, ev);
            if (j + 2 < my_tiles) convert_rows(j + 2);
            if (j + 3 < my_tiles) issue_tile_load(j + 3);
#if !(WM_VQ_TUNE & 8)
            load_rows(j, xo);                              // re-read of tile j (L2 hit)
#endif
#else
            if (j + 2 < my_tiles) convert_rows(j + 2);
            if (j + 3 < my_tiles) issue_tile_load(j + 3);
            load_rows(j, xo);                              // re-read of tile j (L2 hit)
            output_decide(j, ev);
#endif
            output_write(j, ev, xo);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

static size_t smem_bytes(int K, int D) {
    const int slabs = D / 64;
    return 1024 + 2ul * slabs * K * 128 + 4ul * slabs * kTileM * 128 + 2ul * 2 * kTileM * 16 + 2ul * K * 16 + 3 * kTileM * 4 +
           2 * 2 * kTileM * 16 + 128 + 256;
}

}  // namespace vq

#if WM_VQ_EXP & 32
extern "C" __attribute__((visibility("default"))) int wm_vq_debug_read(unsigned long long* out) {
    return (int)cudaMemcpyFromSymbol(out, vq::g_vq_dbg, sizeof(unsigned long long) * 64 * 24);
}
#endif

bool vq_tc_supported(long N, int L, int K, int D) {
    if (D != 64) return false;                             // one 64-channel (128-byte) operand slab; the codebook's hi / lo
                                                           // tiles of a 128-channel code would not fit next to the latents
    if (K % 64 != 0 || K > 512 || K < 64) return false;    // four quarters of <= 128 codes, multiples of 16
    return vq::smem_bytes(K, D) <= 227ul * 1024 && N >= 1 && L <= 65535;
}

int vq_nearest_tc(const void* x, const void* cb, int64_t* idx, void* quantized, float* sq_err, long N, int L, int K,
                  int D, cudaStream_t st) {
    using namespace vq;
    Params prm{static_cast<const float*>(x), static_cast<const float*>(cb), idx, static_cast<float*>(quantized), sq_err,
               N, L, K, D, (int)((N + kTileM - 1) / kTileM), 16u};
    const size_t smem = smem_bytes(K, D);
    const int sms = sm_count();
    int ctas = prm.tiles < sms ? prm.tiles : sms;
    if (L > 1 && ctas > sms / L) ctas = sms / L > 0 ? sms / L : 1;
    const dim3 grid((unsigned)ctas, (unsigned)L);
    WM_CUDA_CHECK(cudaFuncSetAttribute(vq_nearest_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vq_nearest_tc_kernel<64><<<grid, kThreads, smem, st>>>(prm);
    WM_CUDA_CHECK(cudaGetLastError());
    const long rounds = (N * L + kSettleRound - 1) / kSettleRound;
    const long wave = 4L * sms;                            // 64 registers x 256 threads: four blocks per SM
    vq_settle_kernel<64><<<(unsigned)(rounds < wave ? rounds : wave), kSettleThreads, 0, st>>>(prm);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}

}  // namespace wm
