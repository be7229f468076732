// Backward of the local windowed 3D attention on tcgen05 tensor cores (bf16, fp32 accumulate).
//
// Same brick / halo-block tiling as the forward kernel (attn_tc.cu).  Two kernels, both
// deterministic (no atomics):
//
//   dQ kernel   (query-stationary)  rows = 128 queries of a brick, blocks = K/V halo
//       S  = Q K_j^T, dP = dO V_j^T              (two tcgen05.mma chains into TMEM)
//       dS = exp2(S*c - lse) * (dP - delta) * scale   (one thread per query row)
//       dQ += dS K_j                              (K_j read MN-major from the same smem)
//   dK/dV kernel (key-stationary)   rows = 128 keys of a brick, blocks = Q/dO halo.  The
//       queries that see a key are exactly the keys that key would see (the window is
//       symmetric), so the halo geometry and the masks are the forward's, transposed:
//       S^T = K Q_j^T, dP^T = V dO_j^T
//       P^T = exp2(S^T*c - lse_col), dS^T = P^T * (dP^T - delta_col) * scale
//       dV += P^T dO_j,  dK += dS^T Q_j
//   plus a small pre-pass delta = rowsum(dO * O).
//
// Replaces the autograd graph of Local3dAttention.local_attention
// (local_3d_attention.py:78-99; recomputed under checkpoint at :110-111).
#include "attn_tc.cuh"

// -DWM_EXPERIMENT=7 compiles the clock64 timeline instrumentation in (tools/build_timeline_lib.sh, tools/dbg_timeline.py)
#ifndef WM_EXPERIMENT
#define WM_EXPERIMENT 0
#endif
#if WM_EXPERIMENT == 7
namespace wm { namespace tc { __device__ long long g_dbg_bwd[2 * 64 * 16]; } }
#define DBGB(slot) do { if (dbg_on && t < 64) g_dbg_bwd[(MODE - 1) * 1024 + t * 16 + (slot)] = clock64(); } while (0)
#else
#define DBGB(slot) do { } while (0)
#endif

#include <math.h>

namespace wm {
namespace tc {

struct BwdParams {
    AttnShape sh;
    Plan pl;
    const float* lse;        // [B,S,H,W,heads] natural-log LSE from the forward
    const float* delta;      // [B,S,H,W,heads] rowsum(dO*O)
    __nv_bfloat16* out1;     // dQ kernel: dq.   dK/dV kernel: dv
    __nv_bfloat16* out2;     //                  dK/dV kernel: dk
};

// ---------------------------------------------------------------------------- delta
__global__ void __launch_bounds__(256)
l3d_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                 float* __restrict__ delta, long items, int d) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;       // (token, head)
    if (i >= items) return;
    const uint4* po = reinterpret_cast<const uint4*>(o + i * d);
    const uint4* pg = reinterpret_cast<const uint4*>(dout + i * d);
    float acc = 0.f;
    for (int c = 0; c < d / 8; ++c) {
        const uint4 a = __ldg(po + c), g = __ldg(pg + c);
        const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* hg = reinterpret_cast<const __nv_bfloat162*>(&g);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 fa = __bfloat1622float2(ha[j]), fg = __bfloat1622float2(hg[j]);
            acc = fmaf(fa.x, fg.x, acc);
            acc = fmaf(fa.y, fg.y, acc);
        }
    }
    delta[i] = acc;
}

// --------------------------------------------------------------------------- kernel
// 256 threads: warps w and w+4 share TMEM lane quadrant w&3 and split the row's columns.  A CTA
// walks `hpc` heads of its brick (row tiles double-buffered when they fit); thread 0 issues TMA
// and MMA.  Steps are (head, halo plane, h-chunk).
template <int D, int MODE>
__global__ void __launch_bounds__(kThreads, (D == 128) ? 1 : 2)
l3d_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_a2,
                  const __grid_constant__ CUtensorMap map_b1, const __grid_constant__ CUtensorMap map_b2,
                  const BwdParams prm) {
    using G = Geo<D>;
    constexpr bool kDKV = (MODE == kBwdDKV);
    const AttnShape& sh = prm.sh;
    const Plan& pl = prm.pl;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays in the shared window

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane;
    const int ncols = pl.ncols, ncols_pad = pl.ncols_pad;
    const int row_slab_bytes = 128 * G::kRowBytes;
    const int row_tile_bytes = G::kSlabs * row_slab_bytes;
    const int blk_slab_bytes = ncols_pad * G::kRowBytes;
    const int blk_tile_bytes = G::kSlabs * blk_slab_bytes;
    const int p_slabs = (ncols_pad + 63) / 64;
    const int p_tile_bytes = p_slabs * 128 * 128;

    uint8_t* sA = smem;                                    // [rowbuf][A1 | A2][slabs][128 rows]  (Q,dO | K,V)
    uint8_t* sB = sA + pl.rowbuf * 2 * row_tile_bytes;     // [stage][operand][slab][ncols_pad rows]
    uint8_t* sDS = sB + 4 * blk_tile_bytes;                // dS (or dS^T), bf16, K-major 128B swizzle
    uint8_t* sPT = sDS + p_tile_bytes;                     // P^T (dK/dV kernel only)
    uint32_t* sMask = reinterpret_cast<uint32_t*>(sPT + (kDKV ? p_tile_bytes : 0));
    float* sCol = reinterpret_cast<float*>(sMask + 2 * 9 * 128);           // [2 bufs][lse2|delta][ncols_pad]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sCol + (kDKV ? 4 * ncols_pad : 0));
    uint64_t* bar_a = bars;           // [2]
    uint64_t* bar_b = bars + 2;       // [2]
    uint64_t* bar_mma = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
    uint32_t* myMask = sMask + half * 9 * 128;                              // [9 words][128 rows], own copy

    const int tw_i = blockIdx.x % pl.tilesW, th_i = blockIdx.x / pl.tilesW;
    const int ts_i = blockIdx.y;
    const int hgroups = sh.heads / pl.hpc;
    const int hg = blockIdx.z % hgroups, b = blockIdx.z / hgroups;
    const int head0 = hg * pl.hpc;
    const int s0 = ts_i * pl.tS, h0 = th_i * pl.tH, w0 = tw_i * pl.tW;

    // ---- this thread's row (a query in the dQ kernel, a key in the dK/dV kernel) --------------
    const int plane_mask = (1 << pl.lgPlane) - 1;
    const int rs = row >> pl.lgPlane, rh = (row & plane_mask) >> pl.lgTW, rw = row & (pl.tW - 1);
    const bool row_valid = (s0 + rs < sh.S) && (h0 + rh < sh.H) && (w0 + rw < sh.W);
    const int kh_lo = max(rh, sh.eH - h0), kh_hi = min(rh + 2 * sh.eH, sh.H - 1 - h0 + sh.eH);
    const int kw_lo = max(rw, sh.eW - w0), kw_hi = min(rw + 2 * sh.eW, sh.W - 1 - w0 + sh.eW);
    const uint32_t wbits = (row_valid && kw_hi >= kw_lo) ? ((kw_hi - kw_lo == 31) ? 0xffffffffu : ((1u << (kw_hi - kw_lo + 1)) - 1u)) << kw_lo : 0u;
    const int w_rs = (quad * 32) >> pl.lgPlane;
    const int w_rh_lo = ((quad * 32) & plane_mask) >> pl.lgTW, w_rh_hi = ((quad * 32 + 31) & plane_mask) >> pl.lgTW;
    const int ks_first = max(0, sh.eS - s0), ks_last = min(pl.hS - 1, sh.S - 1 - s0 + sh.eS);
    const int khg_lo = max(0, sh.eH - h0), khg_hi = min(pl.hH - 1, sh.H - 1 - h0 + sh.eH);
    int chunk_first = 0, chunk_last = 0;
    for (int c = 0; c < pl.nchunk; ++c) {
        if (khg_lo >= (c + 1) * pl.ch) chunk_first = c + 1;
        if (khg_hi >= c * pl.ch) chunk_last = c;
    }
    const int nplanes = ks_last - ks_first + 1;
    const int nblocks = nplanes * (chunk_last - chunk_first + 1);
    const int nsteps = nblocks * pl.hpc;
    const long row_tok = (((long)b * sh.S + (s0 + rs)) * sh.H + (h0 + rh)) * sh.W + (w0 + rw);
    constexpr float kLog2e = 1.4426950408889634f;

    if (tid == 0) {
        tma_prefetch_desc(&map_a1); tma_prefetch_desc(&map_a2); tma_prefetch_desc(&map_b1); tma_prefetch_desc(&map_b2);
        mbar_init(&bar_a[0], 1);
        mbar_init(&bar_a[1], 1);
        mbar_init(&bar_b[0], 1);
        mbar_init(&bar_b[1], 1);
        mbar_init(bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        if (pl.tmem_cols <= 128) tmem_alloc<128>(tmem_slot);
        else if (pl.tmem_cols <= 256) tmem_alloc<256>(tmem_slot);
        else tmem_alloc<512>(tmem_slot);
    }
    if (ncols_pad > ncols) {      // rows TMA never writes must stay finite (they meet zero dS / P columns)
        const int pad_bytes = (ncols_pad - ncols) * G::kRowBytes;
        for (int t = 0; t < 4 * G::kSlabs; ++t) {
            uint8_t* base = sB + t * blk_slab_bytes + ncols * G::kRowBytes;
            for (int i = tid * 16; i < pad_bytes; i += kThreads * 16) *reinterpret_cast<uint4*>(base + i) = make_uint4(0, 0, 0, 0);
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_acc1 = tmem_base;                                   // dQ | dV
    const uint32_t tmem_acc2 = tmem_base + D;                               //      dK
    const uint32_t tmem_t1 = tmem_base + (kDKV ? 2 * D : D);                // S   | S^T
    const uint32_t tmem_t2 = tmem_t1 + ncols_pad;                           // dP  | dP^T
    const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;

    struct Cursor { int hd, ks, chunk; };
    auto advance = [&](Cursor& c) {
        if (++c.ks > ks_last) {
            c.ks = ks_first;
            if (++c.chunk > chunk_last) { c.chunk = chunk_first; ++c.hd; }
        }
    };
    // TMA / MMA issue: executed by the whole of warp 0 so that addresses and descriptors stay warp-uniform
    // (uniform datapath); only the instructions themselves are predicated on one lane.
    const bool leader = (lane == 0);
    auto a_buf = [&](int hd) { return sA + (pl.rowbuf == 2 ? (hd & 1) : 0) * 2 * row_tile_bytes; };
    auto issue_row_load = [&](int hd) {
        if (leader) {
            uint64_t* bar = &bar_a[hd & 1];
            const int cb = (head0 + hd) * D;
            mbar_expect_tx(bar, 2u * (uint32_t)row_tile_bytes);
#pragma unroll
            for (int sl = 0; sl < G::kSlabs; ++sl) {
                tma_load_5d(a_buf(hd) + sl * row_slab_bytes, &map_a1, bar, cb + sl * G::kSlabCh, w0, h0, s0, b);
                tma_load_5d(a_buf(hd) + row_tile_bytes + sl * row_slab_bytes, &map_a2, bar, cb + sl * G::kSlabCh, w0, h0, s0, b);
            }
        }
    };
    auto issue_block_load = [&](int t, const Cursor& c) {
        if (leader) {
            const int stage = t & 1;
            const int cb = (head0 + c.hd) * D;
            uint8_t* dst = sB + stage * 2 * blk_tile_bytes;
            mbar_expect_tx(&bar_b[stage], 2u * G::kSlabs * (uint32_t)ncols * G::kRowBytes);
#pragma unroll
            for (int sl = 0; sl < G::kSlabs; ++sl) {
                tma_load_5d(dst + sl * blk_slab_bytes, &map_b1, &bar_b[stage], cb + sl * G::kSlabCh, w0 - sh.eW,
                            h0 - sh.eH + c.chunk * pl.ch, s0 - sh.eS + c.ks, b);
                tma_load_5d(dst + blk_tile_bytes + sl * blk_slab_bytes, &map_b2, &bar_b[stage], cb + sl * G::kSlabCh,
                            w0 - sh.eW, h0 - sh.eH + c.chunk * pl.ch, s0 - sh.eS + c.ks, b);
            }
        }
    };
    const uint32_t idesc_t = make_idesc_bf16(ncols_pad, false, false);
    const uint32_t idesc_acc = make_idesc_bf16(D, false, true);
    // descriptor bases; byte offsets are added as (offset >> 4) to the low (start address) field
    const uint64_t da0 = make_smem_desc(smem_u32(sA), 16, G::kAtomBytes, G::kSwizzleCode);
    const uint64_t dbk0 = make_smem_desc(smem_u32(sB), 16, G::kAtomBytes, G::kSwizzleCode);                         // block, K-major
    const uint64_t dbm0 = make_smem_desc(smem_u32(sB), (uint32_t)blk_slab_bytes, G::kAtomBytes, G::kSwizzleCode);   // block, MN-major
    const uint64_t dds0 = make_smem_desc(smem_u32(sDS), 16, 1024, 2u);
    const uint64_t dpt0 = make_smem_desc(smem_u32(sPT), 16, 1024, 2u);
    const uint32_t a_buf_step = (pl.rowbuf == 2) ? (uint32_t)((2 * row_tile_bytes) >> 4) : 0u;
    const uint32_t stage_step = (uint32_t)((2 * blk_tile_bytes) >> 4);
    auto issue_t_mma = [&](int t, int hd) {         // T1 = A1 B1_t^T, T2 = A2 B2_t^T
        const uint64_t da_h = da0 + (hd & 1) * a_buf_step, db_s = dbk0 + (t & 1) * stage_step;
#pragma unroll
        for (int op = 0; op < 2; ++op) {
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk) {
                const uint32_t sl = (uint32_t)((kk * 16) / G::kSlabCh);
                const uint32_t koff = (uint32_t)((((kk * 16) % G::kSlabCh) * 2) >> 4);
                const uint64_t da = da_h + op * (uint32_t)(row_tile_bytes >> 4) + sl * (uint32_t)(row_slab_bytes >> 4) + koff;
                const uint64_t db = db_s + op * (uint32_t)(blk_tile_bytes >> 4) + sl * (uint32_t)(blk_slab_bytes >> 4) + koff;
                if (leader) umma_bf16_ss(op ? tmem_t2 : tmem_t1, da, db, idesc_t, kk > 0);
            }
        }
    };
    const int nk_acc = ncols_pad / 16;
    auto issue_acc_mma = [&](int t, bool accumulate) {
        uint64_t d_ds = dds0, d_pt = dpt0;
        uint64_t d_b1 = dbm0 + (t & 1) * stage_step;
        uint64_t d_b2 = d_b1 + (uint32_t)(blk_tile_bytes >> 4);
        for (int kk = 0; kk < nk_acc; ++kk) {
            const uint32_t acc = (accumulate || kk > 0) ? 1u : 0u;
            if constexpr (kDKV) {
                if (leader) {
                    umma_bf16_ss(tmem_acc1, d_pt, d_b2, idesc_acc, acc);      // dV += P^T dO_t
                    umma_bf16_ss(tmem_acc2, d_ds, d_b1, idesc_acc, acc);      // dK += dS^T Q_t
                }
            } else {
                if (leader) umma_bf16_ss(tmem_acc1, d_ds, d_b1, idesc_acc, acc);      // dQ += dS K_t
            }
            const uint32_t a_step = ((kk & 3) == 3) ? (uint32_t)((128 * 128 - 96) >> 4) : 2u;
            d_ds += a_step; d_pt += a_step;
            d_b1 += (uint32_t)((16 * G::kRowBytes) >> 4); d_b2 += (uint32_t)((16 * G::kRowBytes) >> 4);
        }
    };
    // per-column lse / delta of a block's queries (dK/dV kernel): one column per thread
    auto load_colvec = [&](const Cursor& c, float& lse2, float& dl) {
        const int gs = s0 - sh.eS + c.ks;
        const int col = tid;
        lse2 = 0.f;
        dl = 0.f;
        if (col < ncols) {
            const int khl = col / pl.hW, kw = col - khl * pl.hW;
            const int gh = h0 - sh.eH + c.chunk * pl.ch + khl, gw = w0 - sh.eW + kw;
            if (gs >= 0 && gs < sh.S && gh >= 0 && gh < sh.H && gw >= 0 && gw < sh.W) {
                const long idx = ((((long)b * sh.S + gs) * sh.H + gh) * sh.W + gw) * sh.heads + head0 + c.hd;
                lse2 = __ldg(prm.lse + idx) * kLog2e;
                dl = __ldg(prm.delta + idx) * sh.scale;
            }
        }
    };
    auto store_colvec = [&](int buf, float lse2, float dl) {
        if (tid < ncols_pad) {
            sCol[(buf * 2 + 0) * ncols_pad + tid] = lse2;
            sCol[(buf * 2 + 1) * ncols_pad + tid] = dl;
        }
    };
    auto store_group = [&](uint8_t* tile, int g, const uint32_t (&pk)[8]) {    // 16 bf16 of this row -> swizzled tile
        const uint32_t slab_off = (g >> 2) * (128 * 128);
        const int c16 = (g & 3) * 2;
        *reinterpret_cast<uint4*>(tile + slab_off + sw128_offset(row, c16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(tile + slab_off + sw128_offset(row, c16 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    };
    // accumulators of head `hd` -> bf16 -> global; called by all threads once its last step has retired
    auto finish_head = [&](int hd) {
        const long row_off = row_tok * (long)sh.inner() + (head0 + hd) * D + half * (D / 2);
#pragma unroll
        for (int which = 0; which < (kDKV ? 2 : 1); ++which) {
            __nv_bfloat16* dst = (which == 0 ? prm.out1 : prm.out2) + row_off;
            const uint32_t src = (which == 0 ? tmem_acc1 : tmem_acc2) + lane_sel + half * (D / 2);
#pragma unroll
            for (int c = 0; c < D / 2; c += 16) {
                uint32_t r[16];
                tmem_ld16(src + c, r);
                tmem_wait_ld();
                if (row_valid) {
                    uint32_t pk[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
                    *reinterpret_cast<uint4*>(dst + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4*>(dst + c + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
            }
        }
    };

    Cursor cur{0, ks_first, chunk_first};
    Cursor nxt = cur;
    advance(nxt);
    if (warp == 0) {
        issue_row_load(0);
        issue_block_load(0, cur);
        if (nsteps > 1) issue_block_load(1, nxt);
        mbar_wait(&bar_a[0], 0);
        mbar_wait(&bar_b[0], 0);
        tc_fence_after();
        issue_t_mma(0, 0);
        if (leader) umma_commit(bar_mma);
    }
    float row_lse2 = 0.f, row_delta = 0.f;
    if constexpr (kDKV) {
        float a, c;
        load_colvec(cur, a, c);
        store_colvec(0, a, c);
        __syncthreads();
    }

    bool p_zero = false;
    int mask_chunk = -1;
    int g_lo = 0, g_hi = 0;
    bool chunk_live = false;
    const int nwords = (ncols_pad + 31) / 32;
    const int ngroups = ncols_pad >> 4;
    const uint32_t zero8[8] = {0, 0, 0, 0, 0, 0, 0, 0};

    const bool dbg_on = (blockIdx.x == 1 && blockIdx.y == 1 && blockIdx.z == 0 && tid == 0);
    (void)dbg_on;
    for (int t = 0; t < nsteps; ++t) {
        DBGB(0);
        const bool head_start = (cur.ks == ks_first) && (cur.chunk == chunk_first);
        float nxt_lse2 = 0.f, nxt_dl = 0.f;
        if constexpr (kDKV) {
            if (t + 1 < nsteps) load_colvec(nxt, nxt_lse2, nxt_dl);        // global loads in flight during the wait
        } else {
            if (head_start && row_valid) {
                row_lse2 = __ldg(prm.lse + row_tok * sh.heads + head0 + cur.hd) * kLog2e;
                row_delta = __ldg(prm.delta + row_tok * sh.heads + head0 + cur.hd) * sh.scale;
            }
        }
        mbar_wait(bar_mma, t & 1);                  // T_t ready; accumulations of step t-1 retired
        tc_fence_after();
        DBGB(1);
        if (warp == 0) {
            if (head_start && pl.rowbuf == 2 && cur.hd + 1 < pl.hpc) issue_row_load(cur.hd + 1);
            if (t >= 1 && t + 1 < nsteps) issue_block_load(t + 1, nxt);
        }
        if (head_start && t > 0) finish_head(cur.hd - 1);

        const int kh0 = cur.chunk * pl.ch;
        if (cur.chunk != mask_chunk) {
            mask_chunk = cur.chunk;
            for (int w = 0; w <= nwords; ++w) myMask[w * 128 + row] = 0u;
            if (wbits != 0u) {
                const int ra = max(kh_lo, kh0), rb = min(kh_hi, kh0 + pl.ch - 1);
                for (int kh = ra; kh <= rb; ++kh) {
                    const int pos = (kh - kh0) * pl.hW;
                    const int w = pos >> 5, sft = pos & 31;
                    myMask[w * 128 + row] |= wbits << sft;
                    if (sft != 0 && (wbits >> (32 - sft)) != 0u) myMask[(w + 1) * 128 + row] |= wbits >> (32 - sft);
                }
            }
            const int ua = max(w_rh_lo, kh0), ub = min(w_rh_hi + 2 * sh.eH, kh0 + pl.ch - 1);
            chunk_live = ub >= ua;
            g_lo = ((ua - kh0) * pl.hW) >> 4;
            g_hi = min(((ub - kh0 + 1) * pl.hW + 15) >> 4, ngroups);
        }

        const bool live = chunk_live && (cur.ks >= w_rs) && (cur.ks <= w_rs + 2 * sh.eS);
        if (live) {
            const int g_mid = (g_lo + g_hi + 1) >> 1;
            const int ga = half ? g_mid : g_lo, gb = half ? g_hi : g_mid;            // live groups of this thread
            const int za = half ? g_hi : 0, zb = half ? ngroups : g_lo;              // groups this thread zero-fills
            const float* col_lse2 = sCol + ((t & 1) * 2 + 0) * ncols_pad;
            const float* col_dl = sCol + ((t & 1) * 2 + 1) * ncols_pad;
            for (int g = ga; g < gb; ++g) {
                const uint32_t mword = myMask[(g >> 1) * 128 + row] >> ((g & 1) * 16);
                uint32_t s[16], dp[16];
                tmem_ld16(tmem_t1 + lane_sel + g * 16, s);
                tmem_ld16(tmem_t2 + lane_sel + g * 16, dp);
                float l2[16], dl[16];
                if constexpr (kDKV) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 a = *reinterpret_cast<const float4*>(col_lse2 + g * 16 + 4 * i);
                        const float4 c = *reinterpret_cast<const float4*>(col_dl + g * 16 + 4 * i);
                        l2[4 * i] = a.x; l2[4 * i + 1] = a.y; l2[4 * i + 2] = a.z; l2[4 * i + 3] = a.w;
                        dl[4 * i] = c.x; dl[4 * i + 1] = c.y; dl[4 * i + 2] = c.z; dl[4 * i + 3] = c.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) { l2[i] = row_lse2; dl[i] = row_delta; }
                }
                tmem_wait_ld();
                float pv[16], dsv[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    // dl[] holds delta * scale: dS = P * (dP * scale - delta * scale); masking P masks dS with it
                    const float p = ex2(fmaf(__uint_as_float(s[i]), pl.scale_log2, -l2[i]));
                    pv[i] = ((mword >> i) & 1u) ? p : 0.f;
                    dsv[i] = pv[i] * fmaf(__uint_as_float(dp[i]), sh.scale, -dl[i]);
                }
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(dsv[2 * i], dsv[2 * i + 1]);
                store_group(sDS, g, pk);
                if constexpr (kDKV) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(pv[2 * i], pv[2 * i + 1]);
                    store_group(sPT, g, pk);
                }
            }
            for (int g = za; g < zb; ++g) {
                store_group(sDS, g, zero8);
                if constexpr (kDKV) store_group(sPT, g, zero8);
            }
            p_zero = false;
        } else if (!p_zero) {
            const int za = half ? (ngroups >> 1) : 0, zb = half ? ngroups : (ngroups >> 1);
            for (int g = za; g < zb; ++g) {
                store_group(sDS, g, zero8);
                if constexpr (kDKV) store_group(sPT, g, zero8);
            }
            p_zero = true;
        }
        if constexpr (kDKV) {
            if (t + 1 < nsteps) store_colvec((t + 1) & 1, nxt_lse2, nxt_dl);
        }
        DBGB(2);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        DBGB(3);
        if (warp == 0) {
            tc_fence_after();
            issue_acc_mma(t, !head_start);
            DBGB(4);
            if (t + 1 < nsteps) {
                mbar_wait(&bar_b[(t + 1) & 1], ((t + 1) >> 1) & 1);
                if (nxt.hd != cur.hd) mbar_wait(&bar_a[nxt.hd & 1], (nxt.hd >> 1) & 1);   // hpc > 1 implies two row buffers
                tc_fence_after();
                issue_t_mma(t + 1, nxt.hd);
            }
            if (leader) umma_commit(bar_mma);
            DBGB(5);
        }
        cur = nxt;
        advance(nxt);
    }

    mbar_wait(bar_mma, nsteps & 1);
    tc_fence_after();
    finish_head(pl.hpc - 1);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        if (pl.tmem_cols <= 128) tmem_dealloc<128>(tmem_base);
        else if (pl.tmem_cols <= 256) tmem_dealloc<256>(tmem_base);
        else tmem_dealloc<512>(tmem_base);
    }
}

template <int D, int MODE>
static int launch_one(const void* a1, const void* a2, const void* b1, const void* b2, const float* lse,
                      const float* delta, void* out1, void* out2, const AttnShape& s, cudaStream_t st) {
    using G = Geo<D>;
    Plan pl;
    if (!make_plan(s, (Mode)MODE, pl)) return fail(WM_EUNSUPPORTED, "no tensor-core backward tiling for this shape");
    CUtensorMap ma1, ma2, mb1, mb2;
    const int C = s.inner();
    if (int rc = make_tensor_map_5d(&ma1, a1, s.B, s.S, s.H, s.W, C, G::kSlabCh, pl.tW, pl.tH, pl.tS, G::kSwizzleBytes)) return rc;
    if (int rc = make_tensor_map_5d(&ma2, a2, s.B, s.S, s.H, s.W, C, G::kSlabCh, pl.tW, pl.tH, pl.tS, G::kSwizzleBytes)) return rc;
    if (int rc = make_tensor_map_5d(&mb1, b1, s.B, s.S, s.H, s.W, C, G::kSlabCh, pl.hW, pl.ch, 1, G::kSwizzleBytes)) return rc;
    if (int rc = make_tensor_map_5d(&mb2, b2, s.B, s.S, s.H, s.W, C, G::kSlabCh, pl.hW, pl.ch, 1, G::kSwizzleBytes)) return rc;
    BwdParams prm{s, pl, lse, delta, static_cast<__nv_bfloat16*>(out1), static_cast<__nv_bfloat16*>(out2)};
    WM_CUDA_CHECK(cudaFuncSetAttribute(l3d_bwd_tc_kernel<D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem_bytes));
    const dim3 grid((unsigned)(pl.tilesW * pl.tilesH), (unsigned)pl.tilesS, (unsigned)(s.B * (s.heads / pl.hpc)));
    if (grid.y > 65535u || grid.z > 65535u) return fail(WM_EUNSUPPORTED, "grid too large for the tensor-core kernel");
    l3d_bwd_tc_kernel<D, MODE><<<grid, kThreads, pl.smem_bytes, st>>>(ma1, ma2, mb1, mb2, prm);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}

template <int D>
static int launch_bwd_d(const void* q, const void* k, const void* v, const void* o, const float* lse, const void* dout,
                        void* dq, void* dk, void* dv, float* delta, const AttnShape& s, cudaStream_t st) {
    const long items = s.tokens() * s.heads;
    l3d_delta_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(o),
                                                                      static_cast<const __nv_bfloat16*>(dout), delta,
                                                                      items, s.d);
    WM_CUDA_CHECK(cudaGetLastError());
    // warp-specialised kernels where the shape has a tiling for them (dim_head <= 64), else the two-role kernels
    int rc = launch_bwd_ws(kBwdDQws, q, dout, k, v, lse, delta, dq, nullptr, s, st);
    if (rc == WM_EUNSUPPORTED) rc = launch_one<D, kBwdDQ>(q, dout, k, v, lse, delta, dq, nullptr, s, st);
    if (rc) return rc;
    rc = launch_bwd_ws(kBwdDKVws, k, v, q, dout, lse, delta, dv, dk, s, st);
    if (rc == WM_EUNSUPPORTED) rc = launch_one<D, kBwdDKV>(k, v, q, dout, lse, delta, dv, dk, s, st);
    return rc;
}

#if WM_EXPERIMENT == 7
extern "C" __attribute__((visibility("default"))) int wm_debug_read_bwd(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, g_dbg_bwd, sizeof(long long) * 2 * 64 * 16);
}
#endif

int launch_bwd_tc(const void* q, const void* k, const void* v, const void* o, const float* lse, const void* dout,
                  void* dq, void* dk, void* dv, float* delta, const AttnShape& s, cudaStream_t st) {
    switch (s.d) {
        case 32: return launch_bwd_d<32>(q, k, v, o, lse, dout, dq, dk, dv, delta, s, st);
        case 64: return launch_bwd_d<64>(q, k, v, o, lse, dout, dq, dk, dv, delta, s, st);
        case 128: return launch_bwd_d<128>(q, k, v, o, lse, dout, dq, dk, dv, delta, s, st);
        default: return fail(WM_EUNSUPPORTED, "dim_head=%d has no tensor-core kernel", s.d);
    }
}

}  // namespace tc
}  // namespace wm
