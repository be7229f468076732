// Warp-specialised backward kernels of the local windowed 3D attention.
//
// Same math, brick / halo-block tiling, masks and determinism as attn_tc_bwd.cu; what changes is
// who does what and where the probability tiles live:
//   warps 0-7  compute (two threads per brick row)
//   warp 8     issues the (S, dP) MMAs of the NEXT step, in two column halves (A, then B)
//   warp 9     issues the accumulating MMAs of the current step
//   warp 10    TMA loader (row tiles of the next head, halo blocks nstage steps ahead)
//   T = (S, dP) accumulators: one TMEM buffer.  Half A of step t+1 is produced as soon as every
//       compute thread is done with half A of step t, i.e. while they work on half B (and vice versa),
//       so the compute warps never wait for the tensor pipe.
//   dQ kernel : dS (bf16) goes back to TENSOR MEMORY (two buffers) and is the A operand of
//               dQ += dS K (TS-mode MMA): no shared-memory round trip
//   dK/dV     : P^T in tensor memory (two buffers, A of dV += P^T dO), dS^T in shared memory
//               (two buffers, A of dK += dS^T Q)
// One issuing thread costs ~20-60 cycles per tcgen05.mma (tools/micro/mma_issue_clean.cu); three
// single-purpose warps keep each of the per-step chains short.
//
// The window / border mask is part of the S (resp. S^T) MMA (build_mask_tiles in attn_tc.cuh); the element loops
// run on packed fp32x2 arithmetic with part of the exponentials on the FMA pipe (exp2_poly2).
#include "attn_tc.cuh"

// -DWM_EXPERIMENT=7 compiles the clock64 timeline instrumentation in (tools/build_timeline_lib.sh, tools/dbg_timeline.py)
#ifndef WM_EXPERIMENT
#define WM_EXPERIMENT 0
#endif
#ifndef WM_WS_PROBE
#define WM_WS_PROBE 1
#endif
#if WM_EXPERIMENT == 7
namespace wm { namespace tc { __device__ long long g_dbg_ws[2 * 64 * 16]; } }
#define DBGW(slot) do { if (dbg_on && t < 64) g_dbg_ws[(MODE - 3) * 1024 + t * 16 + (slot)] = clock64(); } while (0)
#else
#define DBGW(slot) do { } while (0)
#endif

#include <math.h>
#include <stdlib.h>
#include <type_traits>

namespace wm {
namespace tc {

#ifndef WM_BWD_POLY
#define WM_BWD_POLY 2      // of every 8 column pairs, this many take the FMA-pipe exp2 (the rest go to the MUFU)
#endif

// ---------------------------------------------------------------------------- delta = rowsum(dO * O)
__global__ void __launch_bounds__(256)
l3d_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                 float* __restrict__ delta, long items, int d) {
    // d / 8 consecutive lanes share one (token, head): every load instruction of a warp covers 512 contiguous bytes
    const int lanes = d >> 3;                                            // 4, 8 or 16 (d = 32, 64, 128)
    const long chunk = (long)blockIdx.x * blockDim.x + threadIdx.x;      // 16-byte chunk of the [items, d] tensors
    const long i = chunk / lanes;
    float acc = 0.f;
    if (i < items) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(o) + chunk), g = __ldg(reinterpret_cast<const uint4*>(dout) + chunk);
        const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* hg = reinterpret_cast<const __nv_bfloat162*>(&g);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 fa = __bfloat1622float2(ha[j]), fg = __bfloat1622float2(hg[j]);
            acc = fmaf(fa.x, fg.x, acc);
            acc = fmaf(fa.y, fg.y, acc);
        }
    }
    for (int sft = lanes >> 1; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
    if (i < items && (threadIdx.x & (lanes - 1)) == 0) delta[i] = acc;
}

struct BwdWsParams {
    AttnShape sh;
    Plan pl;
    const float* lse;
    const float* delta;
    __nv_bfloat16* out1;     // dQ kernel: dq.   dK/dV kernel: dv
    __nv_bfloat16* out2;     //                  dK/dV kernel: dk
    long ld_out;             // elements between consecutive tokens of the outputs
};

constexpr int kWsThreads = 352;        // 8 compute warps + (S,dP) issuer + accumulate issuer + TMA loader

template <int D, int MODE>
__global__ void __launch_bounds__(kWsThreads, 1)
l3d_bwd_ws_kernel(const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_a2,
                  const __grid_constant__ CUtensorMap map_b1, const __grid_constant__ CUtensorMap map_b2,
                  const BwdWsParams prm) {
    using G = Geo<D>;
    constexpr bool kDKV = (MODE == kBwdDKVws);
    const AttnShape& sh = prm.sh;
    const Plan& pl = prm.pl;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ncols = pl.ncols, ncols_pad = pl.ncols_pad;
    const int nstage = pl.nstage;
    const int row_slab_bytes = 128 * G::kRowBytes;
    const int row_tile_bytes = G::kSlabs * row_slab_bytes;
    const int blk_slab_bytes = ncols_pad * G::kRowBytes;
    const int blk_tile_bytes = G::kSlabs * blk_slab_bytes;
    const int p_slabs = (ncols_pad + 63) / 64;
    const int p_tile_bytes = p_slabs * 128 * 128;

    uint8_t* sA = smem;                                    // [rowbuf][A1 | A2][slabs][128 rows]  (Q,dO | K,V)
    uint8_t* sB = sA + pl.rowbuf * 2 * row_tile_bytes;     // [stage][B1 | B2][slab][ncols_pad rows]  (K,V | Q,dO)
    uint8_t* sDS = sB + nstage * 2 * blk_tile_bytes;       // dK/dV kernel: dS^T x2, bf16, K-major 128B swizzle
    const int cm_tile_bytes = ncols_pad * pl.km * 2;
    uint8_t* sRm = sDS + (kDKV ? 2 * p_tile_bytes : 0);    // mask operand of the rows:    [128 x km] bf16
    uint8_t* sCm = sRm + 128 * pl.km * 2;                  // mask operand of the columns: [nchunk][ncols_pad x km] bf16
    float* sCol = reinterpret_cast<float*>(sCm + pl.nchunk * cm_tile_bytes);              // [3 bufs][-lse2|-delta][ncols_pad]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sCol + (kDKV ? 6 * ncols_pad : 0));
    uint64_t* bar_a = bars;           // [2] row tiles of a head landed
    uint64_t* bar_b = bars + 2;       // [3] halo block landed
    // (S, dP) of a step are produced and drained in two column halves, A = [0, nA) and B = [nA, ncols_pad): half A of
    // step t+1 is issued as soon as every compute thread is done with half A of step t, i.e. while they work on half B.
    uint64_t* bar_tA = bars + 5;      //     columns A of (S, dP) of a step computed     (tcgen05.commit)
    uint64_t* bar_p = bars + 6;       // [2] all of dS / P written, columns B drained     (8 compute warps)
    uint64_t* bar_acc = bars + 8;     // [2] accumulating MMAs of a step retired          (tcgen05.commit)
    uint64_t* bar_tB = bars + 10;     //     columns B of (S, dP) of a step computed     (tcgen05.commit)
    uint64_t* bar_pA = bars + 11;     // [2] columns A drained                            (8 compute warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
    const int gA = (ncols_pad >> 4) >= 2 ? (ncols_pad >> 5) : 0;   // 16-column groups in half A (0: a single half)
    const int nA = gA * 16;

    // Persistent CTAs.  A CTA serves ONE (h, w) brick position -- its masks, border ranges and live column ranges depend on
    // nothing else -- and walks the work items (batch element, s position, head group) of that position back to back:
    // TMEM, barriers and mask tiles are set up once, and the loader runs ahead across item boundaries, so the pipeline
    // never drains between bricks (a one-item CTA spent ~16 % of its life in prologue and drain).
    const int nclass = pl.tilesH * pl.tilesW;
    const int cls = (int)blockIdx.x % nclass, rank = (int)blockIdx.x / nclass;
    const int nrank = ((int)gridDim.x - cls + nclass - 1) / nclass;          // CTAs sharing this brick position
    const int tw_i = cls % pl.tilesW, th_i = cls / pl.tilesW;
    const int hgroups = sh.heads / pl.hpc;
    const int n_items = sh.B * pl.tilesS * hgroups;
    const int h0 = th_i * pl.tH, w0 = tw_i * pl.tW;
    const int khg_lo = max(0, sh.eH - h0), khg_hi = min(pl.hH - 1, sh.H - 1 - h0 + sh.eH);
    int chunk_first = 0, chunk_last = 0;
    for (int c = 0; c < pl.nchunk; ++c) {
        if (khg_lo >= (c + 1) * pl.ch) chunk_first = c + 1;
        if (khg_hi >= c * pl.ch) chunk_last = c;
    }
    // a step = (item, head, h-chunk, plane); H counts heads across items (buffer / barrier parities run on it)
    // The items of this CTA live in a small shared-memory table {b, s0, head0, ks_first, ks_last}; a cursor is five
    // integers (slot, head, plane, chunk, H), so that the three or four cursors a warp keeps cost few registers.
    struct Item { int b, s0, head0, ks_first, ks_last, pad0, pad1, pad2; };
    Item* sItems = reinterpret_cast<Item*>(bars + 16);                  // [kMaxItemsPerCta], inside the fixed part of the budget
    const int my_items = rank < n_items ? (n_items - rank + nrank - 1) / nrank : 0;
    if (tid < my_items) {
        const int item = rank + tid * nrank;
        const int hg = item % hgroups, ts_i = (item / hgroups) % pl.tilesS;
        Item it;
        it.b = item / (hgroups * pl.tilesS);
        it.s0 = ts_i * pl.tS;
        it.head0 = hg * pl.hpc;
        it.ks_first = max(0, sh.eS - it.s0);
        it.ks_last = min(pl.hS - 1, sh.S - 1 - it.s0 + sh.eS);
        it.pad0 = it.pad1 = it.pad2 = 0;
        sItems[tid] = it;
    }
    __syncthreads();
    struct Cursor { int slot, hd, ks, chunk, H; };
    auto next_head = [&](Cursor& c) {                 // first step of the head after c's
        ++c.H;
        if (++c.hd == pl.hpc) { c.hd = 0; ++c.slot; }
        c.ks = sItems[min(c.slot, my_items - 1)].ks_first;
        c.chunk = chunk_first;
    };
    auto advance = [&](Cursor& c) {
        if (++c.ks > sItems[c.slot].ks_last) {
            c.ks = sItems[c.slot].ks_first;
            if (++c.chunk > chunk_last) next_head(c);
        }
    };
    auto is_head_start = [&](const Cursor& c) { return c.ks == sItems[c.slot].ks_first && c.chunk == chunk_first; };
    int nsteps = 0;
    for (int i = 0; i < my_items; ++i)
        nsteps += (sItems[i].ks_last - sItems[i].ks_first + 1) * (chunk_last - chunk_first + 1) * pl.hpc;
    const int nheads = my_items * pl.hpc;
    const Cursor first = {0, 0, my_items > 0 ? sItems[0].ks_first : 0, chunk_first, 0};

    if (tid == 0) {
        tma_prefetch_desc(&map_a1); tma_prefetch_desc(&map_a2); tma_prefetch_desc(&map_b1); tma_prefetch_desc(&map_b2);
        mbar_init(&bar_a[0], 1);
        mbar_init(&bar_a[1], 1);
        for (int i = 0; i < 3; ++i) mbar_init(&bar_b[i], 1);
        mbar_init(bar_tA, 1);
        mbar_init(bar_tB, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_p[i], 8);           // one arrival per compute warp (256 arrivals on one word serialise)
            mbar_init(&bar_pA[i], 8);
            mbar_init(&bar_acc[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc<512>(tmem_slot);
    if (ncols_pad > ncols) {      // rows TMA never writes must stay finite (they meet zero dS / P columns)
        const int pad_bytes = (ncols_pad - ncols) * G::kRowBytes;
        for (int t = 0; t < nstage * 2 * G::kSlabs; ++t) {
            uint8_t* base = sB + t * blk_slab_bytes + ncols * G::kRowBytes;
            for (int i = tid * 16; i < pad_bytes; i += kWsThreads * 16) *reinterpret_cast<uint4*>(base + i) = make_uint4(0, 0, 0, 0);
        }
    }
    build_mask_tiles(sRm, sCm, pl, sh, h0, w0, tid, kWsThreads);
    fence_proxy_async();                 // generic-proxy writes above -> visible to tcgen05.mma
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns: accumulators | T1 (S) | T2 (dP) | two buffers of bf16-pair columns (dS, resp. P^T)
    const uint32_t tmem_acc1 = tmem_base;                                   // dQ | dV
    const uint32_t tmem_acc2 = tmem_base + D;                               //      dK
    const uint32_t tmem_t1 = tmem_base + (kDKV ? 2 * D : D);
    const uint32_t tmem_t2 = tmem_t1 + ncols_pad;
    const uint32_t tmem_pa = tmem_t2 + ncols_pad;                           // A operand buffers: [2][ncols_pad / 2]
    const int pa_cols = ncols_pad >> 1;

    if (warp >= 8) {
        // =============================== issuing warps (warp-uniform code; instructions on one elected lane) =========
        const bool leader = elect_one();      // each of these warps is converged here
        auto a_buf = [&](int H) { return sA + (pl.rowbuf == 2 ? (H & 1) : 0) * 2 * row_tile_bytes; };
        auto issue_row_load = [&](const Cursor& c) {      // row tiles of c's head
            if (leader) {
                const int hd = c.H;
                const Item it = sItems[c.slot];
                const int s0 = it.s0, b = it.b;
                uint64_t* bar = &bar_a[hd & 1];
                const int cb = (it.head0 + c.hd) * D;
                mbar_expect_tx(bar, 2u * (uint32_t)row_tile_bytes);
#pragma unroll
                for (int sl = 0; sl < G::kSlabs; ++sl) {
                    tma_load_5d(a_buf(hd) + sl * row_slab_bytes, &map_a1, bar, cb + sl * G::kSlabCh, w0, h0, s0, b);
                    tma_load_5d(a_buf(hd) + row_tile_bytes + sl * row_slab_bytes, &map_a2, bar, cb + sl * G::kSlabCh, w0, h0, s0, b);
                }
            }
        };
        auto issue_block_load = [&](int stage, const Cursor& c) {
            if (leader) {
                const Item it = sItems[c.slot];
                const int s0 = it.s0, b = it.b;
                const int cb = (it.head0 + c.hd) * D;
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
        const uint32_t idesc_tA = make_idesc_bf16(nA > 0 ? nA : 16, false, false), idesc_tB = make_idesc_bf16(ncols_pad - nA, false, false);
        const uint32_t idesc_acc = make_idesc_bf16(D, false, true);
        const uint64_t da0 = make_smem_desc(smem_u32(sA), 16, G::kAtomBytes, G::kSwizzleCode);
        const uint64_t dbk0 = make_smem_desc(smem_u32(sB), 16, G::kAtomBytes, G::kSwizzleCode);                         // block, K-major
        const uint64_t dbm0 = make_smem_desc(smem_u32(sB), (uint32_t)blk_slab_bytes, G::kAtomBytes, G::kSwizzleCode);   // block, MN-major
        const uint64_t dds0 = make_smem_desc(smem_u32(sDS), 16, 1024, 2u);
        const uint32_t a_buf_step = (pl.rowbuf == 2) ? (uint32_t)((2 * row_tile_bytes) >> 4) : 0u;
        const uint32_t stage_step = (uint32_t)((2 * blk_tile_bytes) >> 4);
        const uint64_t drm0 = make_smem_desc(smem_u32(sRm), 2048u, 128u, 0u);                       // mask operands: no swizzle
        const uint64_t dcm0 = make_smem_desc(smem_u32(sCm), (uint32_t)ncols_pad * 16u, 128u, 0u);
        const int nk_m = pl.km >> 4;
        // Issue discipline (tools/micro/mma_cost.cu, sync_latency.cu): ONE branch on the elected lane around a whole chain of
        // tcgen05.mma with compile-time operand offsets costs ~20-35 cycles per MMA; a predicate per MMA or a run-time
        // trip count costs ~100.  The accumulation chains are therefore unrolled templates picked by a switch.
        auto issue_t_mma = [&](int stage, int hd, int chunk, int part) {  // T1 = A1 B1^T + mask, T2 = A2 B2^T; columns A (part 0) or B (part 1)
            if (leader) {
                const uint32_t col0 = part ? (uint32_t)nA : 0u;        // columns = rows of the K-major block
                const uint32_t idesc_t = part ? idesc_tB : idesc_tA;
                const uint64_t da_h = da0 + (hd & 1) * a_buf_step;
                const uint64_t db_s = dbk0 + stage * stage_step + ((col0 * (uint32_t)G::kRowBytes) >> 4);
#pragma unroll
                for (int op = 0; op < 2; ++op) {
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk) {
                        const uint32_t sl = (uint32_t)((kk * 16) / G::kSlabCh);
                        const uint32_t koff = (uint32_t)((((kk * 16) % G::kSlabCh) * 2) >> 4);
                        const uint64_t da = da_h + op * (uint32_t)(row_tile_bytes >> 4) + sl * (uint32_t)(row_slab_bytes >> 4) + koff;
                        const uint64_t db = db_s + op * (uint32_t)(blk_tile_bytes >> 4) + sl * (uint32_t)(blk_slab_bytes >> 4) + koff;
                        umma_bf16_ss((op ? tmem_t2 : tmem_t1) + col0, da, db, idesc_t, kk > 0);
                    }
                    if (op == 0) {      // S += R C^T: window / border mask (16 channels = two core-matrix columns per MMA; km is 16 or 32)
                        const uint64_t dcm = dcm0 + (uint32_t)chunk * (uint32_t)(cm_tile_bytes >> 4) + ((col0 * 16u) >> 4);
                        umma_bf16_ss(tmem_t1 + col0, drm0, dcm, idesc_t, 1u);
                        if (nk_m > 1)
                            umma_bf16_ss(tmem_t1 + col0, drm0 + (uint32_t)((2 * 2048) >> 4), dcm + (uint32_t)((2 * ncols_pad * 16) >> 4), idesc_t, 1u);
                    }
                }
                umma_commit(part ? bar_tB : bar_tA);
            }
            __syncwarp();
        };
        const int nk_acc = ncols_pad / 16;
        auto issue_acc_chain = [&](auto nk_tag, uint32_t ta, uint64_t d_ds, uint64_t d_b1, uint64_t d_b2, bool accumulate, uint64_t* bar) {
            constexpr int NK = decltype(nk_tag)::value;
#pragma unroll
            for (int kk = 0; kk < NK; ++kk) {
                const uint32_t acc = (accumulate || kk > 0) ? 1u : 0u;
                const uint32_t boff = (uint32_t)kk * (uint32_t)((16 * G::kRowBytes) >> 4);      // next 16 rows of the halo block
                if constexpr (kDKV) {
                    // dS^T tile: 64-column (128-byte) slabs of 128 rows; 16 columns = 2 sixteen-byte units inside a slab
                    const uint32_t dsoff = (uint32_t)(kk >> 2) * (uint32_t)((128 * 128) >> 4) + (uint32_t)(kk & 3) * 2u;
                    umma_bf16_ts(tmem_acc1, ta + 8 * kk, d_b2 + boff, idesc_acc, acc);          // dV += P^T dO_t
                    umma_bf16_ss(tmem_acc2, d_ds + dsoff, d_b1 + boff, idesc_acc, acc);         // dK += dS^T Q_t
                } else {
                    umma_bf16_ts(tmem_acc1, ta + 8 * kk, d_b1 + boff, idesc_acc, acc);          // dQ += dS K_t
                }
            }
            umma_commit(bar);
        };
        auto issue_acc_mma = [&](int t, int stage, bool accumulate) {
            if (leader) {
                const uint32_t ta = tmem_pa + (t & 1) * pa_cols;                              // dS (dQ kernel) or P^T (dK/dV kernel), from TMEM
                const uint64_t d_ds = dds0 + (t & 1) * (uint32_t)(p_tile_bytes >> 4);         // dS^T (dK/dV kernel), from shared memory
                const uint64_t d_b1 = dbm0 + stage * stage_step;
                const uint64_t d_b2 = d_b1 + (uint32_t)(blk_tile_bytes >> 4);
                uint64_t* bar = &bar_acc[t & 1];
#define WM_ACC_CASE(n) case n: issue_acc_chain(std::integral_constant<int, n>{}, ta, d_ds, d_b1, d_b2, accumulate, bar); break;
                switch (nk_acc) {
                    WM_ACC_CASE(1) WM_ACC_CASE(2) WM_ACC_CASE(3) WM_ACC_CASE(4) WM_ACC_CASE(5) WM_ACC_CASE(6) WM_ACC_CASE(7) WM_ACC_CASE(8)
                    WM_ACC_CASE(9) WM_ACC_CASE(10) WM_ACC_CASE(11) WM_ACC_CASE(12) WM_ACC_CASE(13) WM_ACC_CASE(14) WM_ACC_CASE(15)
                    default: issue_acc_chain(std::integral_constant<int, 16>{}, ta, d_ds, d_b1, d_b2, accumulate, bar); break;
                }
#undef WM_ACC_CASE
            }
            __syncwarp();
        };
        uint32_t dbg_flag = (blockIdx.x == 1) && leader;
        asm volatile("" : "+r"(dbg_flag));
        const bool dbg_on = dbg_flag != 0;
        (void)dbg_on;
        Cursor cur = first;

        if (warp == 10) {
            // ---- TMA loader: row tiles one head ahead, halo blocks nstage steps ahead -------------------------------
            Cursor ld = cur;
            int ld_t = 0, ld_stage = 0;
            issue_row_load(cur);
            for (; ld_t < nstage && ld_t < nsteps; ++ld_t) {
                issue_block_load(ld_stage, ld);
                advance(ld);
                if (++ld_stage == nstage) ld_stage = 0;
            }
            for (int t = 0; t < nsteps; ++t) {
                const bool head_start = is_head_start(cur);
                if (head_start && pl.rowbuf == 2 && cur.H + 1 < nheads) {
                    // two row buffers: the next head's tiles go into the buffer of head H-1, whose last (S, dP) was drained
                    // when every thread arrived for step t-1
                    if (t > 0) mbar_wait(&bar_p[(t - 1) & 1], ((t - 1) >> 1) & 1);
                    Cursor nh = cur;
                    next_head(nh);
                    issue_row_load(nh);
                } else if (head_start && pl.rowbuf == 1 && t > 0) {
                    // one row buffer: this head's tiles, once the previous head's last (S, dP) has been drained
                    mbar_wait(&bar_p[(t - 1) & 1], ((t - 1) >> 1) & 1);
                    issue_row_load(cur);
                }
                if (t >= 1 && ld_t < nsteps) {   // the stage of step t-1 is free once its accumulating MMAs have retired
                    mbar_wait(&bar_acc[(t - 1) & 1], ((t - 1) >> 1) & 1);
                    issue_block_load(ld_stage, ld);
                    advance(ld);
                    ++ld_t;
                    if (++ld_stage == nstage) ld_stage = 0;
                }
                advance(cur);
            }
        } else if (warp == 8) {
            // ---- (S, dP) of step t+1, half A after the threads leave half A of step t, half B likewise ----------------
            Cursor nxt = cur;
            advance(nxt);
            mbar_wait(&bar_a[0], 0);
            mbar_wait(&bar_b[0], 0);
            tc_fence_after();
            if (gA > 0) issue_t_mma(0, 0, cur.chunk, 0);
            issue_t_mma(0, 0, cur.chunk, 1);
            int st_nxt = (nstage > 1) ? 1 : 0;
            uint32_t b_par = 1u;                         // bit s = parity of stage s's next completion (stage 0 consumed once)
            for (int t = 0; t + 1 < nsteps; ++t) {
                DBGW(0);
                if (gA > 0) {
                    mbar_wait(&bar_pA[t & 1], (t >> 1) & 1);
                    DBGW(1);
                }
                mbar_wait(&bar_b[st_nxt], (b_par >> st_nxt) & 1u);
                b_par ^= 1u << st_nxt;
                if (nxt.H != cur.H) mbar_wait(&bar_a[nxt.H & 1], (nxt.H >> 1) & 1);
                tc_fence_after();
                if (gA > 0) issue_t_mma(st_nxt, nxt.H, nxt.chunk, 0);
                DBGW(2);
                mbar_wait(&bar_p[t & 1], (t >> 1) & 1);
                tc_fence_after();
                issue_t_mma(st_nxt, nxt.H, nxt.chunk, 1);
                DBGW(3);
                cur = nxt;
                advance(nxt);
                if (++st_nxt == nstage) st_nxt = 0;
            }
        } else {
            // ---- accumulating MMAs of step t, once all of its dS / P is written -------------------------------------
            int st_cur = 0;
            for (int t = 0; t < nsteps; ++t) {
                const bool head_start = is_head_start(cur);
                mbar_wait(&bar_p[t & 1], (t >> 1) & 1);
                tc_fence_after();
                DBGW(4);
                issue_acc_mma(t, st_cur, !head_start);
                DBGW(5);
                advance(cur);
                if (++st_cur == nstage) st_cur = 0;
            }
        }
    } else {
        // =============================== compute warps ==================================================
        const int quad = warp & 3, half = warp >> 2;
        const int row = quad * 32 + lane;
        const int ctid = tid;                                            // 0..255
        const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
        const int plane_mask = (1 << pl.lgPlane) - 1;
        const int rs = row >> pl.lgPlane, rh = (row & plane_mask) >> pl.lgTW, rw = row & (pl.tW - 1);
        const bool hw_valid = (h0 + rh < sh.H) && (w0 + rw < sh.W);
        // (item-dependent) is this row a token of the grid, and which
        auto valid_of = [&](const Cursor& c) { return hw_valid && (sItems[c.slot].s0 + rs < sh.S); };
        auto tok_of = [&](const Cursor& c) {
            return (((long)sItems[c.slot].b * sh.S + (sItems[c.slot].s0 + rs)) * sh.H + (h0 + rh)) * sh.W + (w0 + rw);
        };
        const int w_rs = (quad * 32) >> pl.lgPlane;
        const int w_rh_lo = ((quad * 32) & plane_mask) >> pl.lgTW, w_rh_hi = ((quad * 32 + 31) & plane_mask) >> pl.lgTW;
        constexpr float kLog2e = 1.4426950408889634f;

        auto load_colvec = [&](const Cursor& c, float& lse2, float& dl) {     // dK/dV kernel: one halo column per thread
            const Item it = sItems[c.slot];
            const int gs = it.s0 - sh.eS + c.ks;
            const int col = ctid;
            lse2 = 0.f;
            dl = 0.f;
            if (col < ncols) {
                const int khl = col / pl.hW, kw = col - khl * pl.hW;
                const int gh = h0 - sh.eH + c.chunk * pl.ch + khl, gw = w0 - sh.eW + kw;
                if (gs >= 0 && gs < sh.S && gh >= 0 && gh < sh.H && gw >= 0 && gw < sh.W) {
                    const long idx = ((((long)it.b * sh.S + gs) * sh.H + gh) * sh.W + gw) * sh.heads + it.head0 + c.hd;
                    lse2 = __ldg(prm.lse + idx);          // raw: scaled in store_colvec, a whole step later, so that
                    dl = __ldg(prm.delta + idx);          // the load latency is never waited for at the top of a step
                }
            }
        };
        auto store_colvec = [&](int buf, float lse_raw, float dl_raw) {
            if (ctid < ncols_pad) {
                sCol[(buf * 2 + 0) * ncols_pad + ctid] = -lse_raw * kLog2e;
                sCol[(buf * 2 + 1) * ncols_pad + ctid] = -dl_raw * sh.scale;
            }
        };
        auto finish_head = [&](const Cursor& c) {   // accumulators of c's head -> bf16 -> global
            const bool row_valid = valid_of(c);
            const long row_off = tok_of(c) * prm.ld_out + (sItems[c.slot].head0 + c.hd) * D + half * (D / 2);
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

        Cursor cur = first;
        Cursor nxt = cur;
        advance(nxt);
        Cursor done_head = cur;                                          // the head whose accumulators are drained next
        float row_lse2 = 0.f, row_delta = 0.f;
        float ahead_lse = 0.f, ahead_delta = 0.f;                        // dQ kernel: raw values of the next head, in flight
        if constexpr (!kDKV) {
            if (valid_of(cur)) {
                ahead_lse = __ldg(prm.lse + tok_of(cur) * sh.heads + sItems[0].head0);
                ahead_delta = __ldg(prm.delta + tok_of(cur) * sh.heads + sItems[0].head0);
            }
        }
        if constexpr (kDKV) {
            float a, c;
            load_colvec(cur, a, c);
            store_colvec(0, a, c);
            asm volatile("bar.sync 5, 256;" ::: "memory");               // the 256 compute threads only
        }
        bool p_zero[2] = {false, false};
        int mask_chunk = -1;
        int g_lo = 0, g_hi = 0;
        int cbuf = 0;                      // dK/dV kernel: lse / delta column buffer of the current step (t mod 3)
        bool tA_seen = false;              // columns A of this step were already seen complete by last step's probe
        bool chunk_live = false;
        const int ngroups = ncols_pad >> 4;
        const uint32_t zero8[8] = {0, 0, 0, 0, 0, 0, 0, 0};

        uint32_t dbg_flag = (blockIdx.x == 1 && tid == 0);
        asm volatile("" : "+r"(dbg_flag));
        const bool dbg_on = dbg_flag != 0;
        (void)dbg_on;
        for (int t = 0; t < nsteps; ++t) {
            DBGW(8);
            const int buf = t & 1;
            const bool head_start = is_head_start(cur);
            float nxt_lse2 = 0.f, nxt_dl = 0.f;
            if constexpr (kDKV) {
                if (t + 1 < nsteps) load_colvec(nxt, nxt_lse2, nxt_dl);      // global loads in flight during the waits
            } else {
                if (head_start) {
                    asm volatile("" : "+f"(ahead_lse), "+f"(ahead_delta));   // first use of the loads issued one head ago
                    row_lse2 = -ahead_lse * kLog2e;          // kept negated: the addends of the two FFMA2 below
                    row_delta = -ahead_delta * sh.scale;
                    if (cur.H + 1 < nheads) {                // the next head may belong to the next work item
                        Cursor nh = cur;
                        next_head(nh);
                        ahead_lse = ahead_delta = 0.f;
                        if (valid_of(nh)) {
                            ahead_lse = __ldg(prm.lse + tok_of(nh) * sh.heads + sItems[nh.slot].head0 + nh.hd);
                            ahead_delta = __ldg(prm.delta + tok_of(nh) * sh.heads + sItems[nh.slot].head0 + nh.hd);
                        }
                    }
                }
            }
            const int kh0 = cur.chunk * pl.ch;
            if (cur.chunk != mask_chunk) {               // live column range of this quadrant for this h-chunk
                mask_chunk = cur.chunk;
                // its rows see halo rows [w_rh_lo, w_rh_hi + 2 eH]; halo rows outside the grid are never live
                const int ua = max(max(w_rh_lo, kh0), khg_lo), ub = min(min(w_rh_hi + 2 * sh.eH, kh0 + pl.ch - 1), khg_hi);
                chunk_live = ub >= ua;
                g_lo = ((ua - kh0) * pl.hW) >> 4;
                g_hi = min(((ub - kh0 + 1) * pl.hW + 15) >> 4, ngroups);
            }
            const bool live = chunk_live && (cur.ks >= w_rs) && (cur.ks <= w_rs + 2 * sh.eS);
            // The dS / P buffers of parity `buf` are free once the accumulating MMAs of step t-2 have retired.
            if (t >= 2) mbar_wait(&bar_acc[buf], ((t - 2) >> 1) & 1);
            DBGW(9);

            const uint32_t tmem_a = tmem_pa + buf * pa_cols + lane_sel;       // this row's bf16-pair columns
            uint8_t* ds_tile = sDS + buf * p_tile_bytes;
            auto store_ds_smem = [&](int g, const uint32_t (&pk)[8]) {        // dK/dV kernel: 16 bf16 -> swizzled smem tile
                const uint32_t slab_off = (g >> 2) * (128 * 128);
                const int c16 = (g & 3) * 2;
                *reinterpret_cast<uint4*>(ds_tile + slab_off + sw128_offset(row, c16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(ds_tile + slab_off + sw128_offset(row, c16 + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            };
            auto store_tmem = [&](int g, const uint32_t (&pk)[8]) {           // 16 bf16 -> 8 pair columns of the TMEM A operand
                tmem_st8(tmem_a + g * 8, pk);
            };
            auto zero_groups = [&](int za, int zb) {
                for (int g = za; g < zb; ++g) {
                    store_tmem(g, zero8);
                    if constexpr (kDKV) store_ds_smem(g, zero8);
                }
            };
            const float* col_lse2 = sCol + (cbuf * 2 + 0) * ncols_pad;
            const float* col_dl = sCol + (cbuf * 2 + 1) * ncols_pad;
            const uint64_t cc2 = pk2(pl.scale_log2, pl.scale_log2), sc2 = pk2(sh.scale, sh.scale);
            const uint64_t rl2 = pk2(row_lse2, row_lse2), rd2 = pk2(row_delta, row_delta);
            // 16 columns of this row: P = 2^(S c - lse2), dS = P (dP scale - delta scale).  Masked scores are <= -2^60
            // (mask operand of the S MMA), so their P and dS are exactly 0.  nl / nd: NEGATED lse2 and delta*scale.
            auto math_group = [&](int g) {
                uint32_t s[16], dp[16];
                tmem_ld16(tmem_t1 + lane_sel + g * 16, s);
                tmem_ld16(tmem_t2 + lane_sel + g * 16, dp);
                uint64_t nl[8], nd[8];
                if constexpr (kDKV) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(col_lse2 + g * 16 + 4 * i);
                        const ulonglong2 c = *reinterpret_cast<const ulonglong2*>(col_dl + g * 16 + 4 * i);
                        nl[2 * i] = a.x; nl[2 * i + 1] = a.y;
                        nd[2 * i] = c.x; nd[2 * i + 1] = c.y;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { nl[i] = rl2; nd[i] = rd2; }
                }
                tmem_wait_ld();
                uint32_t pkp[8], pkd[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint64_t x = ffma2(pk2u(s[2 * i], s[2 * i + 1]), cc2, nl[i]);
                    uint64_t pr;
                    if (i < WM_BWD_POLY) {
                        pr = exp2_poly2(x);
                    } else {
                        float x0, x1;
                        upk2(x, x0, x1);
                        pr = pk2(ex2(x0), ex2(x1));
                    }
                    const uint64_t y = ffma2(pk2u(dp[2 * i], dp[2 * i + 1]), sc2, nd[i]);
                    pkd[i] = pack_bf16_2(fmul2(pr, y));
                    if constexpr (kDKV) pkp[i] = pack_bf16_2(pr);
                }
                if constexpr (kDKV) {
                    store_ds_smem(g, pkd);
                    store_tmem(g, pkp);
                } else {
                    store_tmem(g, pkd);
                }
            };
            // this warp's live groups inside [x0, x1), split between the two threads of a row; dead groups are zeroed.
            // The barrier needed next is probed before the last group, so that the probe's latency hides under its math.
            auto run_half = [&](int x0, int x1, int odd_to, uint64_t* probe_bar, uint32_t probe_parity) -> bool {
                const int l0 = min(max(g_lo, x0), x1), l1 = max(min(g_hi, x1), l0);
                const int mid = (l0 + l1 + (odd_to == 0 ? 1 : 0)) >> 1;
                const int ga = half ? mid : l0, gb = half ? l1 : mid;
                bool probed = false;
                for (int g = ga; g < gb - 1; ++g) math_group(g);
                if (WM_WS_PROBE) probed = mbar_test(probe_bar, probe_parity);
                if (gb > ga) math_group(gb - 1);
                zero_groups(half ? l1 : x0, half ? x1 : l0);
                return __all_sync(0xffffffffu, probed);
            };
            auto zero_half = [&](int x0, int x1) {
                const int mid = (x0 + x1) >> 1;
                zero_groups(half ? mid : x0, half ? x1 : mid);
            };
            bool tB_seen = false;
            // Every warp waits for both halves of (S, dP), also when the block is dead for it: the parity waits below are
            // only unambiguous for a thread that has seen every phase of the barrier, and nothing else orders a dead warp
            // behind the (S, dP) MMAs now that they and the accumulating MMAs are issued by different warps.
            if (gA > 0) {
                if (!tA_seen) mbar_wait(bar_tA, t & 1);       // columns A of (S, dP) of this step
                if (live) {
                    tc_fence_after();
                    DBGW(10);
                    tB_seen = run_half(0, gA, 0, bar_tB, t & 1);
                } else if (!p_zero[buf]) {
                    zero_half(0, gA);
                }
                if constexpr (kDKV) {
                    // Next step's lse / delta columns, written BEFORE the arrival below: columns A of step t+1 are issued
                    // after it, so whoever sees them also sees these stores.  Three buffers: the one written here was
                    // last read in step t-2, which bar_acc(t-2) has put behind every thread.
                    // (ptxas otherwise hoists the scaling multiplies up to the loads at the top of the step.)
                    asm volatile("" : "+f"(nxt_lse2), "+f"(nxt_dl));
                    if (t + 1 < nsteps) store_colvec(cbuf == 2 ? 0 : cbuf + 1, nxt_lse2, nxt_dl);
                }
                DBGW(13);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_pA[buf]);
            }
            tA_seen = false;
            if (!tB_seen) mbar_wait(bar_tB, t & 1);       // columns B
            DBGW(14);
            if (live) {
                tc_fence_after();
                if (gA == 0) DBGW(10);
                tA_seen = run_half(gA, ngroups, 1, bar_tA, (t + 1) & 1) && gA > 0 && t + 1 < nsteps;
                p_zero[buf] = false;
            } else if (!p_zero[buf]) {
                zero_half(gA, ngroups);
                p_zero[buf] = true;
            }
            if constexpr (kDKV) {
                if (gA == 0) {
                    asm volatile("" : "+f"(nxt_lse2), "+f"(nxt_dl));
                    if (t + 1 < nsteps) store_colvec(cbuf == 2 ? 0 : cbuf + 1, nxt_lse2, nxt_dl);
                }
                fence_proxy_async();              // dS^T (generic proxy) -> visible to tcgen05.mma
            }
            DBGW(11);
            if (head_start && t > 0) {
                // One accumulator set: the previous head must be drained before this head's first accumulating MMA, which
                // the driver issues only after every compute thread has arrived below.  Doing it here, after this step's
                // math, lets the wait for the previous head's last MMAs overlap that math.
                mbar_wait(&bar_acc[(t - 1) & 1], ((t - 1) >> 1) & 1);
                tc_fence_after();
                finish_head(done_head);
                done_head = cur;
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_p[buf]);
            DBGW(12);
            cur = nxt;
            advance(nxt);
            cbuf = (cbuf == 2) ? 0 : cbuf + 1;
        }
        mbar_wait(&bar_acc[(nsteps - 1) & 1], ((nsteps - 1) >> 1) & 1);
        tc_fence_after();
        finish_head(done_head);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<512>(tmem_base);
}

template <int D, int MODE>
static int launch_ws(const void* a1, const void* a2, const void* b1, const void* b2, const long (&ld)[4], const float* lse,
                     const float* delta, void* out1, void* out2, long ld_out, const AttnShape& s, const Plan& pl,
                     cudaStream_t st) {
    using G = Geo<D>;
    CUtensorMap ma1, ma2, mb1, mb2;
    const int C = s.inner();
    if (int rc = make_tensor_map_5d(&ma1, a1, s.B, s.S, s.H, s.W, C, ld[0], G::kSlabCh, pl.tW, pl.tH, pl.tS, G::kSwizzleBytes)) return rc;
    if (int rc = make_tensor_map_5d(&ma2, a2, s.B, s.S, s.H, s.W, C, ld[1], G::kSlabCh, pl.tW, pl.tH, pl.tS, G::kSwizzleBytes)) return rc;
    if (int rc = make_tensor_map_5d(&mb1, b1, s.B, s.S, s.H, s.W, C, ld[2], G::kSlabCh, pl.hW, pl.ch, 1, G::kSwizzleBytes)) return rc;
    if (int rc = make_tensor_map_5d(&mb2, b2, s.B, s.S, s.H, s.W, C, ld[3], G::kSlabCh, pl.hW, pl.ch, 1, G::kSwizzleBytes)) return rc;
    BwdWsParams prm{s, pl, lse, delta, static_cast<__nv_bfloat16*>(out1), static_cast<__nv_bfloat16*>(out2), ld_out};
    WM_CUDA_CHECK(cudaFuncSetAttribute(l3d_bwd_ws_kernel<D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem_bytes));
    // Grid (see the kernel's work decomposition): one persistent CTA per SM when there is more than one work item per SM
    // and its items fit the per-CTA table, else one CTA per item.  Measured on the same box against one CTA per item:
    // config 3 dQ 276 -> 246 us, dK/dV 382 -> 349; config 4 dQ 362 -> 309, dK/dV 547 -> 514.
    const long items = (long)pl.tilesW * pl.tilesH * pl.tilesS * s.B * (s.heads / pl.hpc);
    if (items > 0x7fffffffL) return fail(WM_EUNSUPPORTED, "grid too large for the tensor-core kernel");
    const bool fits = (items + sm_count() - 1) / sm_count() + pl.tilesH * pl.tilesW <= kMaxItemsPerCta;      // per-CTA item table
    const bool every_position_served = pl.tilesH * pl.tilesW <= sm_count();       // a CTA serves ONE brick position
    const unsigned grid = (unsigned)(items > (long)sm_count() && fits && every_position_served ? (long)sm_count() : items);
    l3d_bwd_ws_kernel<D, MODE><<<grid, kWsThreads, pl.smem_bytes, st>>>(ma1, ma2, mb1, mb2, prm);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}

#if WM_EXPERIMENT == 7
extern "C" __attribute__((visibility("default"))) int wm_debug_read_ws(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, g_dbg_ws, sizeof(long long) * 2 * 64 * 16);
}
#endif

template <int D>
static int launch_bwd_d(const void* q, const void* k, const void* v, const void* o, const float* lse, const void* dout,
                        void* dq, void* dk, void* dv, float* delta, const AttnShape& s, cudaStream_t st) {
    Plan pq, pkv;
    if (!make_plan(s, kBwdDQws, pq) || !make_plan(s, kBwdDKVws, pkv))
        return fail(WM_EUNSUPPORTED, "no tensor-core backward tiling for this shape");
    const long items = s.tokens() * s.heads;
    l3d_delta_kernel<<<(unsigned)((items * (s.d / 8) + 255) / 256), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(o),
                                                                      static_cast<const __nv_bfloat16*>(dout), delta,
                                                                      items, s.d);
    WM_CUDA_CHECK(cudaGetLastError());
    const long in = s.inner(), lq = s.q_ld(), lkv = s.kv_ld();
    const long ld_dq[4] = {lq, in, lkv, lkv}, ld_dkv[4] = {lkv, lkv, lq, in};
    if (int rc = launch_ws<D, kBwdDQws>(q, dout, k, v, ld_dq, lse, delta, dq, nullptr, lq, s, pq, st)) return rc;     // rows: queries
    return launch_ws<D, kBwdDKVws>(k, v, q, dout, ld_dkv, lse, delta, dv, dk, lkv, s, pkv, st);                       // rows: keys
}

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
