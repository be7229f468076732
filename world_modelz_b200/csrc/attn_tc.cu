// Local windowed 3D attention on the 5th-generation tensor cores (bf16, fp32 accumulate).
//
// Tiling.  A CTA owns a query tile of 128 tokens, a tS x tH x tW brick of the (S,H,W)
// grid, for one (batch, head).  The keys any of its queries can see form the brick's halo
// (tS+2eS) x (tH+2eH) x (tW+2eW).  The halo is walked in BLOCKS: one s-plane of the halo,
// restricted to `ch` consecutive h-rows, i.e. ch*hW keys (<= 256).  K and V blocks are
// fetched with ONE 5-D TMA box each straight out of the [B,S,H,W,C] activation tensor:
// the box may start at negative coordinates or run past the grid, and TMA zero-fills
// those keys -- the reference's F.pad (local_3d_attention.py:57-63) for free.
//
// Per block (flash-attention style, accumulators in TMEM):
//   S = Q K^T          tcgen05.mma, A = Q tile (K-major), B = K block (K-major), N = block
//   softmax            one thread per query row: tcgen05.ld the row, mask by window /
//                      grid-border geometry, running max / sum, P in bf16 -> smem
//   O += P V           tcgen05.mma, A = P (K-major), B = V block (MN-major), N = dim_head
// Masking (reference :92-94) is a per-thread column bitmask built from coordinates: a
// key column is live iff it lies inside this query's window AND inside the grid.  Warps
// skip whole planes / 16-column groups that none of their 32 queries can see.
//
// Replaces Local3dAttention.local_attention (local_3d_attention.py:78-99).
#include "attn_tc.cuh"

// -DWM_EXPERIMENT=7 compiles the clock64 timeline instrumentation in (tools/build_timeline_lib.sh, tools/dbg_timeline.py)
#ifndef WM_EXPERIMENT
#define WM_EXPERIMENT 0
#endif

#include <math.h>
#include <stdlib.h>
#include <mutex>

namespace wm {
namespace tc {

// --------------------------------------------------------------------- host: tensor map
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int make_tensor_map_5d(CUtensorMap* out, const void* base, int B, int S, int H, int W, int C, int box_c, int box_w,
                       int box_h, int box_s, int swizzle_bytes) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr) return fail(WM_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)S, (cuuint64_t)B};
    const cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H,
                                   (cuuint64_t)C * 2 * W * H * S};
    const cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_s, 1u};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(WM_ECUDA, "cuTensorMapEncodeTiled failed (%d) for box (%d,%d,%d,%d) of [%d,%d,%d,%d,%d]", (int)r,
                    box_c, box_w, box_h, box_s, B, S, H, W, C);
    return WM_OK;
}

// 2-D fp32 tensor map: [rows, cols] with an arbitrary row stride; box = (32 floats = 128 B, box_rows), 128B swizzle
int make_tensor_map_2d_f32(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes,
                           uint32_t box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr) return fail(WM_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)row_stride_bytes};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(WM_ECUDA, "cuTensorMapEncodeTiled (2-D fp32) failed (%d)", (int)r);
    return WM_OK;
}

// ------------------------------------------------------------------------------- plan
static int row_bytes_of(int d) { return d == 32 ? 64 : 128; }
static int slabs_of(int d) { return d == 128 ? 2 : 1; }
static const size_t kFixedSmem = 1024 /*alignment slack*/ + 3 * 2 * 4 * 128 * 4 /*exchange*/ + 256 /*barriers*/;

static size_t mask_tile_bytes(int ncols_pad, int km, int nchunk) {      // row tile [128 x km] + nchunk column tiles, bf16
    return (size_t)128 * km * 2 + (size_t)nchunk * ncols_pad * km * 2;
}

size_t smem_bytes_for(Mode mode, int d, int ncols_pad, int nstage, int rowbuf, int km, int nchunk) {
    const size_t row_tile = (size_t)slabs_of(d) * 128 * row_bytes_of(d);
    const size_t blk = (size_t)slabs_of(d) * ncols_pad * row_bytes_of(d);
    const size_t ptile = (size_t)round_up(ncols_pad, 64) / 64 * 128 * 128;
    const size_t fixed = kFixedSmem + mask_tile_bytes(ncols_pad, km, nchunk);
    switch (mode) {
        case kFwd: return fixed + rowbuf * row_tile + nstage * 2 * blk;                   // Q | (K,V) stages (P lives in TMEM)
        case kBwdDQws: return fixed + rowbuf * 2 * row_tile + nstage * 2 * blk;           // Q,dO | (K,V) stages (dS lives in TMEM)
        default: return fixed + rowbuf * 2 * row_tile + nstage * 2 * blk + 2 * ptile + 6 * ncols_pad * 4;   // K,V | (Q,dO) | dS^T x2 | lse,delta x3 (P^T in TMEM)
    }
}

int tmem_cols_for(Mode mode, int d, int ncols_pad) {
    switch (mode) {
        case kFwd: return d + 3 * ncols_pad;             // O (x2 if it fits) | S x2 | P x2 (bf16, half the columns)
        case kBwdDQws: return d + 3 * ncols_pad;         // dQ | S | dP | dS x2 (bf16 pairs)
        default: return 2 * d + 3 * ncols_pad;           // dV | dK | S^T | dP^T | P^T x2 (bf16 pairs)
    }
}

// Choose the brick and block shape: minimise the dense MMA columns executed per clip.
bool make_plan(const AttnShape& s, Mode mode, Plan& best) {
    if (s.d != 32 && s.d != 64 && s.d != 128) return false;
    static const int bricks[][3] = {{2, 8, 8}, {4, 4, 8}, {4, 8, 4}, {1, 8, 16}, {1, 16, 8}, {2, 4, 16}, {2, 16, 4}};
    double best_cost = 1e300;
    bool found = false;
    const int max_cols = 256;
    for (const auto& b : bricks) {
        Plan p{};
        p.tS = b[0]; p.tH = b[1]; p.tW = b[2];
        if (p.tS * p.tH * p.tW != 128 || (p.tH * p.tW) % 32 != 0) continue;
        p.hS = p.tS + 2 * s.eS; p.hH = p.tH + 2 * s.eH; p.hW = p.tW + 2 * s.eW;
        if (p.hW > 32 || p.hH > 256) continue;
        p.km = round_up(p.tH + p.tW, 16);
        p.tilesS = (s.S + p.tS - 1) / p.tS; p.tilesH = (s.H + p.tH - 1) / p.tH; p.tilesW = (s.W + p.tW - 1) / p.tW;
        const double tiles = (double)p.tilesS * p.tilesH * p.tilesW;
        // heads per CTA (forward kernel): walk as many heads as possible while keeping the grid >= 4 CTAs per SM
        int hpc = 1;
        for (int h = s.heads; h >= 1; --h)
            if (s.heads % h == 0 && (double)s.B * tiles * (s.heads / h) >= 4.0 * 148) { hpc = h; break; }
        for (int nchunk = 1; nchunk <= p.hH; ++nchunk) {
            p.ch = (p.hH + nchunk - 1) / nchunk;
            p.nchunk = (p.hH + p.ch - 1) / p.ch;
            p.ncols = p.ch * p.hW;
            p.ncols_pad = round_up(p.ncols, 16);
            if (p.ncols_pad > max_cols || tmem_cols_for(mode, s.d, p.ncols_pad) > 512) continue;
            // shared-memory variants, best first
            bool fits = false;
            const int opts[7][3] = {{4, 2, hpc}, {3, 2, hpc}, {2, 2, hpc}, {4, 1, 1}, {3, 1, 1}, {2, 1, hpc}, {2, 1, 1}};    // {nstage, rowbuf, hpc}
            for (const auto& o : opts) {
                if (mode != kFwd && o[0] > 3) continue;                                // 4 stages: forward kernel only
                if (mode != kFwd && o[1] == 1 && o[2] != 1) continue;                  // bwd: a head loop only with 2 row buffers
                if (o[1] == 2 && hpc == 1) continue;
                if (smem_bytes_for(mode, s.d, p.ncols_pad, o[0], o[1], p.km, p.nchunk) <= (size_t)kSmemLimit) {
                    p.nstage = o[0]; p.rowbuf = o[1]; p.hpc = o[2];
                    fits = true;
                    break;
                }
            }
            if (!fits) continue;
            const double cost = tiles * p.hS * p.nchunk * (p.ncols_pad + 24.0 /*per-block overhead*/);
            if (cost < best_cost) {
                best_cost = cost;
                p.smem_bytes = (int)smem_bytes_for(mode, s.d, p.ncols_pad, p.nstage, p.rowbuf, p.km, p.nchunk);
                // forward: one O accumulator per head parity when TMEM has room (the previous head is drained off the
                // critical path)
                p.obufs = (mode == kFwd && 2 * s.d + 3 * p.ncols_pad <= 512) ? 2 : 1;
                p.tmem_cols = next_pow2(tmem_cols_for(mode, s.d, p.ncols_pad) + (p.obufs == 2 ? s.d : 0));
                p.scale_log2 = s.scale * 1.4426950408889634f;
                p.lgTW = 0; while ((1 << p.lgTW) < p.tW) ++p.lgTW;
                p.lgPlane = 0; while ((1 << p.lgPlane) < p.tH * p.tW) ++p.lgPlane;
                best = p;
                found = true;
            }
            break;   // the smallest feasible nchunk for this brick
        }
    }
    return found;
}

#if WM_EXPERIMENT == 7
__device__ long long g_dbg[64 * 16];
#define DBG(slot) do { if (dbg_on && t < 64) g_dbg[t * 16 + (slot)] = clock64(); } while (0)
#else
#define DBG(slot) do { } while (0)
#endif

// ------------------------------------------------------------------------------ kernel
struct FwdParams {
    AttnShape sh;
    Plan pl;
    __nv_bfloat16* o;
    float* lse;
};

#ifndef WM_FWD_POLY
#define WM_FWD_POLY 3      // of every 8 column pairs, this many take the FMA-pipe exp2 (the rest go to the MUFU)
#endif

// Warp-specialised forward kernel.
//   warps 0-7  compute: warps w and w+4 share TMEM lane quadrant w&3 (32 query rows) and split
//              the quadrant's live key columns ("part" 0 / 1): exponentials, P -> TMEM, epilogue.
//   warp  8    driver (one elected lane): TMA loads of Q / K / V blocks and every tcgen05.mma.
// S lives in two TMEM buffers and P in two more, so S_{t+1} (and S_{t+2} with four K/V stages) is
// computed while the compute warps are still in step t, and O += P_t V_t runs while they are already in
// step t+1.  All hand-offs are mbarriers; there is no CTA-wide barrier inside the loop.  A CTA walks
// `hpc` heads of its brick back to back.  The window / border mask is part of the score MMA
// (build_mask_tiles in attn_tc.cuh), so the element loop is: scale, 2^x, sum, pack.
template <int D>
__global__ void __launch_bounds__(kFwdThreads, 1)
l3d_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv_k,
                  const __grid_constant__ CUtensorMap map_kv_v, const FwdParams prm) {
    using G = Geo<D>;
    constexpr int NPART = 2;
    const AttnShape& sh = prm.sh;
    const Plan& pl = prm.pl;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays in the shared window

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ncols = pl.ncols, ncols_pad = pl.ncols_pad;
    const int nstage = pl.nstage;
    const int q_slab_bytes = 128 * G::kRowBytes;
    const int q_tile_bytes = G::kSlabs * q_slab_bytes;
    const int kv_slab_bytes = ncols_pad * G::kRowBytes;
    const int kv_tile_bytes = G::kSlabs * kv_slab_bytes;
    const int cm_tile_bytes = ncols_pad * pl.km * 2;

    uint8_t* sQ = smem;                                          // [rowbuf][slabs][128 rows]
    uint8_t* sKV = sQ + pl.rowbuf * q_tile_bytes;                // [nstage][K|V][slabs][ncols_pad rows]
    constexpr int kDriverWarp = 4 * NPART;
    uint8_t* sRm = sKV + nstage * 2 * kv_tile_bytes;             // mask operand of the rows:    [128 x km] bf16
    uint8_t* sCm = sRm + 128 * pl.km * 2;                        // mask operand of the columns: [nchunk][ncols_pad x km] bf16
    float* sX = reinterpret_cast<float*>(sCm + pl.nchunk * cm_tile_bytes);     // [3 uses][2 parities][4][128 rows]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sX + 3 * 2 * 4 * 128);
    uint64_t* bar_q = bars;           // [2]  Q tile of a head landed
    uint64_t* bar_kv = bars + 2;      // [4]  K/V block landed
    uint64_t* bar_s = bars + 6;       // [2]  S buffer computed            (tcgen05.commit)
    uint64_t* bar_p = bars + 8;       // [2]  P buffer written, S buffer drained (8 compute warps)
    uint64_t* bar_o = bars + 10;      // [2]  O += P V of a step retired   (tcgen05.commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    // ---- which brick / head group ----------------------------------------------------------
    const int tw_i = blockIdx.x % pl.tilesW, th_i = blockIdx.x / pl.tilesW;
    const int ts_i = blockIdx.y;
    const int hgroups = sh.heads / pl.hpc;
    const int hg = blockIdx.z % hgroups, b = blockIdx.z / hgroups;
    const int head0 = hg * pl.hpc;
    const int s0 = ts_i * pl.tS, h0 = th_i * pl.tH, w0 = tw_i * pl.tW;

    // block iteration space: planes and h-chunks that intersect the grid
    const int ks_first = max(0, sh.eS - s0), ks_last = min(pl.hS - 1, sh.S - 1 - s0 + sh.eS);
    const int khg_lo = max(0, sh.eH - h0), khg_hi = min(pl.hH - 1, sh.H - 1 - h0 + sh.eH);
    int chunk_first = 0, chunk_last = 0;
    for (int c = 0; c < pl.nchunk; ++c) {           // nchunk is tiny; avoids integer divisions
        if (khg_lo >= (c + 1) * pl.ch) chunk_first = c + 1;
        if (khg_hi >= c * pl.ch) chunk_last = c;
    }
    const int nplanes = ks_last - ks_first + 1;
    const int nblocks = nplanes * (chunk_last - chunk_first + 1);
    const int nsteps = nblocks * pl.hpc;

    // ---- one-time setup ----------------------------------------------------------------------
    if (tid == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_kv_k);
        tma_prefetch_desc(&map_kv_v);
        mbar_init(&bar_q[0], 1);
        mbar_init(&bar_q[1], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&bar_kv[i], 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_s[i], 1);
            mbar_init(&bar_p[i], 4 * NPART);       // one arrival per compute warp (hundreds of arrivals on one word serialise)
            mbar_init(&bar_o[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == kDriverWarp) {
        // power-of-two TMEM allocation
        if (pl.tmem_cols <= 128) tmem_alloc<128>(tmem_slot);
        else if (pl.tmem_cols <= 256) tmem_alloc<256>(tmem_slot);
        else tmem_alloc<512>(tmem_slot);
    }
    // rows [ncols, ncols_pad) of every K / V stage are never written by TMA: keep them zero
    if (ncols_pad > ncols) {
        const int pad_bytes = (ncols_pad - ncols) * G::kRowBytes;
        for (int t = 0; t < nstage * 2 * G::kSlabs; ++t) {
            uint8_t* base = sKV + t * kv_slab_bytes + ncols * G::kRowBytes;
            for (int i = tid * 16; i < pad_bytes; i += kFwdThreads * 16) *reinterpret_cast<uint4*>(base + i) = make_uint4(0, 0, 0, 0);
        }
    }
    build_mask_tiles(sRm, sCm, pl, sh, h0, w0, tid, kFwdThreads);
    fence_proxy_async();                 // generic-proxy writes above -> visible to tcgen05.mma
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns: O (one buffer per head parity when it fits) | S x2 | P x2 (bf16 pairs, ncols_pad/2 columns each)
    const uint32_t tmem_s0 = tmem_base + pl.obufs * D;
    const uint32_t tmem_p0 = tmem_s0 + 2 * ncols_pad;
    const int p_cols = ncols_pad >> 1;

    // a step = (head, plane, h-chunk); steps run head-major, then chunk, then plane
    struct Cursor { int hd, ks, chunk; };
    auto advance = [&](Cursor& c) {
        if (++c.ks > ks_last) {
            c.ks = ks_first;
            if (++c.chunk > chunk_last) { c.chunk = chunk_first; ++c.hd; }
        }
    };

    if (warp == kDriverWarp) {
        // =============================== driver ===============================================
        // The whole warp runs this code so that addresses and descriptors stay warp-uniform (uniform
        // datapath); only the TMA / tcgen05 instructions themselves are predicated on one lane.
        const bool leader = elect_one();      // whole driver warp is converged here
        auto q_buf = [&](int hd) { return sQ + (pl.rowbuf == 2 ? (hd & 1) : 0) * q_tile_bytes; };
        auto issue_q_load = [&](int hd) {
            if (leader) {
                uint64_t* bar = &bar_q[hd & 1];
                mbar_expect_tx(bar, (uint32_t)q_tile_bytes);
#pragma unroll
                for (int sl = 0; sl < G::kSlabs; ++sl)
                    tma_load_5d(q_buf(hd) + sl * q_slab_bytes, &map_q, bar, (head0 + hd) * D + sl * G::kSlabCh, w0, h0, s0, b);
            }
        };
        auto issue_kv_load = [&](int stage, const Cursor& c) {
            if (leader) {
                const int cb = (head0 + c.hd) * D;
                uint8_t* dst = sKV + stage * 2 * kv_tile_bytes;
                mbar_expect_tx(&bar_kv[stage], 2u * G::kSlabs * (uint32_t)ncols * G::kRowBytes);
#pragma unroll
                for (int sl = 0; sl < G::kSlabs; ++sl) {
                    tma_load_5d(dst + sl * kv_slab_bytes, &map_kv_k, &bar_kv[stage], cb + sl * G::kSlabCh, w0 - sh.eW,
                                h0 - sh.eH + c.chunk * pl.ch, s0 - sh.eS + c.ks, b);
                    tma_load_5d(dst + kv_tile_bytes + sl * kv_slab_bytes, &map_kv_v, &bar_kv[stage], cb + sl * G::kSlabCh,
                                w0 - sh.eW, h0 - sh.eH + c.chunk * pl.ch, s0 - sh.eS + c.ks, b);
                }
            }
        };
        const uint32_t idesc_s = make_idesc_bf16(ncols_pad, false, false);
        const uint32_t idesc_o = make_idesc_bf16(D, false, true);
        // descriptor bases; a byte offset is added as (offset >> 4) to the low (start address) field
        const uint64_t dq0 = make_smem_desc(smem_u32(sQ), 16, G::kAtomBytes, G::kSwizzleCode);
        const uint64_t dk0 = make_smem_desc(smem_u32(sKV), 16, G::kAtomBytes, G::kSwizzleCode);
        const uint64_t dv0 = make_smem_desc(smem_u32(sKV + kv_tile_bytes), (uint32_t)kv_slab_bytes, G::kAtomBytes, G::kSwizzleCode);
        const uint64_t drm0 = make_smem_desc(smem_u32(sRm), 2048u, 128u, 0u);                       // mask operands: no swizzle
        const uint64_t dcm0 = make_smem_desc(smem_u32(sCm), (uint32_t)ncols_pad * 16u, 128u, 0u);
        const uint32_t q_buf_step = (pl.rowbuf == 2) ? (uint32_t)(q_tile_bytes >> 4) : 0u;
        const uint32_t stage_step = (uint32_t)((2 * kv_tile_bytes) >> 4);
        const int nk_m = pl.km >> 4;
        auto issue_s_mma = [&](int t, int stage, const Cursor& c) {     // S[t&1] = Q_hd K_t^T + R C_chunk^T
            const uint32_t tmem_s = tmem_s0 + (t & 1) * ncols_pad;
            const uint64_t da0 = dq0 + (c.hd & 1) * q_buf_step, db0 = dk0 + stage * stage_step;
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk) {
                const uint32_t off = (uint32_t)(((kk * 16) / G::kSlabCh) * 1 /*slab*/);
                const uint32_t koff = (uint32_t)((((kk * 16) % G::kSlabCh) * 2) >> 4);
                if (leader)
                    umma_bf16_ss(tmem_s, da0 + off * (uint32_t)(q_slab_bytes >> 4) + koff,
                                 db0 + off * (uint32_t)(kv_slab_bytes >> 4) + koff, idesc_s, kk > 0);
            }
            const uint64_t dcm = dcm0 + (uint32_t)c.chunk * (uint32_t)(cm_tile_bytes >> 4);
            for (int kk = 0; kk < nk_m; ++kk)       // 16 mask channels = two 8-channel core-matrix columns per MMA
                if (leader)
                    umma_bf16_ss(tmem_s, drm0 + (uint32_t)kk * (uint32_t)((2 * 2048) >> 4),
                                 dcm + (uint32_t)kk * (uint32_t)((2 * ncols_pad * 16) >> 4), idesc_s, 1u);
            if (leader) umma_commit(&bar_s[t & 1]);
        };
        const int nk_o = ncols_pad / 16;
        auto issue_o_mma = [&](int t, int stage, int hd, bool accumulate) {   // O[hd&1] += P[t&1] V_t
            const uint32_t tmem_o = tmem_base + (pl.obufs == 2 ? (hd & 1) : 0) * D;
            uint32_t ta = tmem_p0 + (t & 1) * p_cols;                                // A = P from tensor memory
            uint64_t db = dv0 + stage * stage_step;
#pragma unroll 3
            for (int kk = 0; kk < nk_o; ++kk) {
                if (leader) umma_bf16_ts(tmem_o, ta, db, idesc_o, (accumulate || kk > 0) ? 1u : 0u);
                ta += 8;                                                             // next 16 keys: 8 bf16-pair columns of P
                db += (uint32_t)((16 * G::kRowBytes) >> 4);                          // next 16 keys of V
            }
            if (leader) umma_commit(&bar_o[t & 1]);
        };

        Cursor ld = {0, ks_first, chunk_first};      // next block to load, its step index and stage
        int ld_t = 0, ld_stage = 0;
        issue_q_load(0);
        for (; ld_t < nstage && ld_t < nsteps; ++ld_t) {
            issue_kv_load(ld_stage, ld);
            advance(ld);
            if (++ld_stage == nstage) ld_stage = 0;
        }
        Cursor cur = {0, ks_first, chunk_first}, nxt = cur;
        advance(nxt);
        mbar_wait(&bar_q[0], 0);
        mbar_wait(&bar_kv[0], 0);
        tc_fence_after();
        issue_s_mma(0, 0, cur);
        int st_cur = 0;                              // stage of step t
        int st_nxt = (nstage > 1) ? 1 : 0;           // stage of step t+1 ...
        uint32_t kv_par = 1u;                        // bit s = parity of stage s's next completion (stage 0 was consumed once)
        const bool dbg_on = (blockIdx.x == 1 && blockIdx.y == 1 && blockIdx.z == 0) && leader;
        (void)dbg_on;
        // With >= 4 K/V stages S runs TWO steps ahead: S_{t+2} is queued right behind O += P_t V_t the moment the
        // compute warps release step t, so they never wait for a score tile and the tensor pipe never idles on us.
        const bool ahead2 = (nstage >= 4) && (nblocks >= 2) && (pl.rowbuf == 2 || pl.hpc == 1) && (nsteps >= 2);
        if (ahead2) {
            Cursor nn = nxt;                         // step t + 2
            advance(nn);
            int st_nn = 2;                           // its stage (nstage >= 4)
            mbar_wait(&bar_kv[1], 0);
            kv_par ^= 1u << 1;
            tc_fence_after();
            issue_s_mma(1, 1, nxt);
            for (int t = 0; t < nsteps; ++t) {
                DBG(0);
                const bool head_start = (cur.ks == ks_first) && (cur.chunk == chunk_first);
                if (head_start && pl.rowbuf == 2 && cur.hd + 1 < pl.hpc) issue_q_load(cur.hd + 1);
                mbar_wait(&bar_p[t & 1], (t >> 1) & 1);      // step t released: P_t written, S buffer t&1 drained
                tc_fence_after();
                DBG(4);
                issue_o_mma(t, st_cur, cur.hd, !head_start);
                DBG(5);
                if (t + 2 < nsteps) {
                    mbar_wait(&bar_kv[st_nn], (kv_par >> st_nn) & 1u);
                    kv_par ^= 1u << st_nn;
                    if (nn.hd != nxt.hd) mbar_wait(&bar_q[nn.hd & 1], (nn.hd >> 1) & 1);
                    tc_fence_after();
                    DBG(1);
                    issue_s_mma(t + 2, st_nn, nn);
                    DBG(2);
                }
                if (t >= 1 && ld_t < nsteps) {               // refill the stage freed by step t-1
                    mbar_wait(&bar_o[(t - 1) & 1], ((t - 1) >> 1) & 1);
                    issue_kv_load(ld_stage, ld);
                    advance(ld);
                    ++ld_t;
                    if (++ld_stage == nstage) ld_stage = 0;
                }
                DBG(3);
                cur = nxt; nxt = nn;
                advance(nn);
                st_cur = st_nxt; st_nxt = st_nn;
                if (++st_nn == nstage) st_nn = 0;
            }
        } else
        for (int t = 0; t < nsteps; ++t) {
            DBG(0);
            const bool head_start = (cur.ks == ks_first) && (cur.chunk == chunk_first);
            // (a) Q of the next head: its buffer was last read by the S MMAs of head hd-1 (rowbuf 2)
            if (head_start && pl.rowbuf == 2 && cur.hd + 1 < pl.hpc) issue_q_load(cur.hd + 1);
            auto refill = [&]() {        // (b) refill the stage freed by step t-1 once its P V has retired
                if (t >= 1 && ld_t < nsteps) {
                    mbar_wait(&bar_o[(t - 1) & 1], ((t - 1) >> 1) & 1);
                    issue_kv_load(ld_stage, ld);          // == stage of step t-1
                    advance(ld);
                    ++ld_t;
                    if (++ld_stage == nstage) ld_stage = 0;
                }
            };
            if (nstage < 3) refill();    // two stages: the block of step t+1 is the one being refilled
            // (c) S of step t+1 into the other TMEM buffer, as soon as its K block is in.  That buffer was
            //     drained by step t-1, which iteration t-1 already waited for (bar_p) before issuing P V.
            if (t + 1 < nsteps) {
                if (nxt.hd != cur.hd && pl.rowbuf == 1) {
                    // single Q buffer: every S MMA of this head has been issued (S_t was) and retired (bar_s of t)
                    mbar_wait(&bar_s[t & 1], (t >> 1) & 1);
                    issue_q_load(nxt.hd);
                }
                mbar_wait(&bar_kv[st_nxt], (kv_par >> st_nxt) & 1u);
                kv_par ^= 1u << st_nxt;
                if (nxt.hd != cur.hd) mbar_wait(&bar_q[nxt.hd & 1], (nxt.hd >> 1) & 1);
                tc_fence_after();
                DBG(1);
                issue_s_mma(t + 1, st_nxt, nxt);
                DBG(2);
            }
            if (nstage >= 3) refill();
            DBG(3);
            // (d) O += P_t V_t once the compute warps have written P_t
            mbar_wait(&bar_p[t & 1], (t >> 1) & 1);
            tc_fence_after();
            DBG(4);
            issue_o_mma(t, st_cur, cur.hd, !head_start);
            DBG(5);
            cur = nxt;
            advance(nxt);
            st_cur = st_nxt;
            if (++st_nxt == nstage) st_nxt = 0;
        }
    } else {
        // =============================== compute warps ==========================================
        // Two threads per query row: warps w and w+4 share TMEM lane quadrant w&3 and split the
        // quadrant's live key columns (in groups of 8) between them.
        constexpr int CP = D / NPART;                  // O columns per thread in rescale / epilogue
        const int quad = warp & 3, part = warp >> 2;
        const int row = quad * 32 + lane;              // query row of the brick == TMEM lane
        const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
        auto quad_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * NPART) : "memory"); };
        const int plane_mask = (1 << pl.lgPlane) - 1;
        const int qs = row >> pl.lgPlane, qh = (row & plane_mask) >> pl.lgTW, qw = row & (pl.tW - 1);
        const bool q_valid = (s0 + qs < sh.S) && (h0 + qh < sh.H) && (w0 + qw < sh.W);
        // live key range of this row in halo coordinates (window AND grid), per axis
        const int kh_lo = max(qh, sh.eH - h0), kh_hi = min(qh + 2 * sh.eH, sh.H - 1 - h0 + sh.eH);
        const int kw_lo = max(qw, sh.eW - w0), kw_hi = min(qw + 2 * sh.eW, sh.W - 1 - w0 + sh.eW);
        const bool row_can = q_valid && kw_hi >= kw_lo;
        // warp-uniform ranges (identical for all warps of a quadrant)
        const int w_qs = (quad * 32) >> pl.lgPlane;
        const int w_qh_lo = ((quad * 32) & plane_mask) >> pl.lgTW, w_qh_hi = ((quad * 32 + 31) & plane_mask) >> pl.lgTW;
        const long tok = (((long)b * sh.S + (s0 + qs)) * sh.H + (h0 + qh)) * sh.W + (w0 + qw);

        // Reference exponent (log2 domain, scaled) this row's P values are relative to.  It starts at 0 and is
        // only moved when a block's row sum leaves [2^-64, 2^64] relative to it: bf16 P and the fp32 sums
        // keep full precision over that range, so no per-block max pass and (in practice) no O rescale is needed.
        float m_used = 0.f;
        float l_part = 0.f;            // this thread's share of the running sum of P
        bool head_has_blocks = false;  // warp-uniform: this quadrant already accumulated a block of the current head
        bool row_seen = false;         // this row had live columns in an earlier block of the current head
        // exchange slot: [use][parity][part][row]; parity alternates so that a slot is rewritten only
        // after a later quad_sync has proven every reader of its previous contents done
        auto xslot = [&](int use, int parity) { return sX + ((use * 2 + (parity & 1)) * 4) * 128; };
        // O / l -> bf16 and the LSE of head `hd`; called once that head's last P V has retired
        auto finish_head = [&](int hd) {
            const uint32_t tmem_o = tmem_base + (pl.obufs == 2 ? (hd & 1) : 0) * D;
            float* x = xslot(2, hd);
            x[part * 128 + row] = l_part;
            quad_sync();
            float l_run = 0.f;
#pragma unroll
            for (int pp = 0; pp < NPART; ++pp) l_run += x[pp * 128 + row];
            const float inv_l = 1.f / l_run;
            const int cb = (head0 + hd) * D;
            __nv_bfloat16* orow = prm.o + tok * (long)sh.inner() + cb + part * CP;
#pragma unroll
            for (int c = 0; c < CP; c += 16) {
                uint32_t r[16];
                tmem_ld16(tmem_o + lane_sel + part * CP + c, r);      // warp-collective: every lane takes part
                tmem_wait_ld();
                if (q_valid) {
                    uint32_t pk[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        pk[i] = pack_bf16(__uint_as_float(r[2 * i]) * inv_l, __uint_as_float(r[2 * i + 1]) * inv_l);
                    *reinterpret_cast<uint4*>(orow + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4*>(orow + c + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
            }
            if (q_valid && part == 0) prm.lse[tok * sh.heads + head0 + hd] = (m_used + lg2(l_run)) * 0.6931471805599453f;
        };

        Cursor cur{0, ks_first, chunk_first};
        bool p_zero[2] = {false, false};   // this thread's share of P buffer i is known to be all zero
        int mask_chunk = -1;
        int g_lo = 0, g_hi = 0;            // live 8-column groups of this quadrant in the current h-chunk
        bool chunk_live = false, row_has_cols = false;
        const int ngroups = ncols_pad >> 3;

        const bool dbg_on = (blockIdx.x == 1 && blockIdx.y == 1 && blockIdx.z == 0 && tid == 0);
        (void)dbg_on;
        for (int t = 0; t < nsteps; ++t) {
            DBG(8);
            const int buf = t & 1;
            const uint32_t tmem_s = tmem_s0 + buf * ncols_pad + lane_sel;
            const uint32_t tmem_o = tmem_base + (pl.obufs == 2 ? (cur.hd & 1) : 0) * D;
            const uint32_t tmem_p = tmem_p0 + buf * p_cols + lane_sel;
            const bool head_start = (cur.ks == ks_first) && (cur.chunk == chunk_first);
            float prev_m = 0.f, prev_l = 0.f;
            if (head_start && t > 0) {               // the previous head is finished AFTER this step (its O buffer is not reused yet)
                prev_m = m_used; prev_l = l_part;
                if (pl.obufs == 1) {                 // single O accumulator: drain it before this head's first P V can be issued
                    mbar_wait(&bar_o[(t - 1) & 1], ((t - 1) >> 1) & 1);
                    tc_fence_after();
                    finish_head(cur.hd - 1);
                }
                m_used = 0.f; l_part = 0.f;
                head_has_blocks = false;
                row_seen = false;
            }
            const int kh0 = cur.chunk * pl.ch;
            if (cur.chunk != mask_chunk) {               // live column range of this row / quadrant for this h-chunk
                mask_chunk = cur.chunk;
                const int ra = max(kh_lo, kh0), rb = min(kh_hi, kh0 + pl.ch - 1);
                row_has_cols = row_can && (rb >= ra);
                // the quadrant's rows see halo rows [w_qh_lo, w_qh_hi + 2 eH]; rows outside the grid are never live
                const int ua = max(max(w_qh_lo, kh0), khg_lo), ub = min(min(w_qh_hi + 2 * sh.eH, kh0 + pl.ch - 1), khg_hi);
                chunk_live = ub >= ua;
                g_lo = ((ua - kh0) * pl.hW) >> 3;
                g_hi = min(((ub - kh0 + 1) * pl.hW + 7) >> 3, ngroups);
            }
            // warp-uniform: can any of this quadrant's queries see this block?
            const bool live = chunk_live && (cur.ks >= w_qs) && (cur.ks <= w_qs + 2 * sh.eS);
            if (live) {
                // S_t computed.  The driver issued it behind O += P_{t-2} V_{t-2} (tcgen05 operations of one thread retire
                // in order), so this also says that P buffer `buf` is free and that every warp has left step t-2.
                mbar_wait(&bar_s[buf], (t >> 1) & 1);
                tc_fence_after();
                DBG(10);
                const int n8 = g_hi - g_lo;
                const int ga = g_lo + ((n8 * part) >> 1), gb = g_lo + ((n8 * (part + 1)) >> 1);
                // Single pass against the stale reference exponent; P may then exceed 1, which is fine up to 2^64.
                bool two_pass = false;
                {
                    const uint64_t cc = pk2(pl.scale_log2, pl.scale_log2), mm = pk2(-m_used, -m_used);
                    uint64_t acc0 = pk2(0.f, 0.f), acc1 = acc0;
                    // 16 scores of this row -> 16 probabilities (bf16) in the P operand; pairs [0, WM_FWD_POLY) take the
                    // polynomial exp2 on the FMA pipe, the others the MUFU: the two pipes run side by side
                    auto cols16 = [&](const uint32_t (&r)[16], int g) {
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const uint64_t x = ffma2(pk2u(r[2 * i], r[2 * i + 1]), cc, mm);
                            uint64_t e;
                            if (i < WM_FWD_POLY) {
                                e = exp2_poly2(x);
                            } else {
                                float x0, x1;
                                upk2(x, x0, x1);
                                e = pk2(ex2(x0), ex2(x1));
                            }
                            if (i & 1) acc1 = fadd2(acc1, e); else acc0 = fadd2(acc0, e);
                            pk[i] = pack_bf16_2(e);
                        }
                        tmem_st8(tmem_p + g * 4, pk);
                    };
                    auto cols8 = [&](const uint32_t (&r)[8], int g) {
                        uint32_t pk[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint64_t x = ffma2(pk2u(r[2 * i], r[2 * i + 1]), cc, mm);
                            float x0, x1;
                            upk2(x, x0, x1);
                            const uint64_t e = pk2(ex2(x0), ex2(x1));
                            if (i & 1) acc1 = fadd2(acc1, e); else acc0 = fadd2(acc0, e);
                            pk[i] = pack_bf16_2(e);
                        }
                        tmem_st4(tmem_p + g * 4, pk[0], pk[1], pk[2], pk[3]);
                    };
                    // two register sets: the TMEM load of the next 16 columns is in flight while these are processed
                    uint32_t ra[16], rb[16];
                    int g = ga;
                    const int n16 = (gb - ga) >> 1;
                    if (n16 > 0) tmem_ld16(tmem_s + g * 8, ra);
                    for (int i = 0; i < n16; i += 2) {
                        tmem_wait_ld();
                        tmem_regs_ready(ra);
                        if (i + 1 < n16) tmem_ld16(tmem_s + (g + 2) * 8, rb);
                        cols16(ra, g);
                        g += 2;
                        if (i + 1 < n16) {
                            tmem_wait_ld();
                            tmem_regs_ready(rb);
                            if (i + 2 < n16) tmem_ld16(tmem_s + (g + 2) * 8, ra);
                            cols16(rb, g);
                            g += 2;
                        }
                    }
                    if (g < gb) {
                        uint32_t r8[8];
                        tmem_ld8(tmem_s + g * 8, r8);
                        tmem_wait_ld();
                        cols8(r8, g);
                    }
                    float a0, a1, a2, a3;
                    upk2(acc0, a0, a1);
                    upk2(acc1, a2, a3);
                    const float lsum = (a0 + a1) + (a2 + a3);
                    float* x = xslot(0, t);
                    x[part * 128 + row] = lsum;
                    quad_sync();
                    const float ltot = x[row] + x[128 + row];
                    // rows with live columns must land in [2^-64, 2^64]; !(a && b) also catches inf / NaN
                    const bool out_of_range = row_has_cols && !(ltot <= 1.8446744e19f && ltot >= 5.4210109e-20f);
                    two_pass = __any_sync(0xffffffffu, out_of_range);
                    if (!two_pass) l_part += lsum;
                }
                if (two_pass) {
                    // pass 1: row maximum over this thread's columns (masked scores are <= -2^60)
                    float mx = -INFINITY;
                    for (int g = ga; g < gb; ++g) {
                        uint32_t r[8];
                        tmem_ld8(tmem_s + g * 8, r);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 8; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
                    }
                    float* x = xslot(1, t);
                    x[part * 128 + row] = mx;
                    quad_sync();
#pragma unroll
                    for (int pp = 0; pp < NPART; ++pp) mx = fmaxf(mx, x[pp * 128 + row]);
                    const float m_blk = row_has_cols ? mx * pl.scale_log2 : -INFINITY;    // scale > 0
                    float alpha = 1.f;
                    // re-centre the reference on this block's maximum: upwards always (alpha < 2^-32), downwards
                    // only while the row has nothing accumulated yet (alpha would overflow otherwise)
                    const bool move = (m_blk != -INFINITY) && (m_blk > m_used + 32.f || (!row_seen && m_blk < m_used - 32.f));
                    if (move) {
                        alpha = row_seen ? ex2(m_used - m_blk) : 0.f;
                        m_used = m_blk;
                    }
                    l_part *= alpha;
                    if (head_has_blocks && __any_sync(0xffffffffu, move)) {   // rescale this thread's share of the O row
                        mbar_wait(&bar_o[(t - 1) & 1], ((t - 1) >> 1) & 1);     // every earlier P V has retired
                        tc_fence_after();
#pragma unroll
                        for (int c = 0; c < CP; c += 8) {
                            uint32_t r[8];
                            tmem_ld8(tmem_o + lane_sel + part * CP + c, r);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                            tmem_st8(tmem_o + lane_sel + part * CP + c, r);
                        }
                        tmem_wait_st();
                    }
                    // pass 2: P = 2^(s*scale*log2e - m) -> bf16 -> TMEM
                    const float neg_m = -m_used;
                    float lsum = 0.f;
                    for (int g = ga; g < gb; ++g) {
                        uint32_t r[8];
                        tmem_ld8(tmem_s + g * 8, r);
                        tmem_wait_ld();
                        float p[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            p[i] = ex2(fmaf(__uint_as_float(r[i]), pl.scale_log2, neg_m));
                            lsum += p[i];
                        }
                        tmem_st4(tmem_p + g * 4, pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
                    }
                    l_part += lsum;
                }
                // columns outside the quadrant's live range: zero, shared round-robin between the parts
                for (int g = part; g < g_lo; g += NPART) tmem_st4(tmem_p + g * 4, 0u, 0u, 0u, 0u);
                for (int g = g_hi + part; g < ngroups; g += NPART) tmem_st4(tmem_p + g * 4, 0u, 0u, 0u, 0u);
                p_zero[buf] = false;
                head_has_blocks = true;
                row_seen = row_seen || row_has_cols;
            } else {
                // P buffer `buf` is free (and every warp has left step t-2) once the P V of step t-2 has retired
                if (t >= 2) mbar_wait(&bar_o[buf], ((t - 2) >> 1) & 1);
                if (!p_zero[buf]) {
                    for (int g = part; g < ngroups; g += NPART) tmem_st4(tmem_p + g * 4, 0u, 0u, 0u, 0u);
                    p_zero[buf] = true;
                }
            }
            DBG(11);
            tmem_wait_st();               // P is in tensor memory
            tc_fence_before();            // ... and our tcgen05.ld of S_t are complete before the driver reuses the buffers
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_p[buf]);
            DBG(12);
            if (head_start && t > 0 && pl.obufs == 2) {   // epilogue of the previous head, off the critical path
                mbar_wait(&bar_o[(t - 1) & 1], ((t - 1) >> 1) & 1);     // its last P V has retired
                tc_fence_after();
                const float keep_m = m_used, keep_l = l_part;
                m_used = prev_m; l_part = prev_l;
                finish_head(cur.hd - 1);
                m_used = keep_m; l_part = keep_l;
            }
            advance(cur);
        }
        mbar_wait(&bar_o[(nsteps - 1) & 1], ((nsteps - 1) >> 1) & 1);
        tc_fence_after();
        finish_head(pl.hpc - 1);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kDriverWarp) {
        if (pl.tmem_cols <= 128) tmem_dealloc<128>(tmem_base);
        else if (pl.tmem_cols <= 256) tmem_dealloc<256>(tmem_base);
        else tmem_dealloc<512>(tmem_base);
    }
}

template <int D>
static int launch_fwd(const void* q, const void* k, const void* v, void* o, float* lse, const AttnShape& s,
                      const Plan& pl, cudaStream_t st) {
    using G = Geo<D>;
    CUtensorMap mq, mk, mv;
    const int C = s.inner();
    if (int rc = make_tensor_map_5d(&mq, q, s.B, s.S, s.H, s.W, C, G::kSlabCh, pl.tW, pl.tH, pl.tS, G::kSwizzleBytes)) return rc;
    if (int rc = make_tensor_map_5d(&mk, k, s.B, s.S, s.H, s.W, C, G::kSlabCh, pl.hW, pl.ch, 1, G::kSwizzleBytes)) return rc;
    if (int rc = make_tensor_map_5d(&mv, v, s.B, s.S, s.H, s.W, C, G::kSlabCh, pl.hW, pl.ch, 1, G::kSwizzleBytes)) return rc;
    FwdParams prm{s, pl, static_cast<__nv_bfloat16*>(o), lse};
    WM_CUDA_CHECK(cudaFuncSetAttribute(l3d_fwd_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem_bytes));
    const dim3 grid((unsigned)(pl.tilesW * pl.tilesH), (unsigned)pl.tilesS, (unsigned)(s.B * (s.heads / pl.hpc)));
    if (grid.y > 65535u || grid.z > 65535u) return fail(WM_EUNSUPPORTED, "grid too large for the tensor-core kernel");
    l3d_fwd_tc_kernel<D><<<grid, kFwdThreads, pl.smem_bytes, st>>>(mq, mk, mv, prm);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}

}  // namespace tc

#if WM_EXPERIMENT == 7
extern "C" __attribute__((visibility("default"))) int wm_debug_read(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, tc::g_dbg, sizeof(long long) * 64 * 16);
}
#endif

bool attn_tc_supported(const AttnShape& s) {
    tc::Plan pl;
    return tc::make_plan(s, tc::kFwd, pl) && tc::make_plan(s, tc::kBwdDQws, pl) && tc::make_plan(s, tc::kBwdDKVws, pl);
}

int attn_fwd_tc(const void* q, const void* k, const void* v, void* o, float* lse, const AttnShape& s, cudaStream_t st) {
    tc::Plan pl;
    if (!tc::make_plan(s, tc::kFwd, pl)) return fail(WM_EUNSUPPORTED, "no tensor-core tiling for this shape");
    switch (s.d) {
        case 32: return tc::launch_fwd<32>(q, k, v, o, lse, s, pl, st);
        case 64: return tc::launch_fwd<64>(q, k, v, o, lse, s, pl, st);
        default: return tc::launch_fwd<128>(q, k, v, o, lse, s, pl, st);
    }
}

int attn_bwd_tc(const void* q, const void* k, const void* v, const void* o, const float* lse, const void* dout,
                void* dq, void* dk, void* dv, float* delta, const AttnShape& s, cudaStream_t st) {
    return tc::launch_bwd_tc(q, k, v, o, lse, dout, dq, dk, dv, delta, s, st);
}

}  // namespace wm
