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
#include <string.h>
#include <mutex>
#include <type_traits>
#include <unordered_map>

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

// Descriptor cache: the library owns nothing but these -- encoded CUtensorMaps keyed on (pointer, shape,
// box, swizzle).  A training step or a sampling loop presents the same few activations again and again (the caching
// allocator hands the same blocks back), so eager launches skip the ~1-2 us driver call per map.  Mutex-guarded,
// bounded (cleared when full); the map itself is passed to the kernel by value, so eviction is harmless.
struct MapKey {
    const void* base;
    int v[11];
    bool operator==(const MapKey& o) const { return base == o.base && memcmp(v, o.v, sizeof(v)) == 0; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = std::hash<const void*>()(k.base);
        for (int i = 0; i < 11; ++i) h = h * 1000003u ^ (size_t)k.v[i];
        return h;
    }
};
static std::mutex g_map_mutex;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;

int make_tensor_map_5d(CUtensorMap* out, const void* base, int B, int S, int H, int W, int C, long ld, int box_c,
                       int box_w, int box_h, int box_s, int swizzle_bytes) {
    const MapKey key{base, {B, S, H, W, C, box_c, box_w, box_h, box_s, swizzle_bytes, (int)ld}};
    {
        std::lock_guard<std::mutex> lock(g_map_mutex);
        auto it = g_map_cache.find(key);
        if (it != g_map_cache.end()) { *out = it->second; return WM_OK; }
    }
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr) return fail(WM_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)S, (cuuint64_t)B};
    const cuuint64_t strides[4] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * W, (cuuint64_t)ld * 2 * W * H,
                                   (cuuint64_t)ld * 2 * W * H * S};      // ld: elements between consecutive tokens (>= C)
    const cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_s, 1u};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(WM_ECUDA, "cuTensorMapEncodeTiled failed (%d) for box (%d,%d,%d,%d) of [%d,%d,%d,%d,%d]", (int)r,
                    box_c, box_w, box_h, box_s, B, S, H, W, C);
    std::lock_guard<std::mutex> lock(g_map_mutex);
    if (g_map_cache.size() >= 1024) g_map_cache.clear();
    g_map_cache.emplace(key, *out);
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
            if (s.heads % h == 0 && (double)s.B * tiles * (s.heads / h) >= 4.0 * sm_count()) { hpc = h; break; }
        for (int nchunk = 1; nchunk <= p.hH; ++nchunk) {
            p.ch = (p.hH + nchunk - 1) / nchunk;
            p.nchunk = (p.hH + p.ch - 1) / p.ch;
            p.ncols = p.ch * p.hW;
            p.ncols_pad = round_up(p.ncols, 16);
            if (p.ncols_pad > max_cols || tmem_cols_for(mode, s.d, p.ncols_pad) > 512) continue;
            // shared-memory variants, best first
            bool fits = false;
            const int opts[6][3] = {{4, 2, hpc}, {3, 2, hpc}, {2, 2, hpc}, {4, 1, 1}, {3, 1, 1}, {2, 1, 1}};    // {nstage, rowbuf, hpc}: a head loop needs two row buffers
            for (const auto& o : opts) {
                if (mode != kFwd && o[0] > 3) continue;                                // 4 stages: forward kernel only
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
__device__ long long g_dbg[64 * 32];
#define DBG(slot) do { if (dbg_on && t < 64) g_dbg[t * 32 + (slot)] = clock64(); } while (0)
#define DBGT(slot, tt) do { if (dbg_on && (tt) < 64) g_dbg[(tt) * 32 + (slot)] = clock64(); } while (0)
#define DBGQ(slot) do { if (dbg_blk && lane == 0 && t < 64) g_dbg[t * 32 + (slot)] = clock64(); } while (0)
#else
#define DBG(slot) do { } while (0)
#define DBGT(slot, tt) do { } while (0)
#define DBGQ(slot) do { } while (0)
#endif

// ------------------------------------------------------------------------------ kernel
struct FwdParams {
    AttnShape sh;
    Plan pl;
    __nv_bfloat16* o;
    float* lse;
};

#ifndef WM_SKEL
#define WM_SKEL 0             // timing experiments only: 1 no element math, 2 no P V MMAs, 4 no S MMAs, 8 no K/V reloads,
                              // 16 no S loads, 32 no exp / sum, 64 no P stores, 128 no agreement barrier
#endif
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
// NK > 0: the kernel is compiled for blocks of exactly 16 * NK key columns (one P V chain instead of a 16-way switch of
// unrolled chains: the issuing warps' code shrinks by ~10x, which matters to the instruction caches they share with the
// compute warps); NK == 0: any block width.
template <int D, int NK>
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
    uint64_t* bar_p = bars + 8;       // [2]  P buffer written             (the 4 warps of the team that owns the buffer)
    uint64_t* bar_o = bars + 10;      // [2]  O += P V of a step retired   (tcgen05.commit)
    uint64_t* bar_sfree = bars + 12;  // [2]  S buffer drained             (the 4 warps of the team; long before P is complete)
    uint64_t* bar_head = bars + 14;   // [2]  last O += P V of a head retired (tcgen05.commit), per head parity
    uint64_t* bar_l = bars + 16;      // [2]  the other team's share of a head's row sums is in shared memory (its 4 warps)
    uint64_t* bar_ofree = bars + 18;  // [2]  a head's O accumulator has been read out (the 4 warps of its epilogue team)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

    // ---- work: persistent CTAs (see attn_tc_bwd_ws.cu) --------------------------------------------------------------
    // A CTA serves ONE (h, w) brick position -- masks, border ranges and live column ranges depend on nothing else -- and
    // walks the work items (batch element, s position, head group) of that position back to back.
    const int nclass = pl.tilesH * pl.tilesW;
    const int cls = (int)blockIdx.x % nclass, rank = (int)blockIdx.x / nclass;
    const int nrank = ((int)gridDim.x - cls + nclass - 1) / nclass;          // CTAs sharing this brick position
    const int tw_i = cls % pl.tilesW, th_i = cls / pl.tilesW;
    const int hgroups = sh.heads / pl.hpc;
    const int n_items = sh.B * pl.tilesS * hgroups;
    const int h0 = th_i * pl.tH, w0 = tw_i * pl.tW;

    // block iteration space: h-chunks that intersect the grid (planes: per item)
    const int khg_lo = max(0, sh.eH - h0), khg_hi = min(pl.hH - 1, sh.H - 1 - h0 + sh.eH);
    int chunk_first = 0, chunk_last = 0;
    for (int c = 0; c < pl.nchunk; ++c) {           // nchunk is tiny; avoids integer divisions
        if (khg_lo >= (c + 1) * pl.ch) chunk_first = c + 1;
        if (khg_hi >= c * pl.ch) chunk_last = c;
    }
    const int nchunks_live = chunk_last - chunk_first + 1;
    struct Item { int b, s0, head0, ks_first, ks_last, pad0, pad1, pad2; };
    Item* sItems = reinterpret_cast<Item*>(sX + 4 * 128);               // [kMaxItemsPerCta] rows of 32 bytes, behind the four row-sum slots
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
    int nsteps = 0;
    for (int i = 0; i < my_items; ++i) nsteps += (sItems[i].ks_last - sItems[i].ks_first + 1) * nchunks_live * pl.hpc;
    const int nheads = my_items * pl.hpc;

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
            mbar_init(&bar_p[i], 4);               // one arrival per compute warp of the owning team
            mbar_init(&bar_sfree[i], 4);
            mbar_init(&bar_o[i], 1);
            mbar_init(&bar_head[i], 1);
            mbar_init(&bar_l[i], 4);
            mbar_init(&bar_ofree[i], 4);
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

    // a step = (item, head, h-chunk, plane); H counts heads across items (buffer / barrier parities run on it)
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
    const Cursor first = {0, 0, my_items > 0 ? sItems[0].ks_first : 0, chunk_first, 0};

    if (warp >= kDriverWarp) {
        // =============================== issuing warps ===========================================
        // warp 8: S = Q K^T (+ mask) two steps ahead; warp 9: O += P V; warp 10: TMA loads.  Each runs warp-uniform
        // code (addresses and descriptors stay in the uniform datapath); only the TMA / tcgen05 instructions
        // themselves are predicated on one elected lane.  One thread pays 20-70 cycles per tcgen05.mma it issues, so
        // the per-step chains (3 score MMAs, ncols/16 P V MMAs, 2 TMA boxes) are spread over three warps.
        const bool leader = elect_one();      // each of these warps is converged here
        auto q_buf = [&](int H) { return sQ + (pl.rowbuf == 2 ? (H & 1) : 0) * q_tile_bytes; };
        auto issue_q_load = [&](const Cursor& c) {        // Q tile of c's head
            if (leader) {
                const Item it = sItems[c.slot];
                uint64_t* bar = &bar_q[c.H & 1];
                mbar_expect_tx(bar, (uint32_t)q_tile_bytes);
#pragma unroll
                for (int sl = 0; sl < G::kSlabs; ++sl)
                    tma_load_5d(q_buf(c.H) + sl * q_slab_bytes, &map_q, bar, (it.head0 + c.hd) * D + sl * G::kSlabCh, w0, h0, it.s0, it.b);
            }
        };
        auto issue_kv_load = [&](int stage, const Cursor& c) {
            if (leader) {
                const Item it = sItems[c.slot];
                const int s0 = it.s0, b = it.b;
                const int cb = (it.head0 + c.hd) * D;
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
        // Issue discipline (tools/micro/mma_issue_clean.cu, sync_latency.cu): ONE branch on the elected lane around a
        // whole chain of tcgen05.mma whose operand offsets are compile-time constants costs ~20-35 cycles per MMA; a
        // predicate per MMA, or a run-time trip count, costs ~100 (VOTEU / R2UR.BROADCAST per instruction).  Hence the
        // chains below are fully unrolled templates selected by a switch on the (per-launch constant) trip count.
        auto issue_s_mma = [&](int t, int stage, const Cursor& c) {     // S[t&1] = Q_hd K_t^T + R C_chunk^T
            if (leader) {
                const uint32_t tmem_s = tmem_s0 + (t & 1) * ncols_pad;
                const uint64_t da0 = dq0 + (c.H & 1) * q_buf_step, db0 = dk0 + stage * stage_step;
#pragma unroll
                for (int kk = 0; kk < ((WM_SKEL & 4) ? 0 : D / 16); ++kk) {
                    const uint32_t off = (uint32_t)(((kk * 16) / G::kSlabCh) * 1 /*slab*/);
                    const uint32_t koff = (uint32_t)((((kk * 16) % G::kSlabCh) * 2) >> 4);
                    umma_bf16_ss(tmem_s, da0 + off * (uint32_t)(q_slab_bytes >> 4) + koff,
                                 db0 + off * (uint32_t)(kv_slab_bytes >> 4) + koff, idesc_s, kk > 0);
                }
                const uint64_t dcm = dcm0 + (uint32_t)c.chunk * (uint32_t)(cm_tile_bytes >> 4);
                // 16 mask channels = two 8-channel core-matrix columns per MMA (km is 16 or 32)
                if (!(WM_SKEL & 4)) umma_bf16_ss(tmem_s, drm0, dcm, idesc_s, 1u);
                if (nk_m > 1 && !(WM_SKEL & 4)) umma_bf16_ss(tmem_s, drm0 + (uint32_t)((2 * 2048) >> 4), dcm + (uint32_t)((2 * ncols_pad * 16) >> 4), idesc_s, 1u);
                umma_commit(&bar_s[t & 1]);
            }
            __syncwarp();
        };
        const int nk_o = ncols_pad / 16;
        auto issue_o_chain = [&](auto nk_tag, uint32_t tmem_o, uint32_t ta, uint64_t db, bool accumulate, uint64_t* bar) {
            constexpr int NK = decltype(nk_tag)::value;
#pragma unroll
            for (int kk = 0; kk < ((WM_SKEL & 2) ? 0 : NK); ++kk)
                umma_bf16_ts(tmem_o, ta + 8 * kk /*8 bf16-pair columns of P per 16 keys*/,
                             db + (uint32_t)kk * (uint32_t)((16 * G::kRowBytes) >> 4), idesc_o, (accumulate || kk > 0) ? 1u : 0u);
            umma_commit(bar);
        };
        auto issue_o_mma = [&](int t, int stage, int hd, bool accumulate, bool head_end) {   // O[H&1] += P[t&1] V_t  (hd: head counter H)
            if (leader) {
                const uint32_t tmem_o = tmem_base + (pl.obufs == 2 ? (hd & 1) : 0) * D;
                const uint32_t ta = tmem_p0 + (t & 1) * p_cols;                          // A = P from tensor memory
                const uint64_t db = dv0 + stage * stage_step;
                uint64_t* bar = &bar_o[t & 1];
                if constexpr (NK > 0) {
                    issue_o_chain(std::integral_constant<int, NK>{}, tmem_o, ta, db, accumulate, bar);
                } else {
#define WM_O_CASE(n) case n: issue_o_chain(std::integral_constant<int, n>{}, tmem_o, ta, db, accumulate, bar); break;
                    switch (nk_o) {
                        WM_O_CASE(1) WM_O_CASE(2) WM_O_CASE(3) WM_O_CASE(4) WM_O_CASE(5) WM_O_CASE(6) WM_O_CASE(7) WM_O_CASE(8)
                        WM_O_CASE(9) WM_O_CASE(10) WM_O_CASE(11) WM_O_CASE(12) WM_O_CASE(13) WM_O_CASE(14) WM_O_CASE(15)
                        default: issue_o_chain(std::integral_constant<int, 16>{}, tmem_o, ta, db, accumulate, bar); break;
                    }
#undef WM_O_CASE
                }
                if (head_end) umma_commit(&bar_head[hd & 1]);
            }
            __syncwarp();
        };
        uint32_t dbg_flag = (blockIdx.x == 1) && leader;
        asm volatile("" : "+r"(dbg_flag));
        const bool dbg_on = dbg_flag != 0;
        (void)dbg_on;
        Cursor cur = first;

        if (warp == kDriverWarp + 2) {
            // ---- TMA loader: K/V blocks nstage steps ahead, Q one head ahead ---------------------------------------
            Cursor ld = cur;
            int ld_t = 0, ld_stage = 0;
            issue_q_load(cur);
            for (; ld_t < nstage && ld_t < nsteps; ++ld_t) {
                issue_kv_load(ld_stage, ld);
                advance(ld);
                if (++ld_stage == nstage) ld_stage = 0;
            }
            for (int t = 0; t < nsteps; ++t) {
                const bool head_start = is_head_start(cur);
                // O += P V of step t-1 has retired: its K/V stage is free, and (the compute warps had to see S_{t-1} before
                // they released that step) every S MMA of the head that ended with step t-1 has retired as well
                if (t >= 1) mbar_wait(&bar_o[(t - 1) & 1], ((t - 1) >> 1) & 1);
                if (head_start && pl.rowbuf == 2 && cur.H + 1 < nheads) {            // two Q buffers: the next head's tile
                    Cursor nh = cur;
                    next_head(nh);
                    issue_q_load(nh);
                } else if (head_start && pl.rowbuf == 1 && t >= 1) {
                    issue_q_load(cur);                                              // one Q buffer: this head's tile, now that it is free
                }
                if (t >= 1 && ld_t < nsteps) {
                    if (WM_SKEL & 8) { if (leader) mbar_arrive(&bar_kv[ld_stage]); } else
                    issue_kv_load(ld_stage, ld);          // == stage of step t-1
                    advance(ld);
                    ++ld_t;
                    if (++ld_stage == nstage) ld_stage = 0;
                }
                advance(cur);
            }
        } else if (warp == kDriverWarp) {
            // ---- S of step t+2 as soon as the compute warps have drained S buffer t&1 and the K block is in ----------
            uint32_t kv_par = 0u;                        // bit s = parity of stage s's next completion
            int stage = 0;
            auto wait_inputs = [&](const Cursor& c, bool new_head) {
                mbar_wait(&bar_kv[stage], (kv_par >> stage) & 1u);
                kv_par ^= 1u << stage;
                if (new_head) mbar_wait(&bar_q[c.H & 1], (c.H >> 1) & 1);
                tc_fence_after();
            };
            Cursor c2 = cur;
            int prev_hd = -1;
            for (int u = 0; u < 2 && u < nsteps; ++u) {                  // S_0, S_1
                wait_inputs(c2, c2.H != prev_hd);
                issue_s_mma(u, stage, c2);
                prev_hd = c2.H;
                advance(c2);
                if (++stage == nstage) stage = 0;
            }
            for (int t = 0; t + 2 < nsteps; ++t) {
                DBG(2);
                mbar_wait(&bar_sfree[t & 1], (t >> 1) & 1);              // S buffer t&1 drained
                DBG(3);
                wait_inputs(c2, c2.H != prev_hd);
                DBG(6);
                issue_s_mma(t + 2, stage, c2);
                DBG(7);
                prev_hd = c2.H;
                advance(c2);
                if (++stage == nstage) stage = 0;
            }
        } else {
            // ---- O += P_t V_t once the compute warps have written P_t -------------------------------------------------
            uint32_t kv_par = 0u;
            int stage = 0;
            for (int t = 0; t < nsteps; ++t) {
                DBG(0);
                const bool head_start = is_head_start(cur);
                mbar_wait(&bar_kv[stage], (kv_par >> stage) & 1u);       // V block (long since there: S_t came from its K)
                kv_par ^= 1u << stage;
                DBG(1);
                mbar_wait(&bar_p[t & 1], (t >> 1) & 1);
                if (head_start && cur.H >= pl.obufs) {       // the accumulator's previous head (H - obufs) has been read out
                    const int h = cur.H - pl.obufs;
                    mbar_wait(&bar_ofree[h & 1], (h >> 1) & 1);
                }
                tc_fence_after();
                DBG(4);
                const int hd_now = cur.H;
                advance(cur);
                issue_o_mma(t, stage, hd_now, !head_start, cur.H != hd_now);
                DBG(5);
                if (++stage == nstage) stage = 0;
            }
        }
    } else {
        // =============================== compute warps ==========================================
        // Two TEAMS of four warps (one per TMEM lane quadrant).  Team b owns S / P buffer b, i.e. the steps t with
        // t & 1 == b, and one of its threads handles a whole query row of such a step.  The two warps of an SM
        // sub-partition (w, w + 4: same quadrant, different teams) are therefore always in different steps: while
        // one sits in the fixed latencies around a step (TMEM store drain, hand-off to the issuing warps, S MMA of its
        // next step) the other one is in its element math.  The S buffer is handed back the moment its last
        // tcgen05.ld has returned (bar_sfree), long before the step's P is complete (bar_p).
        const int quad = warp & 3, team = warp >> 2;
        const int row = quad * 32 + lane;              // query row of the brick == TMEM lane
        const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
        const int plane_mask = (1 << pl.lgPlane) - 1;
        const int qs = row >> pl.lgPlane, qh = (row & plane_mask) >> pl.lgTW, qw = row & (pl.tW - 1);
        const bool hw_valid = (h0 + qh < sh.H) && (w0 + qw < sh.W);
        // warp-uniform ranges (identical for all warps of a quadrant)
        const int w_qs = (quad * 32) >> pl.lgPlane;
        const int w_qh_lo = ((quad * 32) & plane_mask) >> pl.lgTW, w_qh_hi = ((quad * 32 + 31) & plane_mask) >> pl.lgTW;

        // Softmax without a running maximum.  P = 2^(s*scale*log2e) is taken relative to the FIXED exponent 0: bf16 P, the
        // fp32 row sums and the fp32 accumulation of P V keep full relative precision as long as the row sum stays
        // inside [2^-100, 2^100], i.e. for logits within about +-69 nats -- no max pass, no O rescale, and the two
        // threads of a row never have to agree on anything.  A row whose sum leaves that range (or is inf / NaN) gets
        // LSE = NaN here and is recomputed exactly by l3d_fwd_fixup_kernel, launched right behind this kernel.
        float l_part = 0.f;            // this thread's share (its team's steps) of the running sum of P
        // Head epilogue without a rendezvous between the teams (they must stay out of step).  The team that owns the
        // LAST step of a head only deposits its share of the row sums (shared memory + bar_l) and moves on; the other
        // team -- whose own work on that head ended a step earlier -- writes the whole row: it picks the deposit up,
        // reads all D columns of O once the head's last P V has retired (bar_head), releases the accumulator to the
        // P V issuer (bar_ofree) and stores O / l and the LSE.  It does so after its first step of the NEXT head, so
        // that the wait for the retiring MMAs overlaps its own math.
        auto xslot = [&](int hd) { return sX + (hd & 3) * 128; };      // four slots: see the reuse argument at bar_ofree
        auto deposit_l = [&](int hd, float l_mine) {
            xslot(hd)[row] = l_mine;
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_l[hd & 1]);
        };
        auto finish_head = [&](int hd, float l_mine) {          // hd: head counter H (across items)
            const Item it = sItems[hd / pl.hpc];
            const int head = it.head0 + hd % pl.hpc;
            const bool q_valid = hw_valid && (it.s0 + qs < sh.S);
            const long tok = (((long)it.b * sh.S + (it.s0 + qs)) * sh.H + (h0 + qh)) * sh.W + (w0 + qw);
            mbar_wait(&bar_l[hd & 1], (hd >> 1) & 1);
            const float l_run = l_mine + xslot(hd)[row];
            const bool ok = l_run >= 7.8886091e-31f && l_run <= 1.2676506e30f;       // [2^-100, 2^100]; false for NaN
            const float inv_l = 1.f / l_run;
            mbar_wait(&bar_head[hd & 1], (hd >> 1) & 1);
            tc_fence_after();
            const uint32_t tmem_o = tmem_base + (pl.obufs == 2 ? (hd & 1) : 0) * D + lane_sel;
            __nv_bfloat16* orow = prm.o + tok * (long)sh.inner() + head * D;
            const float lse_out = ok ? lg2(l_run) * 0.6931471805599453f : __int_as_float(0x7fc00000);
#pragma unroll
            for (int c0 = 0; c0 < D; c0 += 32) {           // 32 columns at a time: bounded register footprint at D = 128
                uint32_t r[32];
                tmem_ld32(tmem_o + c0, r);
                tmem_wait_ld();
                if (c0 + 32 == D) {                        // the accumulator has been read out
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_ofree[hd & 1]);
                }
                if (q_valid) {
#pragma unroll
                    for (int c = 0; c < 32; c += 8) {
                        uint32_t pk[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            pk[i] = pack_bf16(__uint_as_float(r[c + 2 * i]) * inv_l, __uint_as_float(r[c + 2 * i + 1]) * inv_l);
                        *reinterpret_cast<uint4*>(orow + c0 + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
            }
            if (q_valid) prm.lse[tok * sh.heads + head] = lse_out;
        };

        const int ngroups = ncols_pad >> 3;
        int dlo = 0, dhi = ngroups;            // 8-column groups of this team's P buffer that may be non-zero
        uint32_t dbg_flag = (blockIdx.x == 1 && (tid & 127) == 0);     // one thread per team
        asm volatile("" : "+r"(dbg_flag));
        const bool dbg_on = dbg_flag != 0;
        (void)dbg_on;
        uint32_t dbg_blk_flag = (blockIdx.x == 1);
        asm volatile("" : "+r"(dbg_blk_flag));
        const bool dbg_blk = dbg_blk_flag != 0;
        (void)dbg_blk;
        const uint64_t cc = pk2(pl.scale_log2, pl.scale_log2), zz = pk2(0.f, 0.f);
        const uint32_t tmem_s = tmem_s0 + team * ncols_pad + lane_sel;
        const uint32_t tmem_p = tmem_p0 + team * p_cols + lane_sel;
        auto s_drained = [&]() {               // every tcgen05.ld of this step's S has returned: the issuer may refill the buffer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_sfree[team]);
        };

        int t = 0;
        bool owned_last_prev = false;            // did this team own the last step of the previous head?
        for (int hd = 0; hd < nheads; ++hd) {    // hd: head counter across this CTA's items
            DBGT(14, t);
            const int ks_first = sItems[hd / pl.hpc].ks_first, ks_last = sItems[hd / pl.hpc].ks_last;
            const bool owns_last = ((t + (ks_last - ks_first + 1) * nchunks_live - 1) & 1) == team;     // of THIS head
            const float prev_l = l_part;
            bool drain_prev = hd > 0 && !owned_last_prev;        // the previous head's epilogue is this team's, still to be done
            l_part = 0.f;
            for (int chunk = chunk_first; chunk <= chunk_last; ++chunk) {
                // ---- live column range of this quadrant in this h-chunk (warp-uniform) ----
                // the quadrant's rows see halo rows [w_qh_lo, w_qh_hi + 2 eH]; rows outside the grid are never live
                const int kh0 = chunk * pl.ch;
                const int ua = max(max(w_qh_lo, kh0), khg_lo), ub = min(min(w_qh_hi + 2 * sh.eH, kh0 + pl.ch - 1), khg_hi);
                const bool chunk_live = ub >= ua;
                const int g_lo = chunk_live ? ((ua - kh0) * pl.hW) >> 3 : 0;
                const int g_hi = chunk_live ? min(((ub - kh0 + 1) * pl.hW + 7) >> 3, ngroups) : 0;
                const int n16 = (g_hi - g_lo) >> 1;
                const bool rem8 = ((g_hi - g_lo) & 1) != 0;
                DBGT(15, t);
                for (int ks = ks_first; ks <= ks_last; ++ks, ++t) {
                    if ((t & 1) != team) continue;
                    DBG(8);
                    // single O accumulator: this head's first P V waits for the read-out (bar_ofree) -- do it right away
                    if (drain_prev && pl.obufs == 1) {
                        finish_head(hd - 1, prev_l);
                        drain_prev = false;
                    }
                    // warp-uniform: can any of this quadrant's queries see this block?
                    const bool live = chunk_live && (ks >= w_qs) && (ks <= w_qs + 2 * sh.eS);
                    DBGQ(20 + quad);
                    if (t >= 2) mbar_wait(&bar_o[team], ((t - 2) >> 1) & 1);     // P buffer free: P V of step t-2 retired
                    DBG(9);
                    mbar_wait(&bar_s[team], (t >> 1) & 1);                       // S_t computed
                    int nl = 0, nh = 0;
                    if (live) {
                        tc_fence_after();
                        DBG(10);
#if !(WM_SKEL & 1)     // WM_SKEL bit 0: pipeline skeleton only (no element math) -- timing floor of TMA + MMA + barriers
                        uint64_t acc0 = zz, acc1 = zz;
                        // 16 scores of this row -> 16 probabilities (bf16) in the P operand; pairs [0, WM_FWD_POLY) take
                        // the polynomial exp2 on the FMA pipe, the others the MUFU: the two pipes run side by side.
                        // Masked scores are <= -2^60 (mask operand of the S MMA): their exponential is 0.
                        auto cols16 = [&](const uint32_t (&r)[16], int g) {
                            uint32_t pk[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const uint64_t x = fmul2(pk2u(r[2 * i], r[2 * i + 1]), cc);
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
                        // two register sets: the TMEM load of the next 16 columns is in flight while these are processed
                        uint32_t r0[16], r1[16];
                        int g = g_lo;
                        if (n16 > 0) tmem_ld16(tmem_s + g * 8, r0);
                        else if (!rem8) s_drained();
                        for (int i = 0; i < n16; i += 2) {
                            tmem_wait_ld();
                            tmem_regs_ready(r0);
                            if (i + 1 < n16) tmem_ld16(tmem_s + (g + 2) * 8, r1);
                            else if (!rem8) s_drained();
                            cols16(r0, g);
                            g += 2;
                            if (i + 1 < n16) {
                                tmem_wait_ld();
                                tmem_regs_ready(r1);
                                if (i + 2 < n16) tmem_ld16(tmem_s + (g + 2) * 8, r0);
                                else if (!rem8) s_drained();
                                cols16(r1, g);
                                g += 2;
                            }
                        }
                        if (rem8) {                   // odd group count: 8 columns left
                            uint32_t r8[8];
                            tmem_ld8(tmem_s + g * 8, r8);
                            tmem_wait_ld();
                            s_drained();
                            uint32_t pk[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint64_t x = fmul2(pk2u(r8[2 * i], r8[2 * i + 1]), cc);
                                float x0, x1;
                                upk2(x, x0, x1);
                                const uint64_t e = pk2(ex2(x0), ex2(x1));
                                if (i & 1) acc1 = fadd2(acc1, e); else acc0 = fadd2(acc0, e);
                                pk[i] = pack_bf16_2(e);
                            }
                            tmem_st4(tmem_p + g * 4, pk[0], pk[1], pk[2], pk[3]);
                        }
                        DBG(13);
                        float a0, a1, a2, a3;
                        upk2(acc0, a0, a1);
                        upk2(acc1, a2, a3);
                        l_part += (a0 + a1) + (a2 + a3);
#else
                        s_drained();
#endif
                        // Columns outside the quadrant's live range must be zero in the P operand.  They stay zero from step
                        // to step: only what an earlier step left non-zero outside today's range is cleared (normally nothing).
                        for (int gz = dlo; gz < min(dhi, g_lo); ++gz) tmem_st4(tmem_p + gz * 4, 0u, 0u, 0u, 0u);
                        for (int gz = max(dlo, g_hi); gz < dhi; ++gz) tmem_st4(tmem_p + gz * 4, 0u, 0u, 0u, 0u);
                        nl = g_lo; nh = g_hi;
                    } else {
                        s_drained();
                        for (int gz = dlo; gz < dhi; ++gz) tmem_st4(tmem_p + gz * 4, 0u, 0u, 0u, 0u);
                    }
                    dlo = nl; dhi = nh;
                    DBG(11);
                    tmem_wait_st();               // P is in tensor memory
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_p[team]);
                    DBG(12);
                    DBGQ(16 + quad);
                    if (drain_prev) {             // epilogue of the previous head, off the critical path (two O accumulators)
                        finish_head(hd - 1, prev_l);
                        drain_prev = false;
                    }
                }
            }
            if (drain_prev) finish_head(hd - 1, prev_l);      // this team had no step in this head
            if (owns_last) deposit_l(hd, l_part);
            owned_last_prev = owns_last;
        }
        if (nheads > 0 && !owned_last_prev) finish_head(nheads - 1, l_part);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kDriverWarp) {
        if (pl.tmem_cols <= 128) tmem_dealloc<128>(tmem_base);
        else if (pl.tmem_cols <= 256) tmem_dealloc<256>(tmem_base);
        else tmem_dealloc<512>(tmem_base);
    }
}

template <int D, int NK>
static int launch_fwd(const void* q, const void* k, const void* v, void* o, float* lse, const AttnShape& s,
                      const Plan& pl, cudaStream_t st) {
    using G = Geo<D>;
    CUtensorMap mq, mk, mv;
    const int C = s.inner();
    if (int rc = make_tensor_map_5d(&mq, q, s.B, s.S, s.H, s.W, C, s.q_ld(), G::kSlabCh, pl.tW, pl.tH, pl.tS, G::kSwizzleBytes)) return rc;
    if (int rc = make_tensor_map_5d(&mk, k, s.B, s.S, s.H, s.W, C, s.kv_ld(), G::kSlabCh, pl.hW, pl.ch, 1, G::kSwizzleBytes)) return rc;
    if (int rc = make_tensor_map_5d(&mv, v, s.B, s.S, s.H, s.W, C, s.kv_ld(), G::kSlabCh, pl.hW, pl.ch, 1, G::kSwizzleBytes)) return rc;
    FwdParams prm{s, pl, static_cast<__nv_bfloat16*>(o), lse};
    WM_CUDA_CHECK(cudaFuncSetAttribute(l3d_fwd_tc_kernel<D, NK>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem_bytes));
    // persistent grid when there is more than one work item per SM (see the kernel's work decomposition), else one CTA per item
    const long items = (long)pl.tilesW * pl.tilesH * pl.tilesS * s.B * (s.heads / pl.hpc);
    if (items > 0x7fffffffL) return fail(WM_EUNSUPPORTED, "grid too large for the tensor-core kernel");
    const bool fits = (items + sm_count() - 1) / sm_count() + pl.tilesH * pl.tilesW <= kMaxItemsPerCta;      // per-CTA item table
    const bool every_position_served = pl.tilesH * pl.tilesW <= sm_count();                                  // a CTA serves ONE brick position
    const unsigned grid = (unsigned)(items > (long)sm_count() && fits && every_position_served ? (long)sm_count() : items);
    l3d_fwd_tc_kernel<D, NK><<<grid, kFwdThreads, pl.smem_bytes, st>>>(mq, mk, mv, prm);
    WM_CUDA_CHECK(cudaGetLastError());
    return WM_OK;
}

}  // namespace tc

#if WM_EXPERIMENT == 7
extern "C" __attribute__((visibility("default"))) int wm_debug_read(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, tc::g_dbg, sizeof(long long) * 64 * 32);
}
#endif

bool attn_tc_supported(const AttnShape& s) {
    tc::Plan pl;
    return tc::make_plan(s, tc::kFwd, pl) && tc::make_plan(s, tc::kBwdDQws, pl) && tc::make_plan(s, tc::kBwdDKVws, pl);
}

int attn_fwd_tc(const void* q, const void* k, const void* v, void* o, float* lse, const AttnShape& s, cudaStream_t st) {
    tc::Plan pl;
    if (!tc::make_plan(s, tc::kFwd, pl)) return fail(WM_EUNSUPPORTED, "no tensor-core tiling for this shape");
    int rc;
    const int nk = pl.ncols_pad / 16;
    // block widths of the named configurations get their own instantiation (3x5x5 window, 2x8x8 brick: 144 columns;
    // 5x7x7: 112), everything else the generic kernel
    if (s.d == 32 && nk == 9) rc = tc::launch_fwd<32, 9>(q, k, v, o, lse, s, pl, st);
    else if (s.d == 128 && nk == 7) rc = tc::launch_fwd<128, 7>(q, k, v, o, lse, s, pl, st);
    else if (s.d == 32) rc = tc::launch_fwd<32, 0>(q, k, v, o, lse, s, pl, st);
    else if (s.d == 64) rc = tc::launch_fwd<64, 0>(q, k, v, o, lse, s, pl, st);
    else rc = tc::launch_fwd<128, 0>(q, k, v, o, lse, s, pl, st);
    if (rc) return rc;
    // rows whose softmax left the range of the max-free formulation (LSE = NaN) are recomputed exactly
    return attn_fwd_fixup(q, k, v, o, lse, s, st);
}

int attn_bwd_tc(const void* q, const void* k, const void* v, const void* o, const float* lse, const void* dout,
                void* dq, void* dk, void* dv, float* delta, const AttnShape& s, cudaStream_t st) {
    return tc::launch_bwd_tc(q, k, v, o, lse, dout, dq, dk, dv, delta, s, st);
}

}  // namespace wm
