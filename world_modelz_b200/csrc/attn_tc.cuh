// Tiling plan and shared-memory geometry shared by the tensor-core attention kernels.
#pragma once

#include "tc_common.cuh"
#include "wm_common.cuh"

namespace wm {
namespace tc {

enum Mode { kFwd = 0, kBwdDQws = 3, kBwdDKVws = 4 };   // forward | backward dQ | backward dK/dV (all warp-specialised)

struct Plan {
    int tS, tH, tW;            // row brick (queries; keys in the dK/dV kernel), tS*tH*tW == 128
    int hS, hH, hW;            // halo = brick + 2*extent
    int ch;                    // halo h-rows per block
    int nchunk;                // ceil(hH / ch)
    int ncols;                 // ch*hW: halo columns per block that TMA writes
    int ncols_pad;             // rounded up to 16 (MMA K granularity of the second GEMM)
    int tilesS, tilesH, tilesW;
    int lgTW, lgPlane;         // log2(tW), log2(tH*tW): brick dims are powers of two
    int hpc;                   // heads walked by one CTA (divides heads)
    int nstage;                // halo-block stages in shared memory (2 or 3)
    int rowbuf;                // row-brick buffers (2 when the CTA walks several heads)
    int obufs;                 // forward: O accumulator sets in TMEM (2 = one per head parity)
    int km;                    // channels of the mask operand appended to S = Q K^T: round_up(tH + tW, 16)
    int smem_bytes;
    int tmem_cols;             // power of two
    float scale_log2;
};

constexpr int kMaxItemsPerCta = 64;  // rows of the per-CTA work table of the persistent kernels (2 KB of shared memory)
constexpr int kFwdThreads = 352;     // fwd: 8 compute warps + S issuer + P V issuer + TMA loader
constexpr int kSmemLimit = 227 * 1024;

template <int D> struct Geo {
    static constexpr int kRowBytes = (D == 32) ? 64 : 128;        // one smem row of one channel slab
    static constexpr int kSlabs = (D == 128) ? 2 : 1;             // 64-channel slabs
    static constexpr int kSlabCh = (D == 32) ? 32 : 64;
    static constexpr int kSwizzleBytes = kRowBytes;
    static constexpr uint32_t kSwizzleCode = (D == 32) ? 4u : 2u;  // UMMA layout type
    static constexpr int kAtomBytes = 8 * kRowBytes;               // 8-row swizzle atom
};

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
inline int next_pow2(int v) { int p = 32; while (p < v) p <<= 1; return p; }

// shared memory / tensor memory needed by a kernel family for a given block width
size_t smem_bytes_for(Mode mode, int d, int ncols_pad, int nstage, int rowbuf, int km, int nchunk);
int tmem_cols_for(Mode mode, int d, int ncols_pad);
bool make_plan(const AttnShape& s, Mode mode, Plan& best);

// ---------------------------------------------------------------------------------------------------------------
// Window / grid-border masking inside the score MMA.
//
// The reference overwrites the scores of keys outside the grid with -1e9 (local_3d_attention.py:92-94); keys of the
// dense brick x halo tile that lie outside a query's window must be excluded as well.  Inside one halo s-plane that
// condition separates per axis: (row r, column c) is live iff  kh_c in [rh_r, rh_r + 2 eH]  and  kw_c in [rw_r, rw_r +
// 2 eW]  and the column's token is inside the grid.  So the mask is a rank-(tH + tW) product that the tensor core
// adds for free:  S += R C^T  with  R[r] = onehot(rh_r) | onehot(rw_r)  and  C[c][j] = 0 where the condition holds
// for row coordinate j, -2^60 where it does not (pad columns: -2^60 everywhere).  A masked score then is <= -2^60 and
// its exponential exactly 0 (ex2.approx.ftz) -- no per-element mask test, no mask words in shared memory.  The
// s-axis needs no term: a TMEM lane quadrant (32 rows) has a single s coordinate, so it is a warp-uniform skip.
// Both tiles are bf16, K-major, UMMA "interleave" (no swizzle) layout: 8-row x 16-byte core matrices, 128 B apart
// along the rows (SBO), rows/8 * 128 B apart along K (LBO).
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t kMaskNeg = 0xDD80u;      // bf16 -2^60
constexpr uint32_t kMaskOne = 0x3F80u;      // bf16 1.0

__device__ __forceinline__ void build_mask_tiles(uint8_t* sRm, uint8_t* sCm, const Plan& pl, const AttnShape& sh, int h0,
                                                 int w0, int tid, int nthreads) {
    const int kc = pl.km >> 3;                                   // 16-byte chunks per row
    const int plane_mask = (1 << pl.lgPlane) - 1;
    for (int i = tid; i < 128 * kc; i += nthreads) {
        const int r = i % 128, j = i / 128;
        const int rh = (r & plane_mask) >> pl.lgTW, rw = r & (pl.tW - 1);
        uint32_t wds[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k0 = 8 * j + 2 * q, k1 = k0 + 1;
            const uint32_t lo = (k0 == rh || k0 == pl.tH + rw) ? kMaskOne : 0u;
            const uint32_t hi = (k1 == rh || k1 == pl.tH + rw) ? kMaskOne : 0u;
            wds[q] = lo | (hi << 16);
        }
        *reinterpret_cast<uint4*>(sRm + j * 2048 + (r >> 3) * 128 + (r & 7) * 16) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
    }
    const int lbo = pl.ncols_pad * 16;
    const int tile = pl.ncols_pad * pl.km * 2;
    for (int i = tid; i < pl.nchunk * pl.ncols_pad * kc; i += nthreads) {
        const int c = i % pl.ncols_pad, j = (i / pl.ncols_pad) % kc, cidx = i / (pl.ncols_pad * kc);
        const int khl = c / pl.hW, kw = c - khl * pl.hW;
        const int kh = cidx * pl.ch + khl;
        const int gh = h0 - sh.eH + kh, gw = w0 - sh.eW + kw;
        const bool ok_c = (c < pl.ncols) && gh >= 0 && gh < sh.H && gw >= 0 && gw < sh.W;
        uint32_t wds[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t pair = 0u;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = 8 * j + 2 * q + e;
                bool ok;
                if (k < pl.tH) ok = ok_c && kh >= k && kh <= k + 2 * sh.eH;
                else if (k < pl.tH + pl.tW) ok = ok_c && kw >= (k - pl.tH) && kw <= (k - pl.tH) + 2 * sh.eW;
                else ok = true;
                pair |= (ok ? 0u : kMaskNeg) << (16 * e);
            }
            wds[q] = pair;
        }
        *reinterpret_cast<uint4*>(sCm + cidx * tile + j * lbo + (c >> 3) * 128 + (c & 7) * 16) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
    }
}

// swizzled byte offset of 16-byte chunk `chunk16` of row `row` inside a 128B-row tile
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk16) {
    return (uint32_t)row * 128u + (uint32_t)((chunk16 ^ (row & 7)) << 4);
}

int launch_bwd_tc(const void* q, const void* k, const void* v, const void* o, const float* lse, const void* dout,
                  void* dq, void* dk, void* dv, float* delta, const AttnShape& s, cudaStream_t st);

}  // namespace tc
}  // namespace wm
