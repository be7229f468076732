// Tiling plan and shared-memory geometry shared by the tensor-core attention kernels.
#pragma once

#include "tc_common.cuh"
#include "wm_common.cuh"

namespace wm {
namespace tc {

enum Mode { kFwd = 0, kBwdDQ = 1, kBwdDKV = 2, kBwdDQws = 3, kBwdDKVws = 4 };   // ws: warp-specialised backward (dim_head <= 64)

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
    int osplit;                // forward: independent accumulation chains of O += P V (summed in the epilogue)
    int smem_bytes;
    int tmem_cols;             // power of two
    float scale_log2;
};

constexpr int kThreads = 256;        // bwd kernels: 2 threads per brick row (column halves)
constexpr int kFwdThreads = 288;     // fwd: 8 compute warps + 1 driver warp (TMA + MMA issue)
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
size_t smem_bytes_for(Mode mode, int d, int ncols_pad, int nstage, int rowbuf);
int tmem_cols_for(Mode mode, int d, int ncols_pad);
bool make_plan(const AttnShape& s, Mode mode, Plan& best);

// swizzled byte offset of 16-byte chunk `chunk16` of row `row` inside a 128B-row tile
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk16) {
    return (uint32_t)row * 128u + (uint32_t)((chunk16 ^ (row & 7)) << 4);
}

int launch_bwd_ws(int mode, const void* a1, const void* a2, const void* b1, const void* b2, const float* lse,
                  const float* delta, void* out1, void* out2, const AttnShape& s, cudaStream_t st);
int launch_bwd_tc(const void* q, const void* k, const void* v, const void* o, const float* lse, const void* dout,
                  void* dq, void* dk, void* dv, float* delta, const AttnShape& s, cudaStream_t st);

}  // namespace tc
}  // namespace wm
