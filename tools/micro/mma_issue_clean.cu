// Micro-benchmark 2: cleanest possible issue loops for tcgen05.mma (M=128, bf16, TS and SS), to separate the
// hardware's per-issuing-thread cost from driver-code overhead.  Variants of how the issuing lane is chosen:
//   0: whole warp runs the loop, `if (lane == 0)` around each MMA (what the attention drivers do)
//   1: `if (lane == 0)` around the whole loop
//   2: whole warp runs the loop, elect.sync picks the lane before every chain
//   3: elect.sync once, before the loop
#include <cstdio>
#include <cuda_runtime.h>
#include "../../world_modelz_b200/csrc/tc_common.cuh"
using namespace wm::tc;

struct Res { long long issue, done; };


// BMODE 0: B K-major SW128; 1: B MN-major SW64 (the V operand of P*V at dim_head 32).  CONTEND: warps 4-7 run tcgen05.ld/st loops
template <int N, bool TS, int VAR, int CHAIN, int BMODE = 0, int CONTEND = 0>
__global__ void __launch_bounds__(256) bench(int issuers, int reps, Res* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    __shared__ uint64_t bar[8];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
    if (warp == 7) tmem_alloc<512>(&tmem_slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp < issuers) {
        const uint64_t da = make_smem_desc(smem_u32(smem), 16, 1024, 2);
        const uint64_t db0 = BMODE == 0 ? make_smem_desc(smem_u32(smem + 16384), 16, 1024, 2) : make_smem_desc(smem_u32(smem + 16384), 16384, 512, 4);
        constexpr uint32_t idesc = make_idesc_bf16(N, false, BMODE == 1);
        const uint32_t d = tmem + (((warp * N) % (496 - N + 1)) & ~31);
        const uint32_t ta0 = tmem + 496;
        const bool leader = lane == 0;
        const bool elected_once = elect_one();
        long long t0 = clock64();
        uint32_t phase = 0;
        auto body = [&](bool go) {
            uint64_t db = db0;
            uint32_t ta = ta0;
#pragma unroll
            for (int i = 0; i < CHAIN; ++i) {
                if (go) {
                    if constexpr (TS) umma_bf16_ts(d, ta, db, idesc, 1);
                    else umma_bf16_ss(d, da + (i & 3) * 2, db, idesc, 1);
                }
                if (BMODE == 0) db += 2 * ((i & 3) == 3 ? -3 : 1);
                else db += 64 * ((i & 7) == 7 ? -7 : 1);
                if (TS && BMODE == 1) ta = ta0 - 64 + 8 * ((i + 1) & 7);
            }
            if (go) umma_commit(&bar[warp]);
        };
        for (int r = 0; r < reps; ++r) {
            if constexpr (VAR == 0) body(leader);
            else if constexpr (VAR == 1) { if (leader) body(true); __syncwarp(); }
            else if constexpr (VAR == 2) body(elect_one());
            else body(elected_once);
            if (r == reps - 1) break;
            mbar_wait(&bar[warp], phase);
            phase ^= 1;
        }
        long long t1 = clock64();
        mbar_wait(&bar[warp], phase);
        long long t2 = clock64();
        if (blockIdx.x == 0 && lane == 0) { out[warp].issue = t1 - t0; out[warp].done = t2 - t0; }
    }
    if (CONTEND && warp >= 4) {
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256;
        uint32_t r[16];
        for (int it = 0; it < CONTEND * reps * CHAIN; ++it) {
            tmem_ld16(base + (it & 7) * 16, r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] += 1;
            tmem_st16(base + 128 + (it & 3) * 16, r);
            tmem_wait_st();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 7) tmem_dealloc<512>(tmem);
}


// The MMA mix of one dK/dV step at the config-3 shape: T = 4 SS MMAs (N=144), dV += P^T dO = 9 TS MMAs (N=32, B MN-major),
// dK += dS^T Q = 9 SS MMAs (N=32, B MN-major).  which: 1 = T only, 2 = TS chain only, 4 = SS chain only, 7 = all.
__global__ void __launch_bounds__(256) bench_mix(int which, int reps, Res* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 100000 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); fence_barrier_init(); }
    if (warp == 7) tmem_alloc<512>(&tmem_slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        const uint64_t da_k = make_smem_desc(smem_u32(smem), 16, 512, 4);                  // 128 x 32 bf16 brick, K-major SW64
        const uint64_t db_k = make_smem_desc(smem_u32(smem + 8192), 16, 512, 4);           // 144 x 32 block, K-major SW64
        const uint64_t db_mn = make_smem_desc(smem_u32(smem + 8192), 16384, 512, 4);       // same block read MN-major
        const uint64_t da_ds = make_smem_desc(smem_u32(smem + 32768), 16, 1024, 2);        // dS^T: 128 x 64 slabs, K-major SW128
        constexpr uint32_t id_t = make_idesc_bf16(144, false, false), id_acc = make_idesc_bf16(32, false, true);
        const bool go = elect_one();
        long long t0 = clock64();
        uint32_t phase = 0;
        for (int r = 0; r < reps; ++r) {
            if (go) {
                if (which & 1) {
#pragma unroll
                    for (int op = 0; op < 2; ++op)
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) umma_bf16_ss(tmem + 64 + op * 144, da_k + kk * 2, db_k + kk * 2, id_t, kk > 0);
                }
#pragma unroll
                for (int kk = 0; kk < 9; ++kk) {
                    if (which & 2) umma_bf16_ts(tmem, tmem + 352 + kk * 8, db_mn + kk * 64, id_acc, 1);
                    if (which & 4) umma_bf16_ss(tmem + 32, da_ds + (kk >> 2) * 1024 + (kk & 3) * 2, db_mn + kk * 64, id_acc, 1);
                }
                umma_commit(&bar[0]);
            }
            __syncwarp();
            mbar_wait(&bar[0], phase);
            phase ^= 1;
        }
        long long t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) out[0].done = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 7) tmem_dealloc<512>(tmem);
}

template <int N, bool TS, int VAR, int CHAIN, int BMODE = 0, int CONTEND = 0>
void run(Res* out) {
    cudaFuncSetAttribute(bench<N, TS, VAR, CHAIN, BMODE, CONTEND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
    for (int issuers : {1, 2, 4}) {
        const int reps = 1024 / CHAIN;
        for (int it = 0; it < 2; ++it) {
            bench<N, TS, VAR, CHAIN, BMODE, CONTEND><<<148, 256, 60000>>>(issuers, reps, out);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
        }
        const double n = (double)reps * CHAIN;
        printf("N=%3d %s var=%d chain=%2d bmode=%d contend=%d issuers=%d | issue %6.1f  done %6.1f cycles/MMA | %6.2f MMAs/kcycle/SM\n", N, TS ? "TS" : "SS", VAR,
               CHAIN, BMODE, CONTEND, issuers, out[0].issue / n, out[0].done / n, 1000.0 * n * issuers / out[0].done);
    }
}

int main() {
    Res* out;
    cudaMallocManaged(&out, 8 * sizeof(Res));
    cudaFuncSetAttribute(bench_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, 110000);
    for (int which : {1, 2, 4, 6, 7}) {
        for (int it = 0; it < 2; ++it) {
            bench_mix<<<148, 256, 110000>>>(which, 64, out);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        }
        printf("dK/dV step mix which=%d: %.0f cycles per step (issue + execute + commit round trip)\n", which, out[0].done / 64.0);
    }
    return 0;
}
