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

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

template <int N, bool TS, int VAR, int CHAIN>
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
        const uint64_t db0 = make_smem_desc(smem_u32(smem + 16384), 16, 1024, 2);
        constexpr uint32_t idesc = make_idesc_bf16(N, false, false);
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
                db += 2 * ((i & 3) == 3 ? -3 : 1);
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
    tc_fence_before();
    __syncthreads();
    if (warp == 7) tmem_dealloc<512>(tmem);
}

template <int N, bool TS, int VAR, int CHAIN>
void run(Res* out) {
    cudaFuncSetAttribute(bench<N, TS, VAR, CHAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
    for (int issuers : {1, 2, 4}) {
        const int reps = 1024 / CHAIN;
        for (int it = 0; it < 2; ++it) {
            bench<N, TS, VAR, CHAIN><<<148, 256, 60000>>>(issuers, reps, out);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
        }
        const double n = (double)reps * CHAIN;
        printf("N=%3d %s var=%d chain=%2d issuers=%d | issue %6.1f  done %6.1f cycles/MMA | %6.2f MMAs/kcycle/SM\n", N, TS ? "TS" : "SS", VAR,
               CHAIN, issuers, out[0].issue / n, out[0].done / n, 1000.0 * n * issuers / out[0].done);
    }
}

int main() {
    Res* out;
    cudaMallocManaged(&out, 8 * sizeof(Res));
    run<32, true, 0, 9>(out);
    run<32, true, 2, 9>(out);
    run<32, true, 3, 9>(out);
    run<32, true, 2, 64>(out);
    run<32, false, 2, 9>(out);
    run<144, false, 0, 2>(out);
    run<144, false, 2, 2>(out);
    run<144, false, 3, 2>(out);
    run<144, false, 2, 64>(out);
    run<64, true, 2, 9>(out);
    run<128, true, 2, 9>(out);
    run<256, false, 2, 64>(out);
    return 0;
}
