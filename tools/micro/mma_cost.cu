// Micro-benchmark 3: tensor-pipe time of small tcgen05.mma (M=128, bf16) in the shapes of the attention kernels, issued
// in the cheapest form (one elected lane, one branch, fully unrolled chain, compile-time offsets), end to end
// (issue + execute + commit -> mbarrier) per MMA.  ACCS = accumulators the chain alternates between.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../world_modelz_b200/csrc/tc_common.cuh"
using namespace wm::tc;

template <int N, bool TS, int CHAIN, int ACCS, int AMODE>   // AMODE (SS only): 0 = A SW128 K-major, 1 = A SW64 K-major (64-byte rows, dim_head 32)
__global__ void __launch_bounds__(128) bench(int reps, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc<512>(&tmem_slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        const uint64_t da = AMODE == 0 ? make_smem_desc(smem_u32(smem), 16, 1024, 2) : make_smem_desc(smem_u32(smem), 16, 512, 4);
        const uint64_t db = AMODE == 0 ? make_smem_desc(smem_u32(smem + 32768), 16, 1024, 2) : make_smem_desc(smem_u32(smem + 32768), 16, 512, 4);
        constexpr uint32_t idesc = make_idesc_bf16(N, false, false);
        const bool go = elect_one();
        long long t0 = clock64();
        uint32_t phase = 0;
        for (int r = 0; r < reps; ++r) {
            if (go) {
#pragma unroll
                for (int i = 0; i < CHAIN; ++i) {
                    const uint32_t d = tmem + (i % ACCS) * 256;
                    if constexpr (TS) umma_bf16_ts(d, tmem + 496 - 8 * (i & 7) - 8, db + (i & 1) * 2, idesc, 1);
                    else umma_bf16_ss(d, da + (i & 1) * 2, db + (i & 1) * 2, idesc, 1);
                }
                umma_commit(&bar);
            }
            __syncwarp();
            mbar_wait(&bar, phase);
            phase ^= 1;
        }
        long long t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem);
}

template <int N, bool TS, int CHAIN, int ACCS, int AMODE = 0>
void run(long long* out) {
    cudaFuncSetAttribute(bench<N, TS, CHAIN, ACCS, AMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int reps = 256;
    for (int it = 0; it < 2; ++it) {
        bench<N, TS, CHAIN, ACCS, AMODE><<<148, 128, 100 * 1024>>>(reps, out);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
    }
    printf("%s N=%3d chain=%2d accumulators=%d amode=%d : %7.1f cycles per chain round trip, %6.1f per MMA\n", TS ? "TS" : "SS", N, CHAIN, ACCS, AMODE,
           out[0] / (double)reps, out[0] / (double)reps / CHAIN);
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    long long* out;
    cudaMallocManaged(&out, 64);
    run<32, true, 9, 1>(out);  run<32, true, 9, 2>(out);  run<32, true, 36, 1>(out); run<32, true, 36, 2>(out);
    run<64, true, 9, 1>(out);  run<128, true, 9, 1>(out); run<256, true, 9, 1>(out); run<128, true, 36, 1>(out);
    run<32, false, 9, 1>(out); run<32, false, 36, 1>(out); run<32, false, 36, 1, 1>(out);
    run<144, false, 3, 1>(out); run<144, false, 3, 1, 1>(out); run<144, false, 36, 1>(out); run<144, false, 36, 1, 1>(out);
    run<256, false, 36, 1>(out); run<112, false, 36, 1>(out);
    return 0;
}
