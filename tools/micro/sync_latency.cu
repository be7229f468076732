// Micro-benchmark: hand-off latencies that bound a warp-specialised tcgen05 pipeline on sm_100a.
//   (1) mbarrier ping-pong between two warps (try_wait loop / test_wait spin), same and different SM sub-partitions
//   (2) tcgen05.mma chain + tcgen05.commit -> mbarrier wait by the issuing warp and by another warp
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../world_modelz_b200/csrc sync_latency.cu -o _bin/sync_latency
#include "tc_common.cuh"
#include <cstdio>
using namespace wm::tc;

__device__ __forceinline__ void spin_test(uint64_t* bar, uint32_t parity) { for (int i = 0; i < (1 << 14) && !mbar_test(bar, parity); ++i) {} }
__device__ __forceinline__ void wait_b(uint64_t* bar, uint32_t parity) {   // bounded try_wait loop
    uint32_t done = 0;
    for (int i = 0; i < (1 << 14) && !done; ++i)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

template <int MODE>   // 0: try_wait loop, 1: test_wait spin
__global__ void pingpong(long long* out, int partner_warp, int iters) {
    __shared__ uint64_t bars[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
    __syncthreads();
    long long t0 = clock64();
    if (warp == 0) {
        for (int i = 0; i < iters; ++i) {
            if (lane == 0) mbar_arrive(&bars[0]);
            if (MODE == 0) wait_b(&bars[1], i & 1); else spin_test(&bars[1], i & 1);
            __syncwarp();
        }
    } else if (warp == partner_warp) {
        for (int i = 0; i < iters; ++i) {
            if (MODE == 0) wait_b(&bars[0], i & 1); else spin_test(&bars[0], i & 1);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[1]);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / iters;
}

// warp 0 issues `nmma` MMAs + commit, then `waiter` warp waits; measures issue time and issue->wake latency
template <bool TS>
__global__ void mma_commit(long long* out, int nmma, int n, int waiter, int iters, int alt = 0) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, bar_back;
    __shared__ uint32_t slot;
    __shared__ long long t_issue_done, t_start;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar_back, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(&slot);
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    const uint64_t da = make_smem_desc(smem_u32(smem), 16, 1024, 2u), db = make_smem_desc(smem_u32(smem + 16384), 16, 1024, 2u);
    const uint32_t idesc = make_idesc_bf16(n, false, false);
    long long issue_sum = 0, wake_sum = 0;
    for (int it = 0; it < iters; ++it) {
        if (warp == 0) {
            const bool leader = elect_one();
            long long a = clock64();
#pragma unroll 1
            for (int k = 0; k < nmma; ++k) {
                if (leader) {
                    const uint32_t dd = tm + (alt ? (k & 1) * 128 : 0);
                    if (TS) umma_bf16_ts(dd, tm + 256 + 8 * (k & 7), db + 2 * (k & 3), idesc, 1);
                    else umma_bf16_ss(dd, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
                }
            }
            if (leader) umma_commit(&bar);
            long long b = clock64();
            issue_sum += b - a;
            if (lane == 0) { t_start = a; t_issue_done = b; }
        }
        if (warp == waiter) {
            wait_b(&bar, it & 1);
            long long c = clock64();
            tc_fence_after();
            if (lane == 0) { wake_sum += c - *(volatile long long*)&t_issue_done; mbar_arrive(&bar_back); }
            __syncwarp();
        }
        if (warp == 0) wait_b(&bar_back, it & 1);
    }
    if (warp == 0 && lane == 0) out[0] = issue_sum / iters;
    if (warp == waiter && lane == 0) out[1] = wake_sum / iters;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    long long* d; cudaMalloc(&d, 64); long long h[2];
    for (int partner : {1, 4}) {
        pingpong<0><<<1, 256>>>(d, partner, 2000); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("mbarrier ping-pong, try_wait loop, warps 0<->%d: %lld cycles per round trip\n", partner, h[0]);
        pingpong<1><<<1, 256>>>(d, partner, 2000); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("mbarrier ping-pong, test_wait spin, warps 0<->%d: %lld cycles per round trip\n", partner, h[0]);
    }
    cudaFuncSetAttribute(mma_commit<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(mma_commit<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int alt : {0, 1})
        for (int n : {32, 64, 128}) {
            mma_commit<true><<<1, 64, 64 * 1024>>>(d, 16, n, 0, 300, alt); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("TS N=%d x16 (accumulators: %d) + commit: issue %lld, issue-done -> done %lld  => %lld cycles per MMA\n", n, alt + 1, h[0], h[1], (h[0] + h[1]) / 16);
            mma_commit<false><<<1, 64, 64 * 1024>>>(d, 16, n, 0, 300, alt); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("SS N=%d x16 (accumulators: %d) + commit: issue %lld, issue-done -> done %lld  => %lld cycles per MMA\n", n, alt + 1, h[0], h[1], (h[0] + h[1]) / 16);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
