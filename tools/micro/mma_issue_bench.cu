// Micro-benchmark: what does one tcgen05.mma (M=128, bf16) cost at small N, as a function of the number of
// issuing warps, SS vs TS operands and accumulator reuse?  Build: see tools/micro/build.sh.  Not part of the library.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../world_modelz_b200/csrc/tc_common.cuh"

using namespace wm::tc;

struct Res { long long issue, done; };

// mode bit0: TS (A from TMEM); rot: number of accumulators each issuer rotates over; K-steps share operands
__global__ void __launch_bounds__(256) bench(int issuers, int N, int ts, int rot, int reps, int per_commit, Res* out) {
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
        const uint64_t da = make_smem_desc(smem_u32(smem), 16, 1024, 2);            // A: 128 rows x 64 bf16, K-major SW128
        const uint64_t db = make_smem_desc(smem_u32(smem + 16384), 16, 1024, 2);    // B: N rows x 64 bf16
        const uint32_t idesc = make_idesc_bf16(N, false, false);
        const int slots = (ts ? 496 : 512) / N;
        long long t0 = clock64();
        uint32_t phase = 0;
        for (int r = 0; r < reps; ++r) {
            if (lane == 0) {
                for (int i = 0; i < per_commit; ++i) {
                    const int slot = (warp * rot + (i % rot)) % slots;
                    const uint64_t koff = (uint64_t)((i & 3) * 2);                   // 32 B along K inside the atom
                    if (ts) umma_bf16_ts(tmem + slot * N, tmem + 496, db + koff, idesc, 1);
                    else umma_bf16_ss(tmem + slot * N, da + koff, db + koff, idesc, 1);
                }
                umma_commit(&bar[warp]);
            }
            __syncwarp();
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

int main() {
    Res* out;
    cudaMallocManaged(&out, 8 * sizeof(Res));
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
    const int Ns[] = {32, 64, 144, 256};
    printf("N ts issuers rot per_commit | cycles/MMA issue-side, cycles/MMA to completion (issuer 0), MMAs/kcycle per SM\n");
    for (int N : Ns)
        for (int ts = 0; ts < 2; ++ts)
            for (int issuers : {1, 2, 4})
                for (int rot : {1, 4})
                    for (int pc : {9, 64}) {
                        if (rot * issuers * N > 496 && rot > 1) continue;
                        const int reps = 512 / pc + 1;
                        for (int it = 0; it < 2; ++it) {
                            bench<<<148, 256, 60000>>>(issuers, N, ts, rot, reps, pc, out);
                            if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                        }
                        const double n = (double)reps * pc;
                        printf("%3d %d %d %d %2d | %7.1f %7.1f | %6.2f\n", N, ts, issuers, rot, pc, out[0].issue / n, out[0].done / n,
                               1000.0 * n * issuers / out[0].done);
                    }
    return 0;
}
