// Micro-benchmark: what bounds the element loop of the attention kernels on one SM sub-partition?
// A CTA of W warps (W = 4, 8, 16: one, two, four warps per sub-partition) runs, per warp and per iteration, the
// forward kernel's unit of work on 16 score columns of its TMEM lane quadrant:
//     tcgen05.ld x16 (software-pipelined one unit ahead) -> scale -> 2^x -> row sum -> bf16 pack -> tcgen05.st x8
// Variants isolate the pieces:  0 = load + store only,  1 = + scale / sum / pack (no exponential),
//     2 = every exponential on the MUFU,  3 = 3 of 8 pairs on the FMA pipe (the shipped mix),  4 = all on the FMA pipe,
//     5 = MUFU only (no load / store: registers),  6 = backward-style unit (two loads, P and dS, one store).
// Prints cycles per unit per warp and per sub-partition.  No MMAs run: with them, TMEM ports are shared with the tensor
// pipe, so these are lower bounds.
// Build: tools/micro/build.sh (needs -I../../world_modelz_b200/csrc)
#include "tc_common.cuh"
#include <cstdio>
using namespace wm::tc;

template <int V>
__global__ void __launch_bounds__(512, 1) unit_kernel(long long* out, int iters, float scale) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc<512>(&slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot + ((uint32_t)((warp & 3) * 32) << 16);
    const int team = warp >> 2;                       // warps of one sub-partition use different column ranges
    const uint32_t s_base = tm + team * 96, p_base = tm + 384 + team * 32;
    // fill the score columns with something finite
    {
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = __float_as_uint(0.01f * (float)(lane + i));
        for (int c = 0; c < 96; c += 16) tmem_st16(s_base + c, z);
        tmem_wait_st();
    }
    __syncthreads();
    const uint64_t cc = pk2(scale, scale);
    uint64_t acc0 = pk2(0.f, 0.f), acc1 = acc0;
    uint32_t r0[16], r1[16], d0[16];
    auto unit = [&](const uint32_t (&r)[16], int g) {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint64_t e;
            if (V == 0) {
                pk[i] = r[2 * i] ^ r[2 * i + 1];
                continue;
            }
            const uint64_t x = fmul2(pk2u(r[2 * i], r[2 * i + 1]), cc);
            const bool poly = (V == 4) || (V == 3 && i < 3);
            if (V == 1) {
                e = x;
            } else if (poly) {
                e = exp2_poly2(x);
            } else {
                float x0, x1;
                upk2(x, x0, x1);
                e = pk2(ex2(x0), ex2(x1));
            }
            if (i & 1) acc1 = fadd2(acc1, e); else acc0 = fadd2(acc0, e);
            pk[i] = pack_bf16_2(e);
        }
        tmem_st8(p_base + (g & 3) * 8, pk);
    };
    const long long t0 = clock64();
    if (V == 5) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.001f * (lane + i);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = ex2(v[i]) * 0.5f;
        }
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) sum += v[i];
        acc0 = pk2(sum, sum);
    } else if (V == 6) {
        const uint64_t sc2 = pk2(0.17f, 0.17f), nd = pk2(-0.3f, -0.3f), nl = pk2(-1.f, -1.f);
        for (int it = 0; it < iters; ++it) {
            const int g = it % 5;
            tmem_ld16(s_base + g * 16, r0);
            tmem_ld16(s_base + ((g + 1) % 6) * 16, d0);
            tmem_wait_ld();
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint64_t x = ffma2(pk2u(r0[2 * i], r0[2 * i + 1]), cc, nl);
                uint64_t pr;
                if (i < 2) pr = exp2_poly2(x);
                else { float x0, x1; upk2(x, x0, x1); pr = pk2(ex2(x0), ex2(x1)); }
                const uint64_t y = ffma2(pk2u(d0[2 * i], d0[2 * i + 1]), sc2, nd);
                pk[i] = pack_bf16_2(fmul2(pr, y));
            }
            tmem_st8(p_base + (g & 3) * 8, pk);
        }
    } else {
        tmem_ld16(s_base, r0);
        for (int it = 0; it < iters; it += 2) {
            const int g = (it >> 1) % 3;
            tmem_wait_ld();
            tmem_regs_ready(r0);
            tmem_ld16(s_base + (2 * g + 1) * 16, r1);
            unit(r0, 2 * g);
            tmem_wait_ld();
            tmem_regs_ready(r1);
            tmem_ld16(s_base + ((2 * g + 2) % 6) * 16, r0);
            unit(r1, 2 * g + 1);
        }
        tmem_wait_ld();
    }
    tmem_wait_st();
    const long long t1 = clock64();
    float a, b;
    upk2(fadd2(acc0, acc1), a, b);
    if (lane == 0) out[warp] = (t1 - t0) + (a + b == 1234.5f ? 1 : 0);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(slot);
}

template <int V>
static void run(const char* name, long long* d_out) {
    for (int warps : {4, 8, 16}) {
        const int iters = 4096;
        unit_kernel<V><<<1, warps * 32>>>(d_out, iters, 0.25f);
        cudaDeviceSynchronize();
        unit_kernel<V><<<1, warps * 32>>>(d_out, iters, 0.25f);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s: launch failed: %s\n", name, cudaGetErrorString(cudaGetLastError())); return; }
        long long h[16];
        cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
        const double per_warp = (double)mx / iters;
        printf("%-44s warps/SMSP %d: %7.1f cycles per unit per warp, %7.1f per unit per sub-partition   per warp:", name, warps / 4,
               per_warp, per_warp / (warps / 4));
        for (int w = 0; w < warps; ++w) printf(" %.0f", (double)h[w] / iters);
        printf("\n");
    }
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 16 * sizeof(long long));
    run<0>("0 ld16 + st8 only", d_out);
    run<1>("1 + scale, sum, pack (no exp)", d_out);
    run<2>("2 all MUFU", d_out);
    run<3>("3 mix: 3 of 8 pairs polynomial", d_out);
    run<4>("4 all polynomial", d_out);
    run<5>("5 16 independent ex2 + fmul (registers)", d_out);
    run<6>("6 backward unit (2 ld16, P*dS, st8)", d_out);
    return 0;
}
