// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, bf16, SS mode, no-swizzle K-major) as a function of N and
// of the number of distinct accumulators the instruction stream cycles through (dependency distance).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../delivr_cfos_b200/csrc umma_bench.cu -o umma_bench
#include <cstdio>
#include <cstdlib>
#include "dlv_common.cuh"
using namespace dlv;

struct Args { int n, nacc, iters, astep, rl, bstep, concurrent_sts; long long* out; };

__global__ void __launch_bounds__(256, 1) k(Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 0) {
        long long t0 = 0, t1 = 0;
        if (elect_one_sync()) {
            const uint32_t idesc = umma_idesc_bf16_m128(a.n);
            const uint32_t abase = smem_u32(smem), bbase = smem_u32(smem + 160 * 1024);
            const uint64_t ad = umma_desc_kmajor_noswz(abase, a.rl * 16, 128);
            const uint64_t bd = umma_desc_kmajor_noswz(bbase, a.n * 16, 128);
            // warm
            for (int i = 0; i < a.nacc; ++i) umma_bf16(tm + i * a.n, ad, bd, idesc, 0u);
            umma_commit(&bar);
            while (!mbar_try_wait(&bar, 0)) {}
            t0 = clock64();
            int acc = 0; int ao = 0; int bo = 0;
            for (int i = 0; i < a.iters; ++i) {
                umma_bf16(tm + acc * a.n, ad + ao, bd + bo, idesc, 1u);
                if (++acc == a.nacc) acc = 0;
                ao += a.astep; if (ao > 2048) ao = 0;
                bo += a.bstep; if (bo > 512) bo = 0;
            }
            umma_commit(&bar);
            while (!mbar_try_wait(&bar, 1)) {}
            t1 = clock64();
            if (blockIdx.x == 0) a.out[0] = t1 - t0;
        }
        __syncwarp();
    } else if (a.concurrent_sts && warp >= 4) {
        // background shared-memory store traffic (like the transform warps)
        uint32_t addr = smem_u32(smem + 100 * 1024) + (threadIdx.x - 128) * 16;
        for (int i = 0; i < a.concurrent_sts; ++i)
            asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" :: "r"(addr + (i & 7) * 2048), "r"(i) : "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
    long long* out; cudaMalloc(&out, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2048;
    printf("N nacc astep bstep sts cycles/mma  (tensor floor N/2)  smemB/clk\n");
    for (int n : {32, 64, 96, 128, 192, 256})
        for (int nacc : {1, 2, 4, 8, 16}) {
            if (nacc * n > 512) continue;
            for (int astep : {0, 1})
                for (int sts : {0}) {
                    Args a{n, nacc, iters, astep, 648, 0, sts, out};
                    k<<<148, 256, 200 * 1024>>>(a);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
                    const double per = double(c) / iters;
                    printf("%3d %4d %5d %5d %3d %9.1f  %6.1f  %7.1f\n", n, nacc, astep, 0, sts, per, n / 2.0, (128 + n) * 32.0 / per);
                }
        }
    // concurrent STS traffic, N = 96
    for (int sts : {0, 20000, 200000}) {
        Args a{96, 4, iters, 1, 648, 0, sts, out};
        k<<<148, 256, 200 * 1024>>>(a);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
        printf("N=96 nacc=4 astep=1 sts=%d: %.1f cycles/mma\n", sts, double(c) / iters);
    }
    return 0;
}
