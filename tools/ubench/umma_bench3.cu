// Microbenchmark 3: depth of the tcgen05.mma issue queue, cost of tcgen05.commit inside a stream, effect of issue gaps.
#include <cstdio>
#include "dlv_common.cuh"
using namespace dlv;
struct Args { int test, n, gap, groups; long long* out; };

__global__ void __launch_bounds__(128, 1) k(Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
    if (warp == 1) tmem_alloc(&slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 1) {
        if (elect_one_sync()) {
            const uint32_t hi = (128u >> 4) | (1u << 14);
            const uint32_t a0 = ((smem_u32(smem) >> 4) & 0x3FFF) | (392u << 16);
            const uint32_t b0 = ((smem_u32(smem + 110 * 1024) >> 4) & 0x3FFF) | (96u << 16);
            const uint32_t id96 = umma_idesc_bf16_m128(96);
            for (int i = 0; i < 4; ++i) umma_bf16_lh(tm + i * 96, a0, b0, hi, id96, 0u);
            umma_commit(&bar[0]);
            while (!mbar_try_wait(&bar[0], 0)) {}
            if (a.test == 0) {                      // queue depth: issue n, time the issue side and the completion
                const long long t0 = clock64();
                for (int i = 0; i < a.n; ++i) umma_bf16_lh(tm + (i & 1) * 256, a0 + (i & 7), b0, hi, id96, 1u);
                const long long t1 = clock64();
                umma_commit(&bar[1]);
                while (!mbar_try_wait(&bar[1], 0)) {}
                const long long t2 = clock64();
                if (blockIdx.x == 0) { a.out[0] = t1 - t0; a.out[1] = t2 - t0; }
            } else {                                // groups of n MMAs, each followed by 2 commits and an issue gap
                uint32_t ph = 0;
                const long long t0 = clock64();
                for (int g = 0; g < a.groups; ++g) {
                    for (int i = 0; i < a.n; ++i) umma_bf16_lh(tm + (i & 1) * 256, a0 + (i & 7), b0, hi, id96, 1u);
                    if (a.test >= 2) { umma_commit(&bar[2]); umma_commit(&bar[3]); }
                    if (a.gap) { const long long s = clock64(); while (clock64() - s < a.gap) {} }
                }
                umma_commit(&bar[1]);
                while (!mbar_try_wait(&bar[1], 0)) {}
                const long long t2 = clock64();
                (void)ph;
                if (blockIdx.x == 0) { a.out[0] = 0; a.out[1] = t2 - t0; }
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tm, 512);
}

int main() {
    long long* out; cudaMalloc(&out, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    auto run = [&](Args a) {
        k<<<148, 128, 200 * 1024>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
        long long c[2]; cudaMemcpy(c, out, 16, cudaMemcpyDeviceToHost);
        return std::pair<long long, long long>(c[0], c[1]);
    };
    for (int n : {1, 2, 4, 8, 16, 32, 64, 128}) {
        auto r = run(Args{0, n, 0, 0, out});
        printf("depth test n=%3d: issue side %6lld cycles, complete %6lld cycles (%.1f/mma)\n", n, r.first, r.second, double(r.second) / n);
    }
    for (int test : {1, 2})
        for (int gap : {0, 100, 200, 400, 800}) {
            auto r = run(Args{test, 45, gap, 40, out});
            printf("groups of 45, commits=%d gap=%3d: %.1f cycles/group (ideal %d)\n", test >= 2, gap, double(r.second) / 40, 45 * 56);
        }
    return 0;
}
