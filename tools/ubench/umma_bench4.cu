// Microbenchmark 4: SS-mode tcgen05.mma rate with cta_group::2 (M = 256 over a CTA pair, each CTA supplies its own
// 128 A rows and HALF of the B rows) against cta_group::1 (M = 128), K = 16, bf16, for the N values of the fused conv.
// Optional background shared-memory traffic from the other warps (ld.shared + st.shared of 16 B per lane) to see how
// much head-room the operand fetch leaves.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I delivr_cfos_b200/csrc tools/ubench/umma_bench4.cu -o /tmp/umma_bench4
#include <cstdio>
#include <utility>
#include "dlv_common.cuh"
using namespace dlv;
struct Args { int pair, n, iters, bg; long long* out; };

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma2_bf16_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                 "setp.ne.b32 p, %5, 0;\n\t"
                 "mov.b64 da, {%1, %3};\n\t"
                 "mov.b64 db, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) k(Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); stop = 0; }
    __syncthreads();
    if (warp == 1) {
        if (a.pair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            tmem_alloc(&slot, 512);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 1) {
        if ((rank == 0 || !a.pair) && elect_one_sync()) {
            const uint32_t hi = (128u >> 4) | (1u << 14);
            const uint32_t a0 = ((smem_u32(smem) >> 4) & 0x3FFF) | (392u << 16);
            const uint32_t b0 = ((smem_u32(smem + 110 * 1024) >> 4) & 0x3FFF) | (96u << 16);
            const uint32_t id = idesc_bf16(a.pair ? 256 : 128, a.n);
            // warm-up
            for (int i = 0; i < 4; ++i) { if (a.pair) umma2_bf16_lh(tm, a0, b0, hi, id, 0u); else umma_bf16_lh(tm, a0, b0, hi, id, 0u); }
            if (a.pair) umma2_commit_mc(&bar[0], 3); else umma_commit(&bar[0]);
            while (!mbar_try_wait(&bar[0], 0)) {}
            const long long t0 = clock64();
            for (int i = 0; i < a.iters; ++i) {
                const uint32_t d = tm + (i & 1) * 256, aa = a0 + (i & 7) + ((i >> 3) & 1) * 128;
                if (a.pair) umma2_bf16_lh(d, aa, b0, hi, id, 1u); else umma_bf16_lh(d, aa, b0, hi, id, 1u);
            }
            if (a.pair) umma2_commit_mc(&bar[1], 3); else umma_commit(&bar[1]);
            while (!mbar_try_wait(&bar[1], 0)) {}
            const long long t1 = clock64();
            if (blockIdx.x == 0) a.out[0] = t1 - t0;
            stop = 1;
        }
        __syncwarp();
    } else if (warp >= 2 && warp < 2 + a.bg && (rank == 0 || !a.pair)) {
        // background shared-memory traffic: in-place 16 B read-modify-write per lane over a 32 KB region
        uint32_t base = smem_u32(smem + 150 * 1024) + lane * 16 + (warp - 2) * 4096;
        long long cnt = 0;
        while (!stop) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + j * 512) : "memory");
                v.x += 1;
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(base + j * 512), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            }
            cnt += 8;
        }
        if (blockIdx.x == 0 && warp == 2 && lane == 0) a.out[1] = cnt;
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        if (a.pair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512u) : "memory");
        else tmem_dealloc(tm, 512);
    }
}

int main() {
    long long* out; cudaMalloc(&out, 64);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    auto run = [&](Args a) {
        cudaMemset(out, 0, 64);
        k<<<148, 256, 200 * 1024>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
        long long c[2]; cudaMemcpy(c, out, 16, cudaMemcpyDeviceToHost);
        return std::pair<long long, long long>(c[0], c[1]);
    };
    const int iters = 2048;
    for (int bg : {0, 2, 6})
        for (int pair : {0, 1})
            for (int n : {32, 64, 96, 128, 192, 256}) {
                auto r = run(Args{pair, n, iters, bg, out});
                printf("cta_group::%d M=%3d N=%3d bg_warps=%d: %.1f cycles/mma  (bg 16B ld+st pairs per warp: %lld, %.1f B/clk)\n", pair + 1, pair ? 256 : 128, n, bg,
                       double(r.first) / iters, r.second, r.first ? double(r.second) * 32 * 32 * bg / double(r.first) : 0.0);
            }
    return 0;
}
