// Throughput of the special-function ops the fused InstanceNorm+Mish staging needs (per SM, many warps).
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float tanhf_(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t tanhh2(uint32_t x) { uint32_t y; asm("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ float mish_fast(float x) {
    float w = ex2f(x * 1.4426950408889634f);
    const float d = fmaf(w, w + 2.f, 2.f);
    return fmaf(-2.f * x, rcpf(d), x);
}
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.001f * (threadIdx.x + i);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) v[i] = ex2f(v[i]);
            if (MODE == 1) v[i] = rcpf(v[i]);
            if (MODE == 2) v[i] = mish_fast(v[i]);
            if (MODE == 3) v[i] = tanhf_(v[i]);
            if (MODE == 4) v[i] = __uint_as_float(ex2h2(__float_as_uint(v[i])));
            if (MODE == 5) v[i] = __uint_as_float(tanhh2(__float_as_uint(v[i])));
            if (MODE == 6) v[i] = fmaf(v[i], 1.0001f, 0.5f);
            if (MODE == 7) { float w = ex2f(v[i]); v[i] = fmaf(w, w + 2.f, 2.f); }        // ex2 + 2 fma-pipe
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const char* names[] = {"ex2.f32", "rcp.f32", "mish_fast", "tanh.f32", "ex2.f16x2", "tanh.f16x2", "ffma", "ex2+2fma"};
    for (int threads : {256, 512, 1024}) {
        for (int mode = 0; mode < 8; ++mode) {
            const int iters = 2000;
            switch (mode) {
                case 0: k<0><<<148, threads>>>(out, iters, cyc); break; case 1: k<1><<<148, threads>>>(out, iters, cyc); break;
                case 2: k<2><<<148, threads>>>(out, iters, cyc); break; case 3: k<3><<<148, threads>>>(out, iters, cyc); break;
                case 4: k<4><<<148, threads>>>(out, iters, cyc); break; case 5: k<5><<<148, threads>>>(out, iters, cyc); break;
                case 6: k<6><<<148, threads>>>(out, iters, cyc); break; case 7: k<7><<<148, threads>>>(out, iters, cyc); break;
            }
            cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            const double ops = double(iters) * 8 * threads;      // per SM
            printf("threads %4d %-10s: %.2f lane-ops/clk/SM  (%.1f cycles per warp-instruction-equivalent per SMSP)\n", threads, names[mode],
                   ops / c, c / (ops / 32 / 4));
        }
    }
    return 0;
}
