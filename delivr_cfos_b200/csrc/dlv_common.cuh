// dlv_common.cuh - shared device helpers for the DELiVR blob_detection kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "delivr_cfos_b200 kernels are written for sm_100a (B200) only"
#endif

namespace dlv {

// ---------------------------------------------------------------- generic device helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
// One lane of a fully converged warp.  Unlike `lane == 0`, ptxas knows the elected branch is single-threaded,
// so tcgen05 operands stay in uniform registers (no per-instruction uniformisation loop).
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ uint4 ld_nc_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_na_u4(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

// mish(x) = x * tanh(softplus(x)) = x * t / (t + 2),  t = e^x (e^x + 2)
__device__ __forceinline__ float mish_f(float x) {
    if (x > 20.f) return x;
    float w = __expf(x);
    float t = w * (w + 2.f);
    return x * __fdividef(t, t + 2.f);
}

// Same function in 8 issue slots and without a branch:  mish(x) = x - 2x / (w(w+2) + 2),  w = e^x.
// w = inf (x > 88) gives 1/inf = 0 -> x; w -> 0 gives x - 2x/2 = 0.
__device__ __forceinline__ float mish_fast(float x) {
    float w, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(x * 1.4426950408889634f));
    const float d = fmaf(w, w + 2.f, 2.f);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return fmaf(-2.f * x, r, x);
}

// ---------------------------------------------------------------- packed fp32 pairs (FFMA2 / FADD2 / FMUL2, sm_100)
// Two fp32 lanes in one 64-bit register: one issue slot does the work of two.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// InstanceNorm affine + Mish on 8 bf16 channels (one 16 B position-chunk), packed arithmetic:
//   y = x*a + b;  w = e^y;  mish(y) = y - y / (0.5 w^2 + w + 1) = y + y * rneg,  rneg = -1 / (0.5 w^2 + w + 1).
// Channel pairs i >= NRP take rneg from MUFU.RCP of the NEGATED denominator (w = +inf -> rcp(-inf) = -0 -> y;
// w = 0 -> rcp(-1) = -1 -> 0): 5 packed FMA-pipe ops + 4 MUFU per pair.
// Channel pairs i < NRP keep the MUFU pipe for ex2 only and take the reciprocal from two Newton steps on the FMA pipe
// (seed = magic - bits(d), 5 % error -> 0.26 % -> 7e-6; the signs are arranged so that the second step lands on
// -1/d): 9 packed ops + 4 ALU ops + 2 MUFU per pair.  The exponent is clamped to 2^63 so that d stays finite.
// NRP trades MUFU cycles (16 lanes / clk / SM) against issue slots.
template <int NRP>
__device__ __forceinline__ void norm_mish8_f32(const uint4 u, const f32x2 (&a)[4], const f32x2 (&b)[4], f32x2 (&out)[4]) {
    const f32x2 kLog2e = pk2(1.4426950408889634f, 1.4426950408889634f);
    const f32x2 kNegHalf = pk2(-0.5f, -0.5f), kNegOne = pk2(-1.f, -1.f);
    const f32x2 kHalf = pk2(0.5f, 0.5f), kOne = pk2(1.f, 1.f), kTwo = pk2(2.f, 2.f), kNegTwo = pk2(-2.f, -2.f);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    f32x2 y[4], t[4], d[4];
    float e0[4], e1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = fma2(pk2(bf16_lo(w[i]), bf16_hi(w[i])), a[i], b[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = mul2(y[i], kLog2e);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        upk2(t[i], e0[i], e1[i]);
        if (i < NRP) { e0[i] = fminf(e0[i], 63.f); e1[i] = fminf(e1[i], 63.f); }
        e0[i] = ex2_approx(e0[i]); e1[i] = ex2_approx(e1[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const f32x2 e = pk2(e0[i], e1[i]);
        d[i] = (i < NRP) ? fma2(e, fma2(e, kHalf, kOne), kOne)                // +(0.5 w^2 + w + 1)
                         : fma2(e, fma2(e, kNegHalf, kNegOne), kNegOne);      // -(0.5 w^2 + w + 1)
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        upk2(d[i], e0[i], e1[i]);
        if (i < NRP) {
            const f32x2 r0 = pk2(__uint_as_float(0x7EF311C7u - __float_as_uint(e0[i])), __uint_as_float(0x7EF311C7u - __float_as_uint(e1[i])));
            const f32x2 r1n = mul2(r0, fma2(d[i], r0, kNegTwo));              // -r1
            const f32x2 r2n = mul2(r1n, fma2(d[i], r1n, kTwo));               // -r2
            out[i] = fma2(y[i], r2n, y[i]);
        } else {
            out[i] = fma2(y[i], pk2(rcp_approx(e0[i]), rcp_approx(e1[i])), y[i]);
        }
    }
}
template <int NRP>
__device__ __forceinline__ uint4 norm_mish8(const uint4 u, const f32x2 (&a)[4], const f32x2 (&b)[4]) {
    f32x2 f[4];
    norm_mish8_f32<NRP>(u, a, b, f);
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float lo, hi;
        upk2(f[i], lo, hi);
        o[i] = pack_bf16x2(lo, hi);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t x, uint32_t y) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&x), *reinterpret_cast<__nv_bfloat162*>(&y));
    return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t bf16x2_min(uint32_t x, uint32_t y) {
    __nv_bfloat162 r = __hmin2(*reinterpret_cast<__nv_bfloat162*>(&x), *reinterpret_cast<__nv_bfloat162*>(&y));
    return *reinterpret_cast<uint32_t*>(&r);
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Non-blocking probe (no hardware suspend): used to look at the NEXT step's barriers from inside an MMA burst.
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a CUDA error, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) {
            printf("dlv: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}

// ---------------------------------------------------------------- TMA bulk copy (UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Same with the descriptors passed as 32-bit halves: the upper half of a K-major no-swizzle descriptor (SBO, version)
// is a constant, so all per-instruction descriptor arithmetic is 32-bit (no carry chains on the uniform datapath).
__device__ __forceinline__ void umma_bf16_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                 "setp.ne.b32 p, %5, 0;\n\t"
                 "mov.b64 da, {%1, %3};\n\t"
                 "mov.b64 db, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate) : "memory");
}
// Same, with a collector hint for the A operand: COLL = 1 "fill" (keep A in the collector buffer after this MMA),
// COLL = 2 "lastuse" (take A from the collector instead of shared memory).  Used where two consecutive MMAs multiply
// the SAME A tile against different B tiles: the second one then costs no shared-memory A fetch.
template <int COLL>
__device__ __forceinline__ void umma_bf16_lh_coll(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    if (COLL == 1) {
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                     "setp.ne.b32 p, %5, 0;\n\t"
                     "mov.b64 da, {%1, %3};\n\t"
                     "mov.b64 db, {%2, %3};\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], da, db, %4, p;\n\t}"
                     :: "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate) : "memory");
    } else if (COLL == 2) {
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                     "setp.ne.b32 p, %5, 0;\n\t"
                     "mov.b64 da, {%1, %3};\n\t"
                     "mov.b64 db, {%2, %3};\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], da, db, %4, p;\n\t}"
                     :: "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        umma_bf16_lh(d_tmem, a_lo, b_lo, hi, idesc, accumulate);
    }
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(smem_u32(bar)) : "memory");
}

// Shared-memory matrix descriptor, K-major, no swizzle ("interleave"):
// core matrix = 8 rows x 16 B stored as 128 contiguous bytes (row r at +16*r);
// next 8-row group at +SBO, next 8-element K chunk at +LBO.
__device__ __forceinline__ uint64_t umma_desc_kmajor_noswz(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
    uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14);   // descriptor version 1 (Blackwell)
    return (static_cast<uint64_t>(hi) << 32) | lo;
}
// Instruction descriptor: bf16 A/B (K-major), fp32 D, M=128, N=n.
__host__ __device__ constexpr uint32_t umma_idesc_bf16_m128(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}

// 32 lanes x 32 consecutive fp32 columns: thread i receives lane (base_lane + i), columns [col, col+32).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Zero 32 lanes x 32 consecutive fp32 columns of TMEM (the warp's own lane quadrant).
__device__ __forceinline__ void tmem_zero32(uint32_t taddr) {
    const uint32_t z = 0u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
                 :: "r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Sum v[c] over the 32 lanes of a warp; on return lane c holds the total of column c.
// 31 shuffles instead of 160 (butterfly that halves the live columns each round).  Split in two so that partial
// results of several tiles can be accumulated after the first round (16 live values instead of 32).
__device__ __forceinline__ void warp_transpose_round1(const float (&v)[32], float (&a16)[16]) {
    const bool up = lane_id() & 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float send = up ? v[i] : v[i + 16];
        float keep = up ? v[i + 16] : v[i];
        a16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
}
__device__ __forceinline__ float warp_transpose_finish(const float (&a16)[16]) {
    const int lane = lane_id();
    float a8[8], a4[4], a2[2];
    {
        const bool up = lane & 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float send = up ? a16[i] : a16[i + 8];
            float keep = up ? a16[i + 8] : a16[i];
            a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    {
        const bool up = lane & 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float send = up ? a8[i] : a8[i + 4];
            float keep = up ? a8[i + 4] : a8[i];
            a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    {
        const bool up = lane & 2;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float send = up ? a4[i] : a4[i + 2];
            float keep = up ? a4[i + 2] : a4[i];
            a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
    }
    const bool up = lane & 1;
    float send = up ? a2[0] : a2[1];
    float keep = up ? a2[1] : a2[0];
    return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}
__device__ __forceinline__ float warp_transpose_sum32(const float (&v)[32]) {
    float a16[16];
    warp_transpose_round1(v, a16);
    return warp_transpose_finish(a16);
}

}  // namespace dlv
