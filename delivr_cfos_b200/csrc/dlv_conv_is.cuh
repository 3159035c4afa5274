// dlv_conv_is.cuh - input-stationary fused 3x3x3 convolution for the Cout = 32 layers of the U-Net
// (91 % of the network's MACs).
//
// Replaces, for those layers, the cuDNN conv3d + ATen instance_norm + mish calls issued from
// predictor(window_data) (inference/sliding_window_inferer.py:222; BasicUNet of inference/inference.py:190-197).
//
// Why a second kernel: with N = 32 an SS-mode tcgen05.mma reads (128 + 32) x 32 B of shared memory for 16
// cycles of tensor work - the shared-memory port (128 B/clk), not the tensor pipe, is the limit (measured:
// l1tex 78 %, tensor pipe 31 %).  Here the roles of the z taps are turned around: ONE staged input plane is
// multiplied against the three dz weight blocks concatenated along N (N = 96), and the three products are
// accumulated into the accumulators of the three output planes (z-1, z, z+1) that this input plane feeds.
// The accumulators live in a ring of S plane slots per 128-row tile in TMEM, so 9 MMAs of N = 96 replace
// 27 MMAs of N = 32 and every input plane is fetched (and normalised) exactly once per column.
//
// Work item = (window, column of R = 128 T in-plane positions, z segment).  Per input plane ("step"):
//   transform warps  raw bf16 activations -> (x*a + b) -> mish -> bf16, halo forced to zero, written straight
//                    into the canonical K-major UMMA layout in shared memory (InstanceNorm3d + Mish of the
//                    PRODUCING layer fused into this layer's operand staging; no separate elementwise pass);
//   producer warp    TMA bulk copies for input chunks that need no transform (pooled / upsampled tensors);
//   MMA warp         T x KB x 9 tcgen05.mma (N = 96, split where the slot ring wraps or a slot starts);
//   epilogue warps   completed output plane: TMEM -> registers -> bf16 raw store + InstanceNorm partial sums.
#pragma once
#include "dlv_common.cuh"

namespace dlv {

constexpr int kIsThreads = 320;          // warp 0 producer, 1 MMA, 2-5 epilogue, 6-9 transform
constexpr int kIsMaxStages = 4;

struct IsArgs {
    const __nv_bfloat16* in0;   // leading input chunks
    const __nv_bfloat16* in1;   // remaining chunks (second tensor of a concatenation) or nullptr
    int nch0;                   // chunks held by in0
    int nchunks;                // 2 * KB
    int xform_chunks;           // leading chunks that hold RAW conv output and get norm+mish applied (0 or 4)
    const double* in_stats;     // [nwin][32][2] sum / sum of squares of the producing layer (when xform_chunks)
    const float* in_gamma;      // [32]
    const float* in_beta;       // [32]
    int64_t inS;                // positions per chunk of the inputs
    int in_guard;
    const __nv_bfloat16* w;     // [KB][9][2][96][8]
    __nv_bfloat16* out;         // raw conv output, 4 chunks
    int64_t outS;
    int out_guard;
    double* part;               // [nwin][nparts][32][2] partial sums of this layer's output
    int nparts;
    int Z, Y, X, Xp, PL, Vp;
    int KB, NC, NZS, Zs, nitems, RL, H, nstages;
    uint32_t stage_bytes, w_bytes;
    double inv_count;           // 1 / (Z*Y*X)
};

// deterministic reduction of the per-item partial sums: stats[win][c][0..1] = sum over parts (fixed order)
__global__ void is_reduce_stats_kernel(const double* __restrict__ part, int nparts, double* __restrict__ stats) {
    const int win = blockIdx.x, t = threadIdx.x;       // 64 threads: (c, k)
    const double* p = part + static_cast<int64_t>(win) * nparts * 64 + t;
    double s = 0.0;
    for (int i = 0; i < nparts; ++i) s += p[static_cast<int64_t>(i) * 64];
    stats[static_cast<int64_t>(win) * 64 + t] = s;
}

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_shared_u4(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int T, int S>
__global__ void __launch_bounds__(kIsThreads, 1) conv_is_kernel(const IsArgs p) {
    static_assert(T * S * 32 <= 512, "accumulator ring exceeds TMEM");
    constexpr int R = 128 * T;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* wsm = smem;
    uint8_t* stages = smem + p.w_bytes;
    uint32_t* masks = reinterpret_cast<uint32_t*>(stages + static_cast<size_t>(p.nstages) * p.stage_bytes);   // [4][32]
    double* comb = reinterpret_cast<double*>(masks + 128);                                                   // [4][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(comb + 256);
    uint64_t* full = bars;                          // [stages]
    uint64_t* empty = bars + kIsMaxStages;          // [stages]
    uint64_t* tfull = bars + 2 * kIsMaxStages;      // [S]
    uint64_t* tempty = tfull + S;                   // [S]
    uint64_t* wfull = tempty + S;                   // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nplain = p.nchunks - p.xform_chunks;

    if (threadIdx.x == 0) {
        const uint32_t full_count = (nplain > 0 ? 1u : 0u) + (p.xform_chunks > 0 ? 4u : 0u);
        for (int s = 0; s < p.nstages; ++s) { mbar_init(&full[s], full_count); mbar_init(&empty[s], 1); }
        for (int s = 0; s < S; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
        mbar_init(wfull, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto item_geom = [&](int item, int& win, int& c, int& za, int& zb) {
        c = item % p.NC;
        const int wz = item / p.NC;
        const int zseg = wz % p.NZS;
        win = wz / p.NZS;
        za = 1 + zseg * p.Zs;
        zb = min(p.Z, za + p.Zs - 1);
    };

    if (warp == 0) {
        // ------------------------------------------------------------ producer: weights once, plain chunks per step
        if (lane == 0) {
            mbar_arrive_expect_tx(wfull, p.w_bytes);
            // bulk copies are limited in size only by the mbarrier tx count (2^20 - 1); split to be safe
            for (uint32_t off = 0; off < p.w_bytes; off += 27648u)
                tma_bulk_g2s(wsm + off, reinterpret_cast<const uint8_t*>(p.w) + off, min(27648u, p.w_bytes - off), wfull);
        }
        if (nplain > 0) {
            int stage = 0; uint32_t phase = 0;
            for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
                int win, c, za, zb;
                item_geom(item, win, c, za, zb);
                const int zi0 = max(za - 1, 1), zi1 = min(zb + 1, p.Z);
                for (int zi = zi0; zi <= zi1; ++zi) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (lane == 0) mbar_arrive_expect_tx(&full[stage], static_cast<uint32_t>(nplain) * p.RL * 16);
                    __syncwarp();
                    const int64_t pos = static_cast<int64_t>(p.in_guard) + static_cast<int64_t>(win) * p.Vp +
                                        static_cast<int64_t>(zi) * p.PL + c * R - p.H;
                    if (lane < nplain) {
                        const int chunk = p.xform_chunks + lane;
                        const __nv_bfloat16* base = (chunk < p.nch0) ? p.in0 + static_cast<int64_t>(chunk) * p.inS * 8
                                                                     : p.in1 + static_cast<int64_t>(chunk - p.nch0) * p.inS * 8;
                        tma_bulk_g2s(stages + static_cast<size_t>(stage) * p.stage_bytes + static_cast<size_t>(chunk) * p.RL * 16,
                                     base + pos * 8, p.RL * 16, &full[stage]);
                    }
                    if (++stage == p.nstages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        int stage = 0; uint32_t phase = 0;
        uint32_t slot_par = 0;       // per-slot use parity (bit s)
        mbar_wait(wfull, 0);
        const uint32_t w_addr = smem_u32(wsm);
        for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
            int win, c, za, zb;
            item_geom(item, win, c, za, zb);
            const int zi0 = max(za - 1, 1), zi1 = min(zb + 1, p.Z);
            for (int zi = zi0; zi <= zi1; ++zi) {
                const int lo = max(zi - 1, za), hi = min(zi + 1, zb);              // output planes fed by this input plane
                const int new_lo = (zi == zi0) ? lo : ((zi + 1 <= zb) ? zi + 1 : hi + 1);   // planes >= new_lo start here
                // the epilogue must have drained the slots that start a new plane
                for (int zo = new_lo; zo <= hi; ++zo) {
                    const int s = zo % S;
                    mbar_wait(&tempty[s], ((slot_par >> s) & 1u) ^ 1u);
                    slot_par ^= 1u << s;
                }
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one_sync()) {
                    // pieces: contiguous runs of output planes in slot order, at most two per range (ring wrap)
                    // {tmem column, B row offset (16 B units), instruction descriptor}; n == 0 -> absent
                    auto piece = [&](int a, int n, uint32_t& col, uint64_t& boff, uint32_t& idesc) {
                        col = tmem_base + (a % S) * 32;
                        boff = static_cast<uint64_t>((a - (zi - 1)) * 32);
                        idesc = umma_idesc_bf16_m128(32 * (n > 0 ? n : 1));
                    };
                    auto split = [&](int a, int b, int& n0, int& n1) {      // planes a..b -> n0 before the wrap, n1 after
                        const int cnt = b - a + 1;
                        n0 = cnt > 0 ? min(cnt, S - a % S) : 0;
                        n1 = cnt > 0 ? cnt - n0 : 0;
                    };
                    // every MMA but the first of the step accumulates into all planes lo..hi
                    int n0, n1;
                    split(lo, hi, n0, n1);
                    uint32_t c0, c1, i0, i1; uint64_t b0, b1;
                    piece(lo, n0, c0, b0, i0);
                    piece(lo + n0, n1, c1, b1, i1);
                    // the first MMA overwrites the planes that start at this step (>= new_lo) and accumulates into the others
                    int on0, on1, nn0, nn1;
                    split(lo, new_lo - 1, on0, on1);
                    split(new_lo, hi, nn0, nn1);
                    uint32_t oc0, oc1, oi0, oi1, nc0, nc1, ni0, ni1; uint64_t ob0, ob1, nb0, nb1;
                    piece(lo, on0, oc0, ob0, oi0);
                    piece(lo + on0, on1, oc1, ob1, oi1);
                    piece(new_lo, nn0, nc0, nb0, ni0);
                    piece(new_lo + nn0, nn1, nc1, nb1, ni1);
                    const uint32_t a_base = smem_u32(stages + static_cast<size_t>(stage) * p.stage_bytes);
                    const uint64_t adesc0 = umma_desc_kmajor_noswz(a_base, p.RL * 16, 128);
                    const uint64_t bdesc0 = umma_desc_kmajor_noswz(w_addr, 96 * 16, 128);
                    {
                        const uint64_t atap = adesc0 + static_cast<uint64_t>(p.H - p.Xp - 1);
#pragma unroll
                        for (int t = 0; t < T; ++t) {
                            if (on0) umma_bf16(oc0 + t * (S * 32), atap + t * 128, bdesc0 + ob0, oi0, 1u);
                            if (on1) umma_bf16(oc1 + t * (S * 32), atap + t * 128, bdesc0 + ob1, oi1, 1u);
                            if (nn0) umma_bf16(nc0 + t * (S * 32), atap + t * 128, bdesc0 + nb0, ni0, 0u);
                            if (nn1) umma_bf16(nc1 + t * (S * 32), atap + t * 128, bdesc0 + nb1, ni1, 0u);
                        }
                    }
#pragma unroll 1
                    for (int kb = 0; kb < p.KB; ++kb) {
#pragma unroll 1
                        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                if (kb == 0 && ky == 0 && kx == 0) continue;
                                const uint64_t atap = adesc0 + static_cast<uint64_t>(kb * 2 * p.RL + p.H + (ky - 1) * p.Xp + (kx - 1));
                                const uint64_t btap = bdesc0 + static_cast<uint64_t>(((kb * 9 + ky * 3 + kx) * 3072) >> 4);
#pragma unroll
                                for (int t = 0; t < T; ++t) {
                                    umma_bf16(c0 + t * (S * 32), atap + t * 128, btap + b0, i0, 1u);
                                    if (n1) umma_bf16(c1 + t * (S * 32), atap + t * 128, btap + b1, i1, 1u);
                                }
                            }
                        }
                    }
                    umma_commit(&empty[stage]);
                    // output planes whose last contributing input plane was this one are complete
                    for (int zo = lo; zo <= hi; ++zo)
                        if (min(zo + 1, zi1) == zi) umma_commit(&tfull[zo % S]);
                }
                __syncwarp();
                if (++stage == p.nstages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp < 6) {
        // ------------------------------------------------------------ epilogue (4 warps = 4 TMEM lane quadrants)
        const int q = warp & 3;
        uint32_t slot_par = 0;
        for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
            int win, c, za, zb;
            item_geom(item, win, c, za, zb);
            // in-plane decode of this thread's T rows (identical for every plane of the column)
            int64_t poff[T];
            bool valid[T];
            bool anyvalid = false;
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const int qq = c * R + t * 128 + q * 32 + lane;
                const int yp = qq / p.Xp, xp = qq - yp * p.Xp;
                valid[t] = qq < p.PL && yp >= 1 && yp <= p.Y && xp >= 1;
                poff[t] = static_cast<int64_t>(p.out_guard) + static_cast<int64_t>(win) * p.Vp + qq;
                anyvalid |= valid[t];
            }
            const unsigned anyw = __ballot_sync(0xffffffffu, anyvalid);
            double run_s = 0.0, run_q = 0.0;       // lane c: channel c
            for (int zo = za; zo <= zb; ++zo) {
                const int s = zo % S;
                mbar_wait(&tfull[s], (slot_par >> s) & 1u);
                slot_par ^= 1u << s;
                tc_fence_after();
                if (anyw) {
                    float acc_s[32], acc_q[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) { acc_s[i] = 0.f; acc_q[i] = 0.f; }
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        if (__ballot_sync(0xffffffffu, valid[t]) == 0) continue;
                        float v[32];
                        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * (S * 32) + s * 32, v);
                        if (valid[t]) {
                            __nv_bfloat16* o = p.out + (poff[t] + static_cast<int64_t>(zo) * p.PL) * 8;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                uint4 u;
                                u.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                                u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                                u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                                u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                                *reinterpret_cast<uint4*>(o + static_cast<int64_t>(j) * p.outS * 8) = u;
                            }
#pragma unroll
                            for (int i = 0; i < 32; ++i) { acc_s[i] += v[i]; acc_q[i] = fmaf(v[i], v[i], acc_q[i]); }
                        }
                    }
                    run_s += static_cast<double>(warp_transpose_sum32(acc_s));
                    run_q += static_cast<double>(warp_transpose_sum32(acc_q));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[s]);
            }
            // combine the four quadrants in a fixed order and write this item's partial sums
            comb[q * 64 + lane * 2] = run_s;
            comb[q * 64 + lane * 2 + 1] = run_q;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (q == 2) {     // warp 2 -> q = 2
                const int zseg = (item / p.NC) % p.NZS;
                double* dst = p.part + (static_cast<int64_t>(win) * p.nparts + zseg * p.NC + c) * 64;
                const double a = comb[lane * 2] + comb[64 + lane * 2] + comb[128 + lane * 2] + comb[192 + lane * 2];
                const double b = comb[lane * 2 + 1] + comb[64 + lane * 2 + 1] + comb[128 + lane * 2 + 1] + comb[192 + lane * 2 + 1];
                dst[lane * 2] = a;
                dst[lane * 2 + 1] = b;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
    } else if (p.xform_chunks > 0) {
        // ------------------------------------------------------------ transform warps: warp w stages chunk w
        const int chunk = warp - 6;
        uint32_t* mymask = masks + chunk * 32;
        const __nv_bfloat16* base = p.in0 + static_cast<int64_t>(chunk) * p.inS * 8;
        const int nit = (p.RL + 31) / 32;
        int stage = 0; uint32_t phase = 0;
        for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
            int win, c, za, zb;
            item_geom(item, win, c, za, zb);
            const int zi0 = max(za - 1, 1), zi1 = min(zb + 1, p.Z);
            // halo mask of the column's run (same for every plane)
            for (int it = 0; it < nit; ++it) {
                const int i = it * 32 + lane;
                const int qq = c * R - p.H + i;
                const int yp = qq / p.Xp, xp = qq - yp * p.Xp;
                const bool in = i < p.RL && qq >= 0 && qq < p.PL && yp >= 1 && yp <= p.Y && xp >= 1;
                const unsigned m = __ballot_sync(0xffffffffu, in);
                if (lane == 0) mymask[it] = m;
            }
            // InstanceNorm scale / shift of the producing layer for this window's 8 channels
            float a[8], b[8];
            {
                float ma = 0.f, mb = 0.f;
                if (lane < 8) {
                    const int ch = chunk * 8 + lane;
                    const double* st = p.in_stats + (static_cast<int64_t>(win) * 32 + ch) * 2;
                    const double mean = st[0] * p.inv_count;
                    double var = st[1] * p.inv_count - mean * mean;
                    var = var > 0.0 ? var : 0.0;
                    const double sc = static_cast<double>(p.in_gamma[ch]) / sqrt(var + 1e-5);
                    ma = static_cast<float>(sc);
                    mb = static_cast<float>(static_cast<double>(p.in_beta[ch]) - mean * sc);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) { a[i] = __shfl_sync(0xffffffffu, ma, i); b[i] = __shfl_sync(0xffffffffu, mb, i); }
            }
            __syncwarp();
            for (int zi = zi0; zi <= zi1; ++zi) {
                mbar_wait(&empty[stage], phase ^ 1);
                const int64_t pos = static_cast<int64_t>(p.in_guard) + static_cast<int64_t>(win) * p.Vp +
                                    static_cast<int64_t>(zi) * p.PL + c * R - p.H;
                const uint4* src = reinterpret_cast<const uint4*>(base + pos * 8);
                const uint32_t dst = smem_u32(stages + static_cast<size_t>(stage) * p.stage_bytes + static_cast<size_t>(chunk) * p.RL * 16);
                constexpr int U = 7;
                for (int it0 = 0; it0 < nit; it0 += U) {
                    uint4 u[U];
#pragma unroll
                    for (int k = 0; k < U; ++k) {
                        const int i = (it0 + k) * 32 + lane;
                        u[k] = (it0 + k < nit && i < p.RL) ? ld_nc_u4(src + i) : make_uint4(0u, 0u, 0u, 0u);
                    }
#pragma unroll
                    for (int k = 0; k < U; ++k) {
                        const int i = (it0 + k) * 32 + lane;
                        if (it0 + k < nit && i < p.RL) {
                            uint4 o = make_uint4(0u, 0u, 0u, 0u);
                            if ((mymask[it0 + k] >> lane) & 1u) {
                                o.x = pack_bf16x2(mish_f(fmaf(bf16_lo(u[k].x), a[0], b[0])), mish_f(fmaf(bf16_hi(u[k].x), a[1], b[1])));
                                o.y = pack_bf16x2(mish_f(fmaf(bf16_lo(u[k].y), a[2], b[2])), mish_f(fmaf(bf16_hi(u[k].y), a[3], b[3])));
                                o.z = pack_bf16x2(mish_f(fmaf(bf16_lo(u[k].z), a[4], b[4])), mish_f(fmaf(bf16_hi(u[k].z), a[5], b[5])));
                                o.w = pack_bf16x2(mish_f(fmaf(bf16_lo(u[k].w), a[6], b[6])), mish_f(fmaf(bf16_hi(u[k].w), a[7], b[7])));
                            }
                            st_shared_u4(dst + static_cast<uint32_t>(i) * 16u, o);
                        }
                    }
                }
                fence_proxy_async_smem();       // generic-proxy stores -> visible to the tensor core's async proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[stage]);
                if (++stage == p.nstages) { stage = 0; phase ^= 1; }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace dlv
