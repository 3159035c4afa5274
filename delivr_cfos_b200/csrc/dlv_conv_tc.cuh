// dlv_conv_tc.cuh - the tcgen05/TMEM implicit-GEMM kernel behind every dense contraction of the U-Net.
//
// Replaces the cuDNN conv3d / conv_transpose3d calls issued from predictor(window_data)
// (inference/sliding_window_inferer.py:222; MONAI BasicUNet instantiated at inference/inference.py:190-197).
//
// GEMM view (per work item):  D[128*T positions][NBLK couts] += A[positions][16 cin] * W[16 cin][NBLK]
// summed over KB cin blocks and NTAPS taps.
//   * A operand: the zero-haloed position-linear activation layout (dlv_internal.h, struct Level) is
//     staged once per cin block by TMA bulk copies as [dz run][k chunk][RL positions][8 ch] - which IS the
//     canonical K-major no-swizzle UMMA layout (8 positions x 16 B core matrices, SBO 128 B, LBO = RL*16 B).
//     A tap (dz,dy,dx) is therefore nothing but a +16 B * (dy*Xp + dx) bump of the descriptor start address
//     inside the dz run: 27 taps re-use one staged tile, nothing is re-fetched or im2col'ed.
//   * B operand: weights pre-packed on the host into the same canonical layout, one bulk copy per stage.
//   * D: fp32 accumulators in TMEM, T tiles x NBLK columns, double-buffered (2 x 256 columns) so the
//     epilogue of item i overlaps the MMAs of item i+1.
// Warp roles: one TMA producer warp, one MMA issuer warp (one thread), 4 epilogue warps
// (TMEM -> registers -> bf16 global stores + InstanceNorm partial sums).
// The k2s2 transposed convolutions are one N = 256 MMA per cin block and 512 B of output per input voxel; their epilogue
// is the critical path.  Measured in round 2 on one 128-window batch of cfg2 (profiles/r02_y_*, r02_z_*, r02_aa_*, r02_ab_*):
//   * eight epilogue warps and a software-pipelined tcgen05.ld: no gain (the next LDTM was waiting for the STOREs of the
//     previous block to drain, not for TMEM) - removed again;
//   * ncu --set full of the large (level 1 -> 0) deconv: L1/TEX 64 % busy, DRAM 46 %; each lane stored its 32 B as two
//     16 B halves, so a warp store covered 1 KB in half-used sectors.  With the halves exchanged by shuffles (XSTORE)
//     every store writes 512 contiguous bytes: 1.49 -> 1.39 ms, 248 -> 212 us on the next smaller one;
//   * shared-memory pipeline 8 stages deep instead of 2 (a stage is 12 KB and one MMA): 8-20 % off the three small
//     deconvs, which wait for TMA round trips, but the large one slows down (1.39 -> 1.48 ms: more 1 KB write streams in
//     flight) - the depth is chosen per launch (ConvArgs::nstages).
#pragma once
#include "dlv_common.cuh"

namespace dlv {

constexpr int kConvThreads = 192;
// warp roles: epilogue warps 0-3 (TMEM lane quadrant q = warp), TMA producer 4, MMA issuer 5 - the two latency-critical
// single-thread roles are the highest warp ids of their sub-partitions (the arbiter favours the highest eligible id)
#ifndef DLV_IS_LEGACY_ROLES
constexpr int kTcWarpProducer = 4, kTcWarpMma = 5;
#else
constexpr int kTcWarpProducer = 0, kTcWarpMma = 1;
#endif
constexpr int kConvStages = 2;         // 3x3x3 convs: a stage is one cin block of all 27 taps (~100 KB, ~7 000 clk of MMAs)
constexpr int kConvStagesMax = 12;     // k2s2 deconvs: a stage is 12 KB and ONE MMA - two of them in flight left the kernel waiting
                                       // for TMA round trips (2.9 us per 128-position tile); barriers for up to 12 stages fit the 256 B tail
constexpr int kTmemCols = 512;
constexpr uint32_t kSmemLimit = 232448;   // 227 KB opt-in maximum per CTA

enum ConvMode { kModeConvStats = 0, kModeDeconvScatter = 1 };

struct ConvArgs {
    const __nv_bfloat16* in0;   // first cin chunks (skip tensor for concatenated inputs)
    const __nv_bfloat16* in1;   // remaining chunks (upsampled tensor) or nullptr
    int nch0;                   // chunks held by in0
    int64_t inS;                // positions per chunk of the inputs
    int in_guard;
    const __nv_bfloat16* w;     // packed weights [KB][NB][NTAPS][2][NBLK][8]
    __nv_bfloat16* out;         // raw conv output (mode 0) / upsampled activation (mode 1)
    int64_t outS;
    int out_guard;
    double* stats;              // mode 0: [nwin][cout][2] sum, sum of squares
    const float* bias;          // mode 1: [cout]
    int Z, Y, X, Yp, Xp, YpXp, Vp;
    int64_t NP;                 // nwin * Vp positions to cover
    int KB, NB, T, RL, H;
    int nitems, items_per_cta;
    int nstages;                // smem pipeline depth (kConvStages for the convs, up to kConvStagesMax for the deconvs)
    uint32_t a_bytes, w_bytes, stage_bytes;
    int cout;
    int oYp, oXp, oVp;          // mode 1: geometry of the finer output level
};

template <int NBLK, int NTAPS, int MODE, bool XSTORE = false>
__global__ void __launch_bounds__(kConvThreads, 1) conv_tc_kernel(const ConvArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int NRUNS = (NTAPS == 27) ? 3 : 1;
    constexpr int kWarpProducer = kTcWarpProducer, kWarpMma = kTcWarpMma;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.nstages * p.stage_bytes);
    uint64_t* full = bars;                       // [stages] TMA -> MMA
    uint64_t* empty = bars + p.nstages;          // [stages] MMA -> TMA
    uint64_t* tfull = bars + 2 * p.nstages;      // [2] MMA -> epilogue
    uint64_t* tempty = tfull + 2;                // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.nstages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 4); }
        fence_mbar_init();
    }
    if (warp == kWarpMma) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int item0 = blockIdx.x * p.items_per_cta;
    const int item1 = min(p.nitems, item0 + p.items_per_cta);

    if (warp == kWarpProducer) {
        // ------------------------------------------------------------ TMA producer
        int stage = 0; uint32_t phase = 0;
        for (int item = item0; item < item1; ++item) {
            const int g = item / p.NB, nb = item - g * p.NB;
            const int64_t s = static_cast<int64_t>(g) * p.T * 128;
            for (int kb = 0; kb < p.KB; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* st = smem + stage * p.stage_bytes;
                if (lane == 0) mbar_arrive_expect_tx(&full[stage], p.a_bytes + p.w_bytes);
                __syncwarp();
                if (lane == 0) {
                    const __nv_bfloat16* wsrc = p.w + (static_cast<int64_t>(kb) * p.NB + nb) * (p.w_bytes / 2);
                    tma_bulk_g2s(st + p.a_bytes, wsrc, p.w_bytes, &full[stage]);
                } else if (lane <= NRUNS * 2) {
                    const int c = lane - 1, d = c >> 1, kc = c & 1;
                    const int chunk = kb * 2 + kc;
                    const __nv_bfloat16* base = (chunk < p.nch0) ? p.in0 + static_cast<int64_t>(chunk) * p.inS * 8
                                                                 : p.in1 + static_cast<int64_t>(chunk - p.nch0) * p.inS * 8;
                    const int64_t dzoff = (NTAPS == 27) ? static_cast<int64_t>(d - 1) * p.YpXp : 0;
                    const int64_t pos = p.in_guard + s + dzoff - p.H;
                    tma_bulk_g2s(st + static_cast<size_t>(c) * p.RL * 16, base + pos * 8, p.RL * 16, &full[stage]);
                }
                if (++stage == p.nstages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == kWarpMma) {
        // ------------------------------------------------------------ MMA issuer
        // The whole warp walks the pipeline (waits are warp-uniform); one elected lane issues the MMAs.
        constexpr uint32_t idesc = umma_idesc_bf16_m128(NBLK);
        int stage = 0; uint32_t phase = 0;
        int it = 0;
        for (int item = item0; item < item1; ++item, ++it) {
            const int buf = it & 1;
            mbar_wait(&tempty[buf], ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t dcol = tmem_base + buf * 256;
            for (int kb = 0; kb < p.KB; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t a_base = smem_u32(smem + stage * p.stage_bytes);
                    uint64_t bdesc = umma_desc_kmajor_noswz(a_base + p.a_bytes, NBLK * 16, 128);
                    const uint64_t adesc0 = umma_desc_kmajor_noswz(a_base, p.RL * 16, 128);
                    uint32_t acc = kb ? 1u : 0u;
                    if (NTAPS == 27) {
                        // descriptor start-address field is in 16 B units = positions: a tap is a +/- position bump
                        const int run = 2 * p.RL;
#pragma unroll 1
                        for (int kz = 0; kz < 3; ++kz) {
#pragma unroll 1
                            for (int ky = 0; ky < 3; ++ky) {
                                const uint64_t arow = adesc0 + static_cast<uint64_t>(kz * run + p.H + (ky - 1) * p.Xp - 1);
#pragma unroll
                                for (int kx = 0; kx < 3; ++kx) {
                                    const uint64_t atap = arow + kx;
#pragma unroll
                                    for (int t = 0; t < 8; ++t)
                                        if (t < p.T) umma_bf16(dcol + t * NBLK, atap + t * 128, bdesc, idesc, acc);
                                    acc = 1u;
                                    bdesc += (2 * NBLK * 16) >> 4;
                                }
                            }
                        }
                    } else {
#pragma unroll 1
                        for (int t = 0; t < p.T; ++t) umma_bf16(dcol + t * NBLK, adesc0 + t * 128, bdesc, idesc, acc);
                    }
                    umma_commit(&empty[stage]);
                    if (kb == p.KB - 1) umma_commit(&tfull[buf]);
                }
                __syncwarp();
                if (++stage == p.nstages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (4 warps = 4 TMEM lane quadrants)
        const int q = warp & 3;
        int it = 0;
        int cur_key = -1;           // win * NB + nb of the running statistics
        double run_s[NBLK / 32], run_q[NBLK / 32];
#pragma unroll
        for (int h = 0; h < NBLK / 32; ++h) { run_s[h] = 0.0; run_q[h] = 0.0; }
        auto flush = [&]() {
            if (MODE == kModeConvStats && cur_key >= 0) {
                const int win = cur_key / p.NB, nb = cur_key - win * p.NB;
#pragma unroll
                for (int h = 0; h < NBLK / 32; ++h) {
                    double* dst = p.stats + (static_cast<int64_t>(win) * p.cout + nb * NBLK + h * 32 + lane) * 2;
                    atomicAdd(dst, run_s[h]);
                    atomicAdd(dst + 1, run_q[h]);
                    run_s[h] = 0.0; run_q[h] = 0.0;
                }
            }
        };
        float bsr[MODE == kModeDeconvScatter ? 32 : 1];
        int bias_nb = -1;
        for (int item = item0; item < item1; ++item, ++it) {
            const int g = item / p.NB, nb = item - g * p.NB;
            const int64_t s = static_cast<int64_t>(g) * p.T * 128;
            const int buf = it & 1;
            if (MODE == kModeDeconvScatter && nb != bias_nb) {
                bias_nb = nb;
#pragma unroll
                for (int i = 0; i < (MODE == kModeDeconvScatter ? 32 : 1); ++i) bsr[i] = __ldg(p.bias + nb * 32 + i);
            }
            mbar_wait(&tfull[buf], (it >> 1) & 1);
            tc_fence_after();
            for (int t = 0; t < p.T; ++t) {
                const int64_t P = s + t * 128 + q * 32 + lane;
                // position -> (win, zp, yp, xp); interior test
                const int win = static_cast<int>(P / p.Vp);
                const int pp = static_cast<int>(P - static_cast<int64_t>(win) * p.Vp);
                const int zp = pp / p.YpXp;
                const int rr = pp - zp * p.YpXp;
                const int yp = rr / p.Xp;
                const int xp = rr - yp * p.Xp;
                const bool valid = (P < p.NP) && zp >= 1 && zp <= p.Z && yp >= 1 && yp <= p.Y && xp >= 1;
                const unsigned vmask = __ballot_sync(0xffffffffu, valid);
                if (vmask == 0) continue;   // warp-uniform: 32 halo positions, nothing to store
                if (MODE == kModeConvStats) {
                    const int w0 = __shfl_sync(0xffffffffu, win, __ffs(vmask) - 1);
                    const int key = w0 * p.NB + nb;
                    if (key != cur_key) { flush(); cur_key = key; }
                }
                if (MODE == kModeConvStats) {
#pragma unroll
                    for (int h = 0; h < NBLK / 32; ++h) {
                        float v[32];
                        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * 256 + t * NBLK + h * 32, v);
                        if (valid) {
                            __nv_bfloat16* o = p.out + (static_cast<int64_t>(nb * (NBLK / 8) + h * 4) * p.outS + p.out_guard + P) * 8;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                uint4 u;
                                u.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                                u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                                u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                                u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                                *reinterpret_cast<uint4*>(o + static_cast<int64_t>(j) * p.outS * 8) = u;
                            }
                        }
                        float sq[32];
#pragma unroll
                        for (int c = 0; c < 32; ++c) { v[c] = valid ? v[c] : 0.f; sq[c] = v[c] * v[c]; }
                        run_s[h] += static_cast<double>(warp_transpose_sum32(v));
                        run_q[h] += static_cast<double>(warp_transpose_sum32(sq));
                    }
                } else {
                    // transposed conv k2 s2: the N = 256 accumulator of this input voxel holds all 8 output
                    // sub-positions; block h = (a, b, jp) carries chunks j = 2 jp, 2 jp + 1, each as the x-even
                    // and x-odd output voxel side by side (pack_deconv) -> 2 x 16 B contiguous stores per chunk.
                    const uint32_t tcol = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * 256 + t * NBLK;
                    if (XSTORE) {
                        // Lane l holds the 32 B (x-even | x-odd output voxel) of input voxel l for every output row.  The halves
                        // are exchanged by shuffles so that every store instruction writes 512 contiguous bytes: instruction
                        // cp serves the input voxels of lanes 16 cp .. 16 cp + 15, lane l writes half (l & 1) of source
                        // lane 16 cp + l / 2 (see the header comment for the measurement).
                        const int half = lane & 1;
                        int64_t sbase[2];       // output position of source voxel's (a = 0, b = 0, x-even) corner, + half
                        bool svalid[2];
#pragma unroll
                        for (int cp = 0; cp < 2; ++cp) {
                            const int src = 16 * cp + (lane >> 1);
                            const int swin = __shfl_sync(0xffffffffu, win, src), szp = __shfl_sync(0xffffffffu, zp, src);
                            const int syp = __shfl_sync(0xffffffffu, yp, src), sxp = __shfl_sync(0xffffffffu, xp, src);
                            svalid[cp] = (vmask >> src) & 1u;
                            sbase[cp] = static_cast<int64_t>(swin) * p.oVp + (static_cast<int64_t>(2 * (szp - 1) + 1) * p.oYp + (2 * (syp - 1) + 1)) * p.oXp
                                        + (2 * (sxp - 1) + 1) + half;
                        }
#pragma unroll
                        for (int h = 0; h < NBLK / 32; ++h) {
                            float v[32];
                            tmem_ld32(tcol + h * 32, v);
                            const int a = h >> 2, b = (h >> 1) & 1, jp = h & 1;
                            const int64_t orow = (static_cast<int64_t>(a) * p.oYp + b) * p.oXp;
#pragma unroll
                            for (int jl = 0; jl < 2; ++jl) {
                                const int j = jp * 2 + jl;
                                const float* bb = bsr + j * 8;
                                uint32_t u[2][4];       // [x-even | x-odd][4 packed channel pairs] of this lane's input voxel
#pragma unroll
                                for (int c = 0; c < 2; ++c) {
                                    const float* vv = v + (jl * 2 + c) * 8;
#pragma unroll
                                    for (int w = 0; w < 4; ++w) u[c][w] = pack_bf16x2(vv[2 * w] + bb[2 * w], vv[2 * w + 1] + bb[2 * w + 1]);
                                }
                                __nv_bfloat16* orow_ptr = p.out + (static_cast<int64_t>(nb * 4 + j) * p.outS + p.out_guard + orow) * 8;
#pragma unroll
                                for (int cp = 0; cp < 2; ++cp) {
                                    const int src = 16 * cp + (lane >> 1);
                                    uint32_t x[4];
#pragma unroll
                                    for (int w = 0; w < 4; ++w) {
                                        const uint32_t e = __shfl_sync(0xffffffffu, u[0][w], src), o = __shfl_sync(0xffffffffu, u[1][w], src);
                                        x[w] = half ? o : e;
                                    }
                                    if (svalid[cp]) *reinterpret_cast<uint4*>(orow_ptr + sbase[cp] * 8) = make_uint4(x[0], x[1], x[2], x[3]);
                                }
                            }
                        }
                    } else {
                    const int64_t Pw = static_cast<int64_t>(win) * p.oVp;
#pragma unroll
                    for (int h = 0; h < NBLK / 32; ++h) {
                        float v[32];
                        tmem_ld32(tcol + h * 32, v);
                        if (valid) {
                            const int a = h >> 2, b = (h >> 1) & 1, jp = h & 1;
                            const int oz = 2 * (zp - 1) + a + 1;
                            const int oy = 2 * (yp - 1) + b + 1;
                            const int ox = 2 * (xp - 1) + 1;
                            const int64_t Po = Pw + (static_cast<int64_t>(oz) * p.oYp + oy) * p.oXp + ox;
#pragma unroll
                            for (int jl = 0; jl < 2; ++jl) {
                                const int j = jp * 2 + jl;
                                __nv_bfloat16* o = p.out + (static_cast<int64_t>(nb * 4 + j) * p.outS + p.out_guard + Po) * 8;
#pragma unroll
                                for (int c = 0; c < 2; ++c) {
                                    const float* vv = v + (jl * 2 + c) * 8;
                                    const float* bb = bsr + j * 8;
                                    uint4 u;
                                    u.x = pack_bf16x2(vv[0] + bb[0], vv[1] + bb[1]);
                                    u.y = pack_bf16x2(vv[2] + bb[2], vv[3] + bb[3]);
                                    u.z = pack_bf16x2(vv[4] + bb[4], vv[5] + bb[5]);
                                    u.w = pack_bf16x2(vv[6] + bb[6], vv[7] + bb[7]);
                                    *reinterpret_cast<uint4*>(o + c * 8) = u;
                                }
                            }
                        }
                    }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
        }
        flush();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kWarpMma) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace dlv
