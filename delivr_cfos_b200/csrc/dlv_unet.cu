// dlv_unet.cu - the 3-D U-Net of DELiVR's blob detector on B200: window gather, tcgen05 convolutions,
// InstanceNorm+Mish(+MaxPool) passes, transposed convolutions, final 1x1 conv fused with the overlap blend.
//
// Network: MONAI 1.2.0 BasicUNet(spatial_dims=3, in=1, out=1, features=(32,32,64,128,256,32), act="mish",
// norm=instance(affine), upsample="deconv") as instantiated at inference/inference.py:190-197 of the reference;
// forward order x0=conv_0(x); x1..x4=down_1..4; u4=upcat_4(x4,x3) ... u1=upcat_1(u2,x0); final_conv(u1).
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <cmath>

#include "dlv_conv_is.cuh"
#include "dlv_conv_tc.cuh"
#include "dlv_internal.h"

namespace dlv {

// =================================================================== elementwise kernels
struct LevelDev {
    int Z, Y, X, Yp, Xp, YpXp, Vp, guard;
    int64_t S;
};
static LevelDev to_dev(const Level& L) { return LevelDev{L.Z, L.Y, L.X, L.Yp, L.Xp, L.YpXp, L.Vp, L.guard, L.S}; }

__device__ __forceinline__ int64_t pos_of(const LevelDev& L, int win, int z, int y, int x) {
    return static_cast<int64_t>(L.guard) + static_cast<int64_t>(win) * L.Vp + (static_cast<int64_t>(z + 1) * L.Yp + (y + 1)) * L.Xp + (x + 1);
}

// uint16 window gather (sliding_window_inferer.py:181-195,207, flips :218-219) fused with the operand format of
// the first convolution.  A uint16 v is split as v = hi + lo, hi = v & 0xFF00, lo = v & 0xFF - both exact in
// bf16 - and the first-layer weights as W = Wh + Wl (two bf16): {hi, lo, hi, lo} against {Wh, Wh, Wl, Wl}
// reproduces the fp32 product v*W to ~2^-16 relative on the tensor cores.  The three kx neighbours of a voxel are
// folded into the K = 16 slot as well (channel = kx*4 + term; zero outside the window = the conv's zero padding),
// so the first layer needs only its centre-kx taps.
__global__ void gather_windows_kernel(const uint16_t* __restrict__ slab, int64_t slabY, int64_t slabX,
                                      const WindowDesc* __restrict__ wd, LevelDev L, __nv_bfloat16* __restrict__ in0) {
    const int win = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L.Z * L.Y * L.X) return;
    const int x = idx % L.X, y = (idx / L.X) % L.Y, z = idx / (L.X * L.Y);
    const WindowDesc w = wd[win];
    const int flip = w.flip & 0xFF;
    const int zs = (flip == 1) ? L.Z - 1 - z : z;
    const int ys = (flip == 2) ? L.Y - 1 - y : y;
    const uint16_t* row = slab + (static_cast<int64_t>(w.oz + zs) * slabY + (w.oy + ys)) * slabX + w.ox;
    uint32_t t[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int xw = x + k - 1;
        uint32_t v = 0u;
        if (xw >= 0 && xw < L.X) v = row[(flip == 3) ? L.X - 1 - xw : xw];
        t[k] = pack_bf16x2(static_cast<float>(v & 0xFF00u), static_cast<float>(v & 0xFFu));
    }
    const int64_t P = pos_of(L, win, z, y, x);
    *reinterpret_cast<uint4*>(in0 + P * 8) = make_uint4(t[0], t[0], t[1], t[1]);
    *reinterpret_cast<uint4*>(in0 + (L.S + P) * 8) = make_uint4(t[2], t[2], 0u, 0u);
}

// per-window input maximum > 0 test (skip rule, sliding_window_inferer.py:198; applied per window)
__global__ void window_active_kernel(const uint16_t* __restrict__ slab, int64_t slabY, int64_t slabX,
                                     const int32_t* __restrict__ origins, int rz, int ry, int rx, int32_t* __restrict__ active) {
    const int win = blockIdx.x;
    const int oz = origins[3 * win], oy = origins[3 * win + 1], ox = origins[3 * win + 2];
    int any = 0;
    if (((ox | rx) & 7) == 0 && (slabX & 7) == 0) {
        // 16 B loads (8 voxels), 8 independent loads per thread in flight; block-wide early exit every round
        const int rx8 = rx >> 3, n8 = rz * ry * rx8;
        for (int base = 0; base < n8; base += 8 * blockDim.x) {
            uint4 acc = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int idx = base + k * blockDim.x + threadIdx.x;
                if (idx < n8) {
                    const int x8 = idx % rx8, y = (idx / rx8) % ry, z = idx / (rx8 * ry);
                    const uint4 u = ld_nc_u4(slab + (static_cast<int64_t>(oz + z) * slabY + (oy + y)) * slabX + ox + 8 * x8);
                    acc.x |= u.x; acc.y |= u.y; acc.z |= u.z; acc.w |= u.w;
                }
            }
            if (__syncthreads_or((acc.x | acc.y | acc.z | acc.w) != 0u)) { any = 1; break; }
        }
    } else {
        const int n = rz * ry * rx;
        for (int idx = threadIdx.x; idx < n && !any; idx += blockDim.x) {
            const int x = idx % rx, y = (idx / rx) % ry, z = idx / (rx * ry);
            any |= slab[(static_cast<int64_t>(oz + z) * slabY + (oy + y)) * slabX + (ox + x)] > 0;
        }
        any = __syncthreads_or(any);
    }
    if (threadIdx.x == 0) active[win] = any;
}

// InstanceNorm3d(affine, eps 1e-5, biased variance) + Mish (+ MaxPool3d(2)) on the raw bf16 conv output.
// grid.y = win * nchunk + chunk; one thread = one position (POOL: one 2x2x2 block) x 8 channels (16 B).
template <bool POOL>
__global__ void norm_mish_kernel(const __nv_bfloat16* __restrict__ raw, LevelDev L, __nv_bfloat16* __restrict__ out,
                                 __nv_bfloat16* __restrict__ pooled, LevelDev Lp, const double* __restrict__ stats,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, int C, int nchunk) {
    __shared__ float sa[8], sb[8];
    const int win = blockIdx.y / nchunk, chunk = blockIdx.y - win * nchunk;
    if (threadIdx.x < 8) {
        const int c = chunk * 8 + threadIdx.x;
        const double inv = 1.0 / (static_cast<double>(L.Z) * L.Y * L.X);
        const double mean = stats[(static_cast<int64_t>(win) * C + c) * 2] * inv;
        double var = stats[(static_cast<int64_t>(win) * C + c) * 2 + 1] * inv - mean * mean;
        var = var > 0.0 ? var : 0.0;
        const double a = static_cast<double>(gamma[c]) / sqrt(var + 1e-5);
        sa[threadIdx.x] = static_cast<float>(a);
        sb[threadIdx.x] = static_cast<float>(static_cast<double>(beta[c]) - mean * a);
    }
    __syncthreads();
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = sa[i]; b[i] = sb[i]; }
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t cbase = static_cast<int64_t>(chunk) * L.S * 8;
    auto apply = [&](int64_t P, float (&m)[8]) {
        const uint4 u = *reinterpret_cast<const uint4*>(raw + cbase + P * 8);
        float f[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y), bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = mish_fast(fmaf(f[i], a[i], b[i]));
        uint4 o = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        if (out) *reinterpret_cast<uint4*>(out + cbase + P * 8) = o;
        // pooling compares the bf16-rounded activations (what the next layer would read)
        m[0] = fmaxf(m[0], bf16_lo(o.x)); m[1] = fmaxf(m[1], bf16_hi(o.x));
        m[2] = fmaxf(m[2], bf16_lo(o.y)); m[3] = fmaxf(m[3], bf16_hi(o.y));
        m[4] = fmaxf(m[4], bf16_lo(o.z)); m[5] = fmaxf(m[5], bf16_hi(o.z));
        m[6] = fmaxf(m[6], bf16_lo(o.w)); m[7] = fmaxf(m[7], bf16_hi(o.w));
    };
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
    if (!POOL) {
        if (idx >= L.Z * L.Y * L.X) return;
        const int x = idx % L.X, y = (idx / L.X) % L.Y, z = idx / (L.X * L.Y);
        apply(pos_of(L, win, z, y, x), m);
    } else if (!out) {
        // Pool only (the full-resolution activation is re-derived by the consuming fused conv): mish is decreasing
        // left of its minimum and increasing right of it, so max over the 2x2x2 block of mish(a*x+b) is
        // max(mish(a*max x + b), mish(a*min x + b)) - two activations per channel instead of eight.
        const int X2 = L.X / 2, Y2 = L.Y / 2, Z2 = L.Z / 2;
        if (idx >= X2 * Y2 * Z2) return;
        const int x = idx % X2, y = (idx / X2) % Y2, z = idx / (X2 * Y2);
        uint4 mx, mn;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint4 u = ld_nc_u4(raw + cbase + pos_of(L, win, 2 * z + (k >> 2), 2 * y + ((k >> 1) & 1), 2 * x + (k & 1)) * 8);
            if (k == 0) { mx = u; mn = u; }
            else {
                mx.x = bf16x2_max(mx.x, u.x); mx.y = bf16x2_max(mx.y, u.y); mx.z = bf16x2_max(mx.z, u.z); mx.w = bf16x2_max(mx.w, u.w);
                mn.x = bf16x2_min(mn.x, u.x); mn.y = bf16x2_min(mn.y, u.y); mn.z = bf16x2_min(mn.z, u.z); mn.w = bf16x2_min(mn.w, u.w);
            }
        }
        f32x2 a2[4], b2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { a2[i] = pk2(a[2 * i], a[2 * i + 1]); b2[i] = pk2(b[2 * i], b[2 * i + 1]); }
        const uint4 hi = norm_mish8<0>(mx, a2, b2), lo = norm_mish8<0>(mn, a2, b2);
        const uint4 o = make_uint4(bf16x2_max(hi.x, lo.x), bf16x2_max(hi.y, lo.y), bf16x2_max(hi.z, lo.z), bf16x2_max(hi.w, lo.w));
        *reinterpret_cast<uint4*>(pooled + static_cast<int64_t>(chunk) * Lp.S * 8 + pos_of(Lp, win, z, y, x) * 8) = o;
    } else {
        const int X2 = L.X / 2, Y2 = L.Y / 2, Z2 = L.Z / 2;
        if (idx >= X2 * Y2 * Z2) return;
        const int x = idx % X2, y = (idx / X2) % Y2, z = idx / (X2 * Y2);
#pragma unroll
        for (int dz = 0; dz < 2; ++dz)
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) apply(pos_of(L, win, 2 * z + dz, 2 * y + dy, 2 * x + dx), m);
        uint4 o = make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]), pack_bf16x2(m[4], m[5]), pack_bf16x2(m[6], m[7]));
        *reinterpret_cast<uint4*>(pooled + static_cast<int64_t>(chunk) * Lp.S * 8 + pos_of(Lp, win, z, y, x) * 8) = o;
    }
}

// Last layer: InstanceNorm + Mish of upcat_1.convs.conv_1, final 1x1x1 conv (32 -> 1, + bias) and the overlap
// blend (sliding_window_inferer.py:222-251): acc[voxel] += w * logit, un-flipped.
// One thread = one window voxel; 4 coalesced 16 B loads; coalesced 4 B integer reductions into the accumulator.
__global__ void final_blend_kernel(const __nv_bfloat16* __restrict__ raw, LevelDev L, const double* __restrict__ stats,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ fw, float fb, const WindowDesc* __restrict__ wd,
                                   int32_t* __restrict__ acc, int64_t slabY, int64_t slabX, const BlendDev bw,
                                   float* __restrict__ logits_out) {
    __shared__ f32x2 sa[16], sb[16], sw[16];
    const int win = blockIdx.y;
    if (threadIdx.x < 32) {
        const int c = threadIdx.x;
        const double inv = 1.0 / (static_cast<double>(L.Z) * L.Y * L.X);
        const double mean = stats[(static_cast<int64_t>(win) * 32 + c) * 2] * inv;
        double var = stats[(static_cast<int64_t>(win) * 32 + c) * 2 + 1] * inv - mean * mean;
        var = var > 0.0 ? var : 0.0;
        const double a = static_cast<double>(gamma[c]) / sqrt(var + 1e-5);
        reinterpret_cast<float*>(sa)[c] = static_cast<float>(a);
        reinterpret_cast<float*>(sb)[c] = static_cast<float>(static_cast<double>(beta[c]) - mean * a);
        reinterpret_cast<float*>(sw)[c] = fw[c];
    }
    __syncthreads();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L.Z * L.Y * L.X) return;
    const int x = idx % L.X, y = (idx / L.X) % L.Y, z = idx / (L.X * L.Y);
    const int64_t P = pos_of(L, win, z, y, x);
    uint4 u[4];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) u[ch] = ld_nc_u4(raw + static_cast<int64_t>(ch) * L.S * 8 + P * 8);
    f32x2 dot = 0ull;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        f32x2 a2[4], b2[4], m[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { a2[i] = sa[ch * 4 + i]; b2[i] = sb[ch * 4 + i]; }
        norm_mish8_f32<4>(u[ch], a2, b2, m);      // reciprocals on the FMA pipe: one MUFU per element
#pragma unroll
        for (int i = 0; i < 4; ++i) dot = fma2(sw[ch * 4 + i], m[i], dot);
    }
    float d0, d1;
    upk2(dot, d0, d1);
    const float logit = fb + d0 + d1;
    if (logits_out) {   // operator-level entry point: plain per-window logits, window order
        logits_out[static_cast<int64_t>(win) * L.Z * L.Y * L.X + idx] = logit;
        return;
    }
    const WindowDesc w = wd[win];
    const int flip = w.flip & 0xFF, repeat = (w.flip >> 8) + 1;
    const int zo = (flip == 1) ? L.Z - 1 - z : z;
    const int yo = (flip == 2) ? L.Y - 1 - y : y;
    const int xo = (flip == 3) ? L.X - 1 - x : x;
    const float wgt = bw.wz ? (bw.wz[zo] * bw.nz[w.oz + zo]) * (bw.wy[yo] * bw.ny[w.oy + yo]) * (bw.wx[xo] * bw.nx[w.ox + xo]) : 1.f;
    // fixed-point accumulation (2^-12 logit units): integer adds are associative, so the blended sum is
    // bit-identical for any window order, batch composition or slab partition across GPUs.  `repeat` identical
    // passes (test-time augmentation evaluates the same flip several times) are one pass added `repeat` times.
    const float v = fminf(fmaxf(wgt * logit, -kAccClamp), kAccClamp);
    atomicAdd(acc + (static_cast<int64_t>(w.oz + zo) * slabY + (w.oy + yo)) * slabX + (w.ox + xo),
              __float2int_rn(v * kAccScale) * repeat);
    (void)xo;
}

// NCDHW fp32 <-> haloed chunked bf16 layout (operator-level entry points / tests only)
__global__ void pack_act_kernel(const float* __restrict__ x, int C, LevelDev L, __nv_bfloat16* __restrict__ out, int nchunk) {
    const int win = blockIdx.y / nchunk, chunk = blockIdx.y - win * nchunk;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = L.Z * L.Y * L.X;
    if (idx >= n) return;
    const int xx = idx % L.X, y = (idx / L.X) % L.Y, z = idx / (L.X * L.Y);
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = chunk * 8 + i;
        f[i] = c < C ? x[(static_cast<int64_t>(win) * C + c) * n + idx] : 0.f;
    }
    uint4 o = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    *reinterpret_cast<uint4*>(out + static_cast<int64_t>(chunk) * L.S * 8 + pos_of(L, win, z, y, xx) * 8) = o;
}
__global__ void unpack_act_kernel(const __nv_bfloat16* __restrict__ in, int C, LevelDev L, float* __restrict__ y_out, int nchunk) {
    const int win = blockIdx.y / nchunk, chunk = blockIdx.y - win * nchunk;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = L.Z * L.Y * L.X;
    if (idx >= n) return;
    const int xx = idx % L.X, y = (idx / L.X) % L.Y, z = idx / (L.X * L.Y);
    const uint4 u = *reinterpret_cast<const uint4*>(in + static_cast<int64_t>(chunk) * L.S * 8 + pos_of(L, win, z, y, xx) * 8);
    const float f[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y), bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = chunk * 8 + i;
        if (c < C) y_out[(static_cast<int64_t>(win) * C + c) * n + idx] = f[i];
    }
}

// =================================================================== host: levels, weights, engine
static Level make_level(int Z, int Y, int X, int batch) {
    Level L;
    L.Z = Z; L.Y = Y; L.X = X;
    L.Zp = Z + 2; L.Yp = Y + 2; L.Xp = X + 1;
    L.YpXp = L.Yp * L.Xp;
    // window stride in positions, rounded to whole 128-row MMA tiles: every window then sees the same tile /
    // warp-tile boundaries whatever its slot in the batch, so its InstanceNorm partial sums (fp32 per 32 rows)
    // are grouped identically -> per-window results do not depend on batch composition or slab partition.
    L.Vp = ((L.Zp * L.YpXp + 127) / 128) * 128;
    L.guard = ((L.YpXp + L.Xp + 1 + 7) / 8) * 8;
    const int64_t np = static_cast<int64_t>(batch) * L.Vp;
    L.S = L.guard + ((np + 1023) / 1024) * 1024 + L.guard + 1024 + 64;
    return L;
}

static bf16 to_bf16(float f) { return __float2bfloat16_rn(f); }

static int upload(Ctx* ctx, const void* host, size_t bytes, void** dev) {
    DLV_CUDA_OK(ctx, cudaMalloc(dev, bytes));
    DLV_CUDA_OK(ctx, cudaMemcpy(*dev, host, bytes, cudaMemcpyHostToDevice));
    return 0;
}

static int pack_conv(Ctx* ctx, ConvLayer& L, const float* W, bool first) {
    // W[cout][cin][3][3][3] (torch Conv3d) -> [KB][NB][27][2][nblk][8] bf16
    L.ntaps = 27;
    L.cin_pad = first ? 16 : ((L.cin + 15) / 16) * 16;
    L.KB = L.cin_pad / 16;
    L.nblk = (L.cout >= 64) ? 64 : 32;
    L.NB = L.cout / L.nblk;
    std::vector<bf16> pk(static_cast<size_t>(L.KB) * L.NB * 27 * 2 * L.nblk * 8, to_bf16(0.f));
    for (int kb = 0; kb < L.KB; ++kb)
        for (int nb = 0; nb < L.NB; ++nb)
            for (int tap = 0; tap < 27; ++tap)
                for (int kc = 0; kc < 2; ++kc)
                    for (int n = 0; n < L.nblk; ++n)
                        for (int e = 0; e < 8; ++e) {
                            const int ci = kb * 16 + kc * 8 + e, co = nb * L.nblk + n;
                            float v = 0.f;
                            if (first) {
                                // kx folded into K (gather_windows_kernel): only the centre-kx taps carry weights
                                if (ci < 12 && tap % 3 == 1) {
                                    const float w = W[static_cast<size_t>(co) * 27 + (tap - 1) + ci / 4];
                                    const float wh = __bfloat162float(to_bf16(w));
                                    v = (ci % 4 < 2) ? wh : (w - wh);
                                }
                            } else if (ci < L.cin) {
                                v = W[(static_cast<size_t>(co) * L.cin + ci) * 27 + tap];
                            }
                            pk[((((static_cast<size_t>(kb) * L.NB + nb) * 27 + tap) * 2 + kc) * L.nblk + n) * 8 + e] = to_bf16(v);
                        }
    return upload(ctx, pk.data(), pk.size() * sizeof(bf16), reinterpret_cast<void**>(&L.w));
}

// Cout = 32 layers, input-stationary kernel: W[32][cin][3][3][3] -> [KB][9 (ky,kx)][2][96][8]; the 96 B rows are the
// three kz blocks in the order of the output planes they feed (z-1: kz = 2, z: kz = 1, z+1: kz = 0).
static int pack_conv_is(Ctx* ctx, ConvLayer& L, const float* W, bool first) {
    const int ntap = first ? 3 : 9;        // first layer: kx folded into K, taps = ky only
    std::vector<bf16> pk(static_cast<size_t>(L.KB) * ntap * 2 * 96 * 8, to_bf16(0.f));
    for (int kb = 0; kb < L.KB; ++kb)
        for (int tap = 0; tap < ntap; ++tap)
            for (int kc = 0; kc < 2; ++kc)
                for (int row = 0; row < 96; ++row)
                    for (int e = 0; e < 8; ++e) {
                        const int kz = 2 - row / 32, co = row % 32, ci = kb * 16 + kc * 8 + e;
                        float v = 0.f;
                        if (first) {
                            if (ci < 12) {
                                const float w = W[static_cast<size_t>(co) * 27 + kz * 9 + tap * 3 + ci / 4];
                                const float wh = __bfloat162float(to_bf16(w));
                                v = (ci % 4 < 2) ? wh : (w - wh);
                            }
                        } else if (ci < L.cin) {
                            v = W[(static_cast<size_t>(co) * L.cin + ci) * 27 + kz * 9 + tap];
                        }
                        pk[((((static_cast<size_t>(kb) * ntap + tap) * 2 + kc) * 96 + row) * 8) + e] = to_bf16(v);
                    }
    return upload(ctx, pk.data(), pk.size() * sizeof(bf16), reinterpret_cast<void**>(&L.w_is));
}

static int pack_deconv(Ctx* ctx, ConvLayer& L, const float* W) {
    // W[cin][cout][2][2][2] (torch ConvTranspose3d) -> [KB][NB = cout/32][1][2][256][8]: one N = 256 MMA produces
    // all 8 sub-positions of 32 output channels for 128 input voxels.  Column order inside the 256:
    // n = ((a*2 + b)*2 + jp)*32 + (jl*2 + c)*8 + e  with output sub-position (a, b, c), output channel
    // co = nb*32 + (jp*2 + jl)*8 + e: a 32-column accumulator block holds, for two 8-channel chunks, the x-even and
    // the x-odd output voxel side by side, so the epilogue stores 32 contiguous bytes per chunk (full sectors).
    L.ntaps = 1;
    L.cin_pad = L.cin;
    L.KB = L.cin / 16;
    L.nblk = 256;
    L.NB = L.cout / 32;
    std::vector<bf16> pk(static_cast<size_t>(L.KB) * L.NB * 2 * L.nblk * 8);
    for (int kb = 0; kb < L.KB; ++kb)
        for (int nb = 0; nb < L.NB; ++nb)
            for (int kc = 0; kc < 2; ++kc)
                for (int n = 0; n < L.nblk; ++n)
                    for (int e = 0; e < 8; ++e) {
                        const int ci = kb * 16 + kc * 8 + e;
                        const int blk = n / 32, r = n % 32;
                        const int a = blk >> 2, b = (blk >> 1) & 1, jp = blk & 1;
                        const int jl = r >> 4, c = (r >> 3) & 1, ce = r & 7;
                        const int abc = a * 4 + b * 2 + c, co = nb * 32 + (jp * 2 + jl) * 8 + ce;
                        const float v = W[(static_cast<size_t>(ci) * L.cout + co) * 8 + abc];
                        pk[(((static_cast<size_t>(kb) * L.NB + nb) * 2 + kc) * L.nblk + n) * 8 + e] = to_bf16(v);
                    }
    return upload(ctx, pk.data(), pk.size() * sizeof(bf16), reinterpret_cast<void**>(&L.w));
}

struct ConvSpec { const char* name; int cin, cout; };
static const ConvSpec kConvs[] = {
    {"conv_0.conv_0", 1, 32},          {"conv_0.conv_1", 32, 32},
    {"down_1.convs.conv_0", 32, 32},   {"down_1.convs.conv_1", 32, 32},
    {"down_2.convs.conv_0", 32, 64},   {"down_2.convs.conv_1", 64, 64},
    {"down_3.convs.conv_0", 64, 128},  {"down_3.convs.conv_1", 128, 128},
    {"down_4.convs.conv_0", 128, 256}, {"down_4.convs.conv_1", 256, 256},
    {"upcat_4.convs.conv_0", 256, 128}, {"upcat_4.convs.conv_1", 128, 128},
    {"upcat_3.convs.conv_0", 128, 64},  {"upcat_3.convs.conv_1", 64, 64},
    {"upcat_2.convs.conv_0", 64, 32},   {"upcat_2.convs.conv_1", 32, 32},
    {"upcat_1.convs.conv_0", 64, 32},   {"upcat_1.convs.conv_1", 32, 32},
};
static const ConvSpec kDeconvs[] = {{"upcat_4", 256, 128}, {"upcat_3", 128, 64}, {"upcat_2", 64, 32}, {"upcat_1", 32, 32}};

void net_free(Ctx* ctx) {
    for (auto* m : {&ctx->net.conv, &ctx->net.deconv})
        for (auto& kv : *m) {
            cudaFree(kv.second.w); cudaFree(kv.second.w_is); cudaFree(kv.second.gamma); cudaFree(kv.second.beta); cudaFree(kv.second.bias);
        }
    ctx->net.conv.clear();
    ctx->net.deconv.clear();
    cudaFree(ctx->net.final_w);
    ctx->net.final_w = nullptr;
    ctx->net.loaded = false;
}

int net_load(Ctx* ctx, int n, const char* const* names, const float* const* data, const int64_t* numel) {
    net_free(ctx);
    std::map<std::string, std::pair<const float*, int64_t>> sd;
    for (int i = 0; i < n; ++i) {
        std::string k = names[i];
        if (k.rfind("module.", 0) == 0) k = k.substr(7);
        sd[k] = {data[i], numel[i]};
    }
    auto need = [&](const std::string& k, int64_t ne, const float** out) -> int {
        auto it = sd.find(k);
        if (it == sd.end()) { set_error(ctx, "dlv_load_weights: missing tensor '%s' (strict load)", k.c_str()); return DLV_ERR_ARG; }
        if (it->second.second != ne) {
            set_error(ctx, "dlv_load_weights: tensor '%s' has %lld elements, expected %lld", k.c_str(), (long long)it->second.second, (long long)ne);
            return DLV_ERR_ARG;
        }
        *out = it->second.first;
        sd.erase(it);
        return 0;
    };
    int rc;
    for (const ConvSpec& s : kConvs) {
        ConvLayer L;
        L.name = s.name; L.cin = s.cin; L.cout = s.cout;
        const float *w, *b, *g, *be;
        if ((rc = need(L.name + ".conv.weight", static_cast<int64_t>(s.cout) * s.cin * 27, &w))) return rc;
        if ((rc = need(L.name + ".conv.bias", s.cout, &b))) return rc;   // cancelled exactly by InstanceNorm: unused
        if ((rc = need(L.name + ".adn.N.weight", s.cout, &g))) return rc;
        if ((rc = need(L.name + ".adn.N.bias", s.cout, &be))) return rc;
        if ((rc = pack_conv(ctx, L, w, s.cin == 1))) return rc;
        if (s.cout == 32 && (rc = pack_conv_is(ctx, L, w, s.cin == 1))) return rc;
        if ((rc = upload(ctx, g, s.cout * sizeof(float), reinterpret_cast<void**>(&L.gamma)))) return rc;
        if ((rc = upload(ctx, be, s.cout * sizeof(float), reinterpret_cast<void**>(&L.beta)))) return rc;
        ctx->net.conv[L.name] = L;
    }
    for (const ConvSpec& s : kDeconvs) {
        ConvLayer L;
        L.name = s.name; L.cin = s.cin; L.cout = s.cout;
        const float *w, *b;
        if ((rc = need(L.name + ".upsample.deconv.weight", static_cast<int64_t>(s.cin) * s.cout * 8, &w))) return rc;
        if ((rc = need(L.name + ".upsample.deconv.bias", s.cout, &b))) return rc;
        if ((rc = pack_deconv(ctx, L, w))) return rc;
        if ((rc = upload(ctx, b, s.cout * sizeof(float), reinterpret_cast<void**>(&L.bias)))) return rc;
        ctx->net.deconv[L.name] = L;
    }
    const float *fw, *fb;
    if ((rc = need("final_conv.weight", 32, &fw))) return rc;
    if ((rc = need("final_conv.bias", 1, &fb))) return rc;
    if ((rc = upload(ctx, fw, 32 * sizeof(float), reinterpret_cast<void**>(&ctx->net.final_w)))) return rc;
    ctx->net.final_b = fb[0];
    if (!sd.empty()) { set_error(ctx, "dlv_load_weights: unexpected tensor '%s' (strict load)", sd.begin()->first.c_str()); return DLV_ERR_ARG; }
    ctx->net.loaded = true;
    return 0;
}

// ------------------------------------------------------------------- conv launcher
template <int NBLK, int NTAPS, int MODE, bool XSTORE = false>
static int launch_conv_t(Ctx* ctx, const ConvArgs& a, int grid, uint32_t smem) {
    auto k = conv_tc_kernel<NBLK, NTAPS, MODE, XSTORE>;
    static bool attr_set = false;
    if (!attr_set) {
        DLV_CUDA_OK(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
        attr_set = true;
    }
    k<<<grid, kConvThreads, smem, ctx->stream>>>(a);
    ctx->launches++;
    DLV_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

// Runs one conv / deconv layer over `nwin` windows.  in1 may be nullptr.
static int run_conv(Ctx* ctx, const ConvLayer& Ly, const Level& L, int nwin, const bf16* in0, int nch0, const bf16* in1,
                    bf16* out, const Level& Lout, double* stats) {
    ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.in0 = in0; a.in1 = in1; a.nch0 = nch0;
    a.inS = L.S; a.in_guard = L.guard;
    a.w = Ly.w; a.out = out; a.outS = Lout.S; a.out_guard = Lout.guard;
    a.stats = stats; a.bias = Ly.bias;
    a.Z = L.Z; a.Y = L.Y; a.X = L.X; a.Yp = L.Yp; a.Xp = L.Xp; a.YpXp = L.YpXp; a.Vp = L.Vp;
    a.NP = static_cast<int64_t>(nwin) * L.Vp;
    a.KB = Ly.KB; a.NB = Ly.NB; a.cout = Ly.cout;
    a.oYp = Lout.Yp; a.oXp = Lout.Xp; a.oVp = Lout.Vp;
    const bool conv = Ly.ntaps == 27;
    a.H = conv ? L.Xp + 1 : 0;
    const int nruns = conv ? 3 : 1;
    a.w_bytes = static_cast<uint32_t>(Ly.ntaps) * 2 * Ly.nblk * 16;
    const int64_t ntiles = (a.NP + 127) / 128;
    // largest T (tiles per work item) that fits TMEM (T*nblk <= 256) and two smem stages, while keeping
    // at least ~2 work items per SM when the layer is big enough
    int T = 1;
    for (int t = 8; t >= 1; --t) {
        if (t * Ly.nblk > 256) continue;
        const int rl = ((128 * t + 2 * a.H + 7) / 8) * 8;
        const uint32_t stage = static_cast<uint32_t>(nruns) * 2 * rl * 16 + a.w_bytes;
        if (kConvStages * stage + 256 > kSmemLimit) continue;
        const int64_t items = ((ntiles + t - 1) / t) * Ly.NB;
        if (t > 1 && items < 2LL * ctx->num_sms) continue;
        T = t;
        break;
    }
    a.T = T;
    a.RL = ((128 * T + 2 * a.H + 7) / 8) * 8;
    a.a_bytes = static_cast<uint32_t>(nruns) * 2 * a.RL * 16;
    a.stage_bytes = a.a_bytes + a.w_bytes;
    if (a.RL > 16383 || kConvStages * a.stage_bytes + 256 > kSmemLimit) {
        set_error(ctx, "conv %s: window level %dx%dx%d does not fit the smem tile (RL=%d)", Ly.name.c_str(), L.Z, L.Y, L.X, a.RL);
        return DLV_ERR_UNSUPPORTED;
    }
    const int64_t ngroups = (ntiles + T - 1) / T;
    a.nitems = static_cast<int>(ngroups * Ly.NB);
    const int grid = std::min<int64_t>(ctx->num_sms, a.nitems);
    a.items_per_cta = (a.nitems + grid - 1) / grid;
    const int grid2 = (a.nitems + a.items_per_cta - 1) / a.items_per_cta;
    a.nstages = kConvStages;
    if (!conv) {
        // Deconv stages are 12 KB and one MMA each.  The small launches wait for TMA round trips with two of them in flight
        // (8 stages: 78 -> 72, 117 -> 92, 292 -> 248 us per 128-window batch of cfg2); the one that writes the level-0
        // tensor (4.8 GB) is bound by its 1 KB write streams and slows down with more of them in flight (1.49 -> 1.80 ms),
        // so the depth follows the size of the output per window (profiles/r02_z_launches_cfg2.txt vs r02_h_launches_cfg2.txt).
        const double out_bytes_per_window = static_cast<double>(L.Vp) * 8.0 * Ly.cout * 2.0;      // 42 MB vs 5.9 / 1.8 / 0.6 MB on 96x96x64 windows
        const int want = ctx->deconv_stages > 0 ? ctx->deconv_stages : (out_bytes_per_window > 16e6 ? kConvStages : 8);
        a.nstages = std::max(kConvStages, std::min({kConvStagesMax, want, static_cast<int>((kSmemLimit - 256) / a.stage_bytes)}));
    }
    const uint32_t smem = a.nstages * a.stage_bytes + 256;
    if (ctx->time_convs) cudaEventRecord(ctx->ev0, ctx->stream);
    int rc;
    if (conv) {
        rc = (Ly.nblk == 32) ? launch_conv_t<32, 27, kModeConvStats>(ctx, a, grid2, smem)
                             : launch_conv_t<64, 27, kModeConvStats>(ctx, a, grid2, smem);
    } else {
        rc = ctx->deconv_xstore ? launch_conv_t<256, 1, kModeDeconvScatter, true>(ctx, a, grid2, smem)
                                : launch_conv_t<256, 1, kModeDeconvScatter, false>(ctx, a, grid2, smem);
    }
    if (ctx->time_convs && rc == 0) {
        cudaEventRecord(ctx->ev1, ctx->stream);
        cudaEventSynchronize(ctx->ev1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        ctx->conv_ms += ms;
    }
    return rc;
}


// ------------------------------------------------------------------- input-stationary fused conv launcher (Cout = 32)
static const int kIsStatGroup = 8;      // output planes per InstanceNorm partial record of the fused conv
// partial records per window of a level: (Z / G) groups x at most ceil(PL / 128) columns (T = 1)
static int is_max_parts(const Level& L) {
    const int G = (L.Z % kIsStatGroup == 0) ? kIsStatGroup : L.Z;
    return (L.Z / G) * ((L.YpXp + 127) / 128);      // columns of one tile (T = 1) at most
}

struct IsPlan {
    int T, S, RL, H, NC, NZS, Zs, nstages, nparts, G, nsub;
    uint32_t stage_bytes, w_bytes, smem;
};

static bool plan_conv_is(const Ctx* ctx, const ConvLayer& Ly, const Level& L, int nwin, IsPlan& P, bool xform = false) {
    if (Ly.cout != 32 || Ly.ntaps != 27 || !Ly.w_is) return false;
    // tile-count override: DLV_IS_T for every layer, DLV_IS_TX for the layers that normalise while staging (32 -> 32)
    // and DLV_IS_TF for the uint16 first layer (3 MMAs per tile and plane: the issuing thread's per-plane bookkeeping, not
    // the tensor pipe, bounds it, so more tiles per plane step is what helps)
    const int force_t = (xform && Ly.KB == 2 && ctx->is_tiles_xf) ? ctx->is_tiles_xf : (Ly.cin == 1 && ctx->is_tiles_fold) ? ctx->is_tiles_fold : ctx->is_tiles;
    const int PL = L.YpXp;
    P.H = L.Xp + 1;
    P.w_bytes = static_cast<uint32_t>(Ly.KB) * (Ly.cin == 1 ? 3 : 9) * 3072;
    // 64 -> 32 layers can stage (and multiply) a plane in two halves of 4 chunks - the skip tensor's raw half, which the
    // transform warps normalise, and the up-sampled half, which is ready as it lands: four half-plane stages fit beside
    // the weights where only two whole-plane stages do.  Measured slower (DLV_IS_NSUB=2: 5 704 instead of 5 387 clk
    // per plane on cfg2 - transform and MMA stream then overlap fully and slow each other), so it is off by default.
    P.nsub = (Ly.KB == 4 && ctx->is_nsub == 2) ? 2 : 1;
    const int nchunks = 2 * Ly.KB / P.nsub;      // chunks per stage
    auto fits = [&](int T, int& nst, int& RL, uint32_t& sb) {
        RL = ((128 * T + 2 * P.H + 7) / 8) * 8;
        sb = static_cast<uint32_t>(nchunks) * RL * 16;
        const uint32_t fixed = P.w_bytes + 1024 * 8 + 512;     // weights + statistics combine buffer + barriers
        nst = 0;
        // plain / fused layers: up to 4 stages; the uint16 first layer: one stage per building warp - 4 of the staging
        // warps when the other 4 serve as a second epilogue set (T >= 2), all 8 otherwise
        const int max_st = (Ly.cin != 1) ? 4 : (T >= 2 && kIsXformWarps >= 8 && kIsFoldSets == 2) ? kIsXformWarps - 4 : std::min(kIsMaxStages, kIsXformWarps);
        for (int n = max_st; n >= 2; --n)
            if (fixed + static_cast<uint64_t>(n) * sb <= kSmemLimit) { nst = n; break; }
        return nst >= 2 && RL <= 1024;      // 32 mask words per transform warp
    };
    // cost model per tile and k-block: 9 taps x (56 clk for an N = 96 MMA, 88 where the slot ring wraps: 2 of S steps),
    // times the column quantisation of the plane
    int bestT = 0; double best = 1e30;
    for (int T : {4, 2}) {
        if (force_t && force_t != T) continue;
        int nst, RL; uint32_t sb;
        if (!fits(T, nst, RL, sb)) continue;
        const int S = 16 / T, R = 128 * T;
        const int NC = (PL + R - 1) / R;
        const double mma = 9.0 * ((S - 2) * 56.0 + 2 * 88.0) / S;
        const double cost = mma * (static_cast<double>(NC) * R / PL);
        if (cost < best) { best = cost; bestT = T; }
    }
    // wide windows (a row halo of 2 (X + 2) positions per column): one tile per column, 16 plane slots - the stages
    // of the two-tile column no longer fit next to the weights (X > ~100 for the 64 -> 32 layers)
    if (!bestT && force_t > 1) {          // the forced tile count does not fit this layer: fall back to the model's choice
        for (int T : {4, 2}) {
            int nst, RL; uint32_t sb;
            if (!fits(T, nst, RL, sb)) continue;
            const int S = 16 / T, R = 128 * T;
            const int NC = (PL + R - 1) / R;
            const double cost = 9.0 * ((S - 2) * 56.0 + 2 * 88.0) / S * (static_cast<double>(NC) * R / PL);
            if (cost < best) { best = cost; bestT = T; }
        }
    }
    if (!bestT || force_t == 1) { int nst, RL; uint32_t sb; if (fits(1, nst, RL, sb)) bestT = 1; }
    if (!bestT) return false;
    P.T = bestT; P.S = 16 / bestT;
    fits(P.T, P.nstages, P.RL, P.stage_bytes);
    P.NC = (PL + 128 * P.T - 1) / (128 * P.T);
    // z segmentation: balance the persistent grid (each extra segment re-stages ~2 input planes).  Segments are
    // whole statistics groups of G planes, so the InstanceNorm partial records are the same for every choice.
    P.G = (L.Z % kIsStatGroup == 0) ? kIsStatGroup : L.Z;
    int bestN = 1; double bestc = 1e30;
    for (int n = 1; n <= 6 && n <= L.Z; ++n) {
        const int Zs = (L.Z + n - 1) / n;
        const int nseg = (L.Z + Zs - 1) / Zs;
        if (nseg != n || Zs % P.G != 0) continue;
        const int64_t items = static_cast<int64_t>(nwin) * P.NC * n;
        const int64_t waves = (items + ctx->num_sms - 1) / ctx->num_sms;
        const double c = static_cast<double>(waves) * (Zs + (n > 1 ? 1.5 : 0.0));
        if (c < bestc) { bestc = c; bestN = n; }
    }
    P.NZS = bestN;
    P.Zs = (L.Z + bestN - 1) / bestN;
    P.nparts = (L.Z / P.G) * P.NC;
    P.smem = kSmemLimit;        // always the full opt-in size: one CTA per SM owns all 512 TMEM columns
    return true;
}


template <int T, int S, bool FOLD>
static int launch_conv_is_t(Ctx* ctx, const IsArgs& a, int grid, uint32_t smem) {
    auto k = conv_is_kernel<T, S, FOLD>;
    static bool attr_set = false;
    if (!attr_set) {
        DLV_CUDA_OK(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
        attr_set = true;
    }
    k<<<grid, kIsThreads, smem, ctx->stream>>>(a);
    ctx->launches++;
    DLV_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

// One Cout = 32 conv layer over nwin windows.  `xform`: the leading 4 chunks of in0 hold the RAW output of `prod`
// (statistics in prod_stats) and get InstanceNorm + Mish applied while being staged.
struct RawWindows {            // the first layer's input: uint16 volume + window descriptors (instead of a gathered tensor)
    const uint16_t* slab = nullptr;
    int64_t sy = 0, sx = 0;
    const WindowDesc* wd = nullptr;
};

static int run_conv_is(Ctx* ctx, const ConvLayer& Ly, const Level& L, int nwin, const bf16* in0, int nch0, const bf16* in1,
                       const ConvLayer* prod, const double* prod_stats, bf16* out, double* part, double* stats,
                       const RawWindows* raw = nullptr) {
    IsPlan P;
    if (!plan_conv_is(ctx, Ly, L, nwin, P, prod != nullptr) || P.nparts > is_max_parts(L)) {
        set_error(ctx, "conv %s: level %dx%dx%d does not fit the input-stationary kernel", Ly.name.c_str(), L.Z, L.Y, L.X);
        return DLV_ERR_UNSUPPORTED;
    }
    IsArgs a;
    memset(&a, 0, sizeof(a));
    a.in0 = in0; a.in1 = in1; a.nch0 = nch0; a.nchunks = 2 * Ly.KB;
    a.xform_chunks = prod ? 4 : 0;
    a.in_stats = prod_stats;
    a.in_gamma = prod ? prod->gamma : nullptr;
    a.in_beta = prod ? prod->beta : nullptr;
    a.inS = L.S; a.in_guard = L.guard;
    a.w = Ly.w_is; a.out = out; a.outS = L.S; a.out_guard = L.guard;
    a.part = part; a.nparts = P.nparts;
    a.Z = L.Z; a.Y = L.Y; a.X = L.X; a.Xp = L.Xp; a.PL = L.YpXp; a.Vp = L.Vp;
    a.xp_magic = static_cast<uint32_t>(((1ull << 32) + L.Xp - 1) / L.Xp);
    a.KB = Ly.KB; a.NC = P.NC; a.NZS = P.NZS; a.Zs = P.Zs; a.G = P.G;
    a.nitems = nwin * P.NC * P.NZS;
    a.RL = P.RL; a.H = P.H; a.nstages = P.nstages; a.nsub = P.nsub;
    a.stage_bytes = P.stage_bytes; a.w_bytes = P.w_bytes;
    a.inv_count = 1.0 / (static_cast<double>(L.Z) * L.Y * L.X);
    if (Ly.cin == 1) {
        if (!raw || !raw->slab || !raw->wd) { set_error(ctx, "conv %s: the first layer reads the uint16 windows directly", Ly.name.c_str()); return DLV_ERR_ARG; }
        a.raw_slab = raw->slab; a.raw_sy = raw->sy; a.raw_sx = raw->sx;
        a.raw_wd = reinterpret_cast<const int4*>(raw->wd);
    }
    const int grid = std::min(ctx->num_sms, a.nitems);
    static const bool dbg = getenv("DLV_IS_DEBUG") != nullptr;
    if (const char* e = getenv("DLV_IS_MODE")) a.dbg_mode = atoi(e);
    if (dbg) { cudaMalloc(reinterpret_cast<void**>(&a.dbg), grid * 64); cudaMemset(a.dbg, 0, grid * 64); }
    if (ctx->time_convs) cudaEventRecord(ctx->ev0, ctx->stream);
    int rc;
    if (Ly.cin == 1)
        rc = (P.T == 4) ? launch_conv_is_t<4, 4, true>(ctx, a, grid, P.smem) : (P.T == 2) ? launch_conv_is_t<2, 8, true>(ctx, a, grid, P.smem)
                                                                                         : launch_conv_is_t<1, 16, true>(ctx, a, grid, P.smem);
    else
        rc = (P.T == 4) ? launch_conv_is_t<4, 4, false>(ctx, a, grid, P.smem) : (P.T == 2) ? launch_conv_is_t<2, 8, false>(ctx, a, grid, P.smem)
                                                                                          : launch_conv_is_t<1, 16, false>(ctx, a, grid, P.smem);
    if (ctx->time_convs && rc == 0) {
        cudaEventRecord(ctx->ev1, ctx->stream);
        cudaEventSynchronize(ctx->ev1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        ctx->conv_ms += ms;
    }
    if (dbg && rc == 0) {
        cudaStreamSynchronize(ctx->stream);
        std::vector<long long> h(grid * 8);
        cudaMemcpy(h.data(), a.dbg, grid * 64, cudaMemcpyDeviceToHost);
        cudaFree(a.dbg);
        double t[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int b = 0; b < grid; ++b) for (int k = 0; k < 7; ++k) t[k] += h[b * 8 + k];
        const double st = (t[4] > 0 ? t[4] : 1) / P.nsub;      // input planes (the kernel counts sub-steps)
        fprintf(stderr, "[is] %-22s nwin %d T %d NZS %d nst %d KB %d xf %d: cycles/CTA %.0f | per step: total %.0f mma-thread waits %.0f (slots %.0f, stage %.0f) | xform raw-wait %.0f work %.0f\n",
                Ly.name.c_str(), nwin, P.T, P.NZS, P.nstages, Ly.KB, a.xform_chunks, t[0] / grid, t[0] / st, t[1] / st, t[2] / st, t[3] / st, t[5] / st, t[6] / st);
    }
    if (rc) return rc;
    is_reduce_stats_kernel<<<nwin, 64 * kIsReduceLanes, 0, ctx->stream>>>(part, P.nparts, stats);
    ctx->launches++;
    DLV_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------- engine (activation buffers for one roi/batch)
struct Engine {
    int roi[3] = {0, 0, 0};
    int batch = 0;
    Level L[5];
    // activations (zero halo, never dirtied)
    bf16 *in0 = nullptr, *c0a = nullptr, *x0 = nullptr, *up1 = nullptr, *c1a = nullptr;
    bf16 *p1 = nullptr, *d1a = nullptr, *x1 = nullptr, *up2 = nullptr, *c2a = nullptr, *u2 = nullptr;
    bf16 *p2 = nullptr, *d2a = nullptr, *x2 = nullptr, *up3 = nullptr, *c3a = nullptr, *u3 = nullptr;
    bf16 *p3 = nullptr, *d3a = nullptr, *x3 = nullptr, *up4 = nullptr, *c4a = nullptr, *u4 = nullptr;
    bf16 *p4 = nullptr, *d4a = nullptr, *x4 = nullptr;
    bf16* raw[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // pre-norm conv outputs, one per level
    double* stats = nullptr;     // [18 layers][batch][256][2]
    double* part = nullptr;      // [batch][is_max_parts][64] partial sums of the layer in flight (fused path)
    std::vector<void*> allocs;
};
static const int kStatsPerLayer = 256 * 2;

void engine_free(Ctx* ctx) {
    if (!ctx->eng) return;
    for (void* p : ctx->eng->allocs) cudaFree(p);
    delete ctx->eng;
    ctx->eng = nullptr;
}
int engine_batch_capacity(Ctx* ctx) { return ctx->eng ? ctx->eng->batch : 0; }

static int alloc_act(Ctx* ctx, Engine* e, const Level& L, int nchunk, bf16** out) {
    const size_t bytes = static_cast<size_t>(nchunk) * L.S * 16;
    void* p = nullptr;
    DLV_CUDA_OK(ctx, cudaMalloc(&p, bytes));
    e->allocs.push_back(p);
    DLV_CUDA_OK(ctx, cudaMemsetAsync(p, 0, bytes, ctx->stream));
    *out = static_cast<bf16*>(p);
    return 0;
}

int engine_prepare(Ctx* ctx, const int32_t roi[3], int batch) {
    if (ctx->eng && ctx->eng->roi[0] == roi[0] && ctx->eng->roi[1] == roi[1] && ctx->eng->roi[2] == roi[2] && ctx->eng->batch >= batch)
        return 0;
    for (int i = 0; i < 3; ++i)
        if (roi[i] < 16 || roi[i] % 16) { set_error(ctx, "window dims must be positive multiples of 16 (got %d)", roi[i]); return DLV_ERR_ARG; }
    engine_free(ctx);
    Engine* e = new Engine();
    ctx->eng = e;
    e->roi[0] = roi[0]; e->roi[1] = roi[1]; e->roi[2] = roi[2];
    e->batch = batch;
    for (int l = 0; l < 5; ++l) e->L[l] = make_level(roi[0] >> l, roi[1] >> l, roi[2] >> l, batch);
    int rc = 0;
#define A(ptr, lvl, nch) if ((rc = alloc_act(ctx, e, e->L[lvl], nch, &e->ptr))) return rc;
    A(in0, 0, 2) A(c0a, 0, 4) A(x0, 0, 4) A(up1, 0, 4) A(c1a, 0, 4)
    A(p1, 1, 4) A(d1a, 1, 4) A(x1, 1, 4) A(up2, 1, 4) A(c2a, 1, 4) A(u2, 1, 4)
    A(p2, 2, 4) A(d2a, 2, 8) A(x2, 2, 8) A(up3, 2, 8) A(c3a, 2, 8) A(u3, 2, 8)
    A(p3, 3, 8) A(d3a, 3, 16) A(x3, 3, 16) A(up4, 3, 16) A(c4a, 3, 16) A(u4, 3, 16)
    A(p4, 4, 16) A(d4a, 4, 32) A(x4, 4, 32)
    A(raw[0], 0, 4) A(raw[1], 1, 4) A(raw[2], 2, 8) A(raw[3], 3, 16) A(raw[4], 4, 32)
#undef A
    void* p = nullptr;
    DLV_CUDA_OK(ctx, cudaMalloc(&p, sizeof(double) * 18 * batch * kStatsPerLayer));
    e->allocs.push_back(p);
    e->stats = static_cast<double*>(p);
    DLV_CUDA_OK(ctx, cudaMalloc(&p, sizeof(double) * batch * std::max(is_max_parts(e->L[0]), is_max_parts(e->L[1])) * 64));
    e->allocs.push_back(p);
    e->part = static_cast<double*>(p);
    DLV_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int run_norm(Ctx* ctx, const ConvLayer& Ly, const Level& L, int nwin, const bf16* raw, bf16* out, bf16* pooled,
                    const Level* Lp, const double* stats) {
    const int nchunk = Ly.cout / 8;
    const LevelDev Ld = to_dev(L);
    StageTimer timer(ctx, kStageNorm);
    if (pooled) {
        const int n = (L.Z / 2) * (L.Y / 2) * (L.X / 2);
        dim3 grid((n + 127) / 128, nwin * nchunk);
        norm_mish_kernel<true><<<grid, 128, 0, ctx->stream>>>(raw, Ld, out, pooled, to_dev(*Lp), stats, Ly.gamma, Ly.beta, Ly.cout, nchunk);
    } else {
        const int n = L.Z * L.Y * L.X;
        dim3 grid((n + 255) / 256, nwin * nchunk);
        norm_mish_kernel<false><<<grid, 256, 0, ctx->stream>>>(raw, Ld, out, nullptr, Ld, stats, Ly.gamma, Ly.beta, Ly.cout, nchunk);
    }
    ctx->launches++;
    DLV_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

// conv -> norm(+pool) for one named layer
static int conv_block(Ctx* ctx, Engine* e, int layer_idx, const char* name, int lvl, int nwin, const bf16* in0, int nch0,
                      const bf16* in1, bf16* out, bf16* pooled) {
    const ConvLayer& Ly = ctx->net.conv.at(name);
    double* st = e->stats + static_cast<size_t>(layer_idx) * e->batch * kStatsPerLayer;
    int rc = run_conv(ctx, Ly, e->L[lvl], nwin, in0, nch0, in1, e->raw[lvl], e->L[lvl], st);
    if (rc) return rc;
    if (!out) return 0;   // last layer: the final kernel consumes raw directly
    return run_norm(ctx, Ly, e->L[lvl], nwin, e->raw[lvl], out, pooled, pooled ? &e->L[lvl + 1] : nullptr, st);
}

static int deconv_block(Ctx* ctx, Engine* e, const char* name, int lvl_in, int nwin, const bf16* in, bf16* out) {
    const ConvLayer& Ly = ctx->net.deconv.at(name);
    return run_conv(ctx, Ly, e->L[lvl_in], nwin, in, Ly.cin / 8, nullptr, out, e->L[lvl_in - 1], nullptr);
}


// Fused path: the eight Cout = 32 layers (levels 0 and 1) run on the input-stationary kernel, which normalises its
// input while staging it; only pooled / deconv-input tensors still need an elementwise pass.
static int fused_block(Ctx* ctx, Engine* e, int layer_idx, const char* name, int lvl, int nwin, const bf16* in0, int nch0,
                       const bf16* in1, const char* prod_name, int prod_idx, bf16* out, const RawWindows* raw = nullptr) {
    const ConvLayer& Ly = ctx->net.conv.at(name);
    const ConvLayer* prod = prod_name ? &ctx->net.conv.at(prod_name) : nullptr;
    const double* pst = prod_name ? e->stats + static_cast<size_t>(prod_idx) * e->batch * kStatsPerLayer : nullptr;
    double* st = e->stats + static_cast<size_t>(layer_idx) * e->batch * kStatsPerLayer;
    return run_conv_is(ctx, Ly, e->L[lvl], nwin, in0, nch0, in1, prod, pst, out, e->part, st, raw);
}
static int norm_block(Ctx* ctx, Engine* e, int layer_idx, const char* name, int lvl, int nwin, const bf16* raw, bf16* out, bf16* pooled) {
    const ConvLayer& Ly = ctx->net.conv.at(name);
    const double* st = e->stats + static_cast<size_t>(layer_idx) * e->batch * kStatsPerLayer;
    return run_norm(ctx, Ly, e->L[lvl], nwin, raw, out, pooled, pooled ? &e->L[lvl + 1] : nullptr, st);
}

// Cout = 32 layers on the fused input-stationary kernel?  (window levels 0 and 1 must fit its shared-memory stages)
static bool fused_path(Ctx* ctx, int nwin) {
    Engine* e = ctx->eng;
    IsPlan probe;
    return ctx->use_fused && plan_conv_is(ctx, ctx->net.conv.at("upcat_1.convs.conv_0"), e->L[0], nwin, probe) &&
           plan_conv_is(ctx, ctx->net.conv.at("upcat_2.convs.conv_0"), e->L[1], nwin, probe) &&
           plan_conv_is(ctx, ctx->net.conv.at("conv_0.conv_0"), e->L[0], nwin, probe);
}

// 18 convs, 4 deconvs, norms, final blend.  Fused path: the first layer reads the uint16 windows itself (no gather);
// per-tap path: from the gathered tensor in0.
static int forward_windows(Ctx* ctx, int nwin, bool fused, const RawWindows& rawin, const WindowDesc* wd_dev, int32_t* acc, int64_t slabY,
                           int64_t slabX, const BlendDev& bw, float* logits_out) {
    Engine* e = ctx->eng;
    int rc;
    DLV_CUDA_OK(ctx, cudaMemsetAsync(e->stats, 0, sizeof(double) * 18 * e->batch * kStatsPerLayer, ctx->stream));
#define CB(i, name, lvl, a, na, b, out, pool) if ((rc = conv_block(ctx, e, i, name, lvl, nwin, a, na, b, out, pool))) return rc;
    const bf16* last_raw = e->raw[0];
    if (fused) {
#define FB(i, name, lvl, a, na, b, prod, pi, out) if ((rc = fused_block(ctx, e, i, name, lvl, nwin, a, na, b, prod, pi, out))) return rc;
        if ((rc = fused_block(ctx, e, 0, "conv_0.conv_0", 0, nwin, nullptr, 2, nullptr, nullptr, 0, e->c0a, &rawin))) return rc;
        FB(1, "conv_0.conv_1", 0, e->c0a, 4, nullptr, "conv_0.conv_0", 0, e->x0)
        if ((rc = norm_block(ctx, e, 1, "conv_0.conv_1", 0, nwin, e->x0, nullptr, e->p1))) return rc;
        FB(2, "down_1.convs.conv_0", 1, e->p1, 4, nullptr, nullptr, 0, e->d1a)
        FB(3, "down_1.convs.conv_1", 1, e->d1a, 4, nullptr, "down_1.convs.conv_0", 2, e->x1)
        if ((rc = norm_block(ctx, e, 3, "down_1.convs.conv_1", 1, nwin, e->x1, nullptr, e->p2))) return rc;
    } else {
    CB(0, "conv_0.conv_0", 0, e->in0, 2, nullptr, e->c0a, nullptr)
    CB(1, "conv_0.conv_1", 0, e->c0a, 4, nullptr, e->x0, e->p1)
    CB(2, "down_1.convs.conv_0", 1, e->p1, 4, nullptr, e->d1a, nullptr)
    CB(3, "down_1.convs.conv_1", 1, e->d1a, 4, nullptr, e->x1, e->p2)
    }
    CB(4, "down_2.convs.conv_0", 2, e->p2, 4, nullptr, e->d2a, nullptr)
    CB(5, "down_2.convs.conv_1", 2, e->d2a, 8, nullptr, e->x2, e->p3)
    CB(6, "down_3.convs.conv_0", 3, e->p3, 8, nullptr, e->d3a, nullptr)
    CB(7, "down_3.convs.conv_1", 3, e->d3a, 16, nullptr, e->x3, e->p4)
    CB(8, "down_4.convs.conv_0", 4, e->p4, 16, nullptr, e->d4a, nullptr)
    CB(9, "down_4.convs.conv_1", 4, e->d4a, 32, nullptr, e->x4, nullptr)
    if ((rc = deconv_block(ctx, e, "upcat_4", 4, nwin, e->x4, e->up4))) return rc;
    CB(10, "upcat_4.convs.conv_0", 3, e->x3, 16, e->up4, e->c4a, nullptr)
    CB(11, "upcat_4.convs.conv_1", 3, e->c4a, 16, nullptr, e->u4, nullptr)
    if ((rc = deconv_block(ctx, e, "upcat_3", 3, nwin, e->u4, e->up3))) return rc;
    CB(12, "upcat_3.convs.conv_0", 2, e->x2, 8, e->up3, e->c3a, nullptr)
    CB(13, "upcat_3.convs.conv_1", 2, e->c3a, 8, nullptr, e->u3, nullptr)
    if ((rc = deconv_block(ctx, e, "upcat_2", 2, nwin, e->u3, e->up2))) return rc;
    if (fused) {
        FB(14, "upcat_2.convs.conv_0", 1, e->x1, 4, e->up2, "down_1.convs.conv_1", 3, e->c2a)
        FB(15, "upcat_2.convs.conv_1", 1, e->c2a, 4, nullptr, "upcat_2.convs.conv_0", 14, e->raw[1])
        if ((rc = norm_block(ctx, e, 15, "upcat_2.convs.conv_1", 1, nwin, e->raw[1], e->u2, nullptr))) return rc;
        if ((rc = deconv_block(ctx, e, "upcat_1", 1, nwin, e->u2, e->up1))) return rc;
        FB(16, "upcat_1.convs.conv_0", 0, e->x0, 4, e->up1, "conv_0.conv_1", 1, e->c1a)
        FB(17, "upcat_1.convs.conv_1", 0, e->c1a, 4, nullptr, "upcat_1.convs.conv_0", 16, e->raw[0])
#undef FB
    } else {
    CB(14, "upcat_2.convs.conv_0", 1, e->x1, 4, e->up2, e->c2a, nullptr)
    CB(15, "upcat_2.convs.conv_1", 1, e->c2a, 4, nullptr, e->u2, nullptr)
    if ((rc = deconv_block(ctx, e, "upcat_1", 1, nwin, e->u2, e->up1))) return rc;
    CB(16, "upcat_1.convs.conv_0", 0, e->x0, 4, e->up1, e->c1a, nullptr)
    CB(17, "upcat_1.convs.conv_1", 0, e->c1a, 4, nullptr, nullptr, nullptr)
    }
#undef CB
    const ConvLayer& last = ctx->net.conv.at("upcat_1.convs.conv_1");
    const Level& L0 = e->L[0];
    const int n = L0.Z * L0.Y * L0.X;
    dim3 grid((n + 255) / 256, nwin);
    StageTimer timer(ctx, kStageBlend);
    final_blend_kernel<<<grid, 256, 0, ctx->stream>>>(last_raw, to_dev(L0), e->stats + static_cast<size_t>(17) * e->batch * kStatsPerLayer,
                                                      last.gamma, last.beta, ctx->net.final_w, ctx->net.final_b, wd_dev, acc,
                                                      slabY, slabX, bw, logits_out);
    ctx->launches++;
    DLV_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

int engine_run_batch(Ctx* ctx, const uint16_t* slab, int64_t slabY, int64_t slabX, const WindowDesc* wd_dev, int nwin,
                     int32_t* acc, const BlendDev& bw, float* logits_out) {
    Engine* e = ctx->eng;
    if (!e || !ctx->net.loaded) { set_error(ctx, "engine_run_batch: weights/engine not ready"); return DLV_ERR_STATE; }
    if (nwin < 1 || nwin > e->batch) { set_error(ctx, "engine_run_batch: nwin %d outside [1,%d]", nwin, e->batch); return DLV_ERR_ARG; }
    const Level& L0 = e->L[0];
    const bool fused = fused_path(ctx, nwin);
    RawWindows rawin;
    rawin.slab = slab; rawin.sy = slabY; rawin.sx = slabX; rawin.wd = wd_dev;
    if (!fused) {
        const int n = L0.Z * L0.Y * L0.X;
        dim3 grid((n + 255) / 256, nwin);
        StageTimer timer(ctx, kStageGather);
        gather_windows_kernel<<<grid, 256, 0, ctx->stream>>>(slab, slabY, slabX, wd_dev, to_dev(L0), e->in0);
        ctx->launches++;
        DLV_CUDA_OK(ctx, cudaGetLastError());
    }
    return forward_windows(ctx, nwin, fused, rawin, wd_dev, acc, slabY, slabX, bw, logits_out);
}

int windows_active(Ctx* ctx, const uint16_t* slab, int64_t slabY, int64_t slabX, const int32_t* origins_dev, int n,
                   const int32_t roi[3], int32_t* active_dev) {
    if (n <= 0) return 0;
    window_active_kernel<<<n, 256, 0, ctx->stream>>>(slab, slabY, slabX, origins_dev, roi[0], roi[1], roi[2], active_dev);
    ctx->launches++;
    DLV_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------- operator-level entry points (tests)
static int temp_level_tensor(Ctx* ctx, const Level& L, int nchunk, bf16** out) {
    const size_t bytes = static_cast<size_t>(nchunk) * L.S * 16;
    DLV_CUDA_OK(ctx, cudaMalloc(reinterpret_cast<void**>(out), bytes));
    DLV_CUDA_OK(ctx, cudaMemsetAsync(*out, 0, bytes, ctx->stream));
    return 0;
}

int op_conv3d(Ctx* ctx, const char* name, const float* x, int n, int D, int H, int W, float* y, double* stats) {
    auto it = ctx->net.conv.find(name);
    if (it == ctx->net.conv.end()) { set_error(ctx, "dlv_op_conv3d: unknown layer '%s'", name); return DLV_ERR_ARG; }
    const ConvLayer& Ly = it->second;
    if (Ly.cin == 1) { set_error(ctx, "dlv_op_conv3d: the uint16 first layer is covered by dlv_unet_forward"); return DLV_ERR_UNSUPPORTED; }
    const Level L = make_level(D, H, W, n);
    bf16 *in = nullptr, *raw = nullptr;
    int rc;
    if ((rc = temp_level_tensor(ctx, L, Ly.cin_pad / 8, &in))) return rc;
    if ((rc = temp_level_tensor(ctx, L, Ly.cout / 8, &raw))) { cudaFree(in); return rc; }
    const int nv = D * H * W;
    pack_act_kernel<<<dim3((nv + 255) / 256, n * (Ly.cin_pad / 8)), 256, 0, ctx->stream>>>(x, Ly.cin, to_dev(L), in, Ly.cin_pad / 8);
    ctx->launches++;
    cudaMemsetAsync(stats, 0, sizeof(double) * n * Ly.cout * 2, ctx->stream);
    IsPlan probe;
    double* part = nullptr;
    if (ctx->use_fused && plan_conv_is(ctx, Ly, L, n, probe)) {
        // Cout = 32 layers: the input-stationary kernel (plain, already-normalised input)
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&part), sizeof(double) * n * is_max_parts(L) * 64);
        if (e != cudaSuccess) { set_error(ctx, "dlv_op_conv3d: %s", cudaGetErrorString(e)); cudaFree(in); cudaFree(raw); return DLV_ERR_CUDA; }
        rc = run_conv_is(ctx, Ly, L, n, in, Ly.cin_pad / 8, nullptr, nullptr, nullptr, raw, part, stats);
    } else {
        rc = run_conv(ctx, Ly, L, n, in, Ly.cin_pad / 8, nullptr, raw, L, stats);
    }
    if (rc == 0) {
        unpack_act_kernel<<<dim3((nv + 255) / 256, n * (Ly.cout / 8)), 256, 0, ctx->stream>>>(raw, Ly.cout, to_dev(L), y, Ly.cout / 8);
        ctx->launches++;
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { set_error(ctx, "dlv_op_conv3d: %s", cudaGetErrorString(e)); rc = DLV_ERR_CUDA; }
    }
    cudaFree(in); cudaFree(raw); cudaFree(part);
    return rc;
}

int op_deconv(Ctx* ctx, const char* name, const float* x, int n, int D, int H, int W, float* y) {
    auto it = ctx->net.deconv.find(name);
    if (it == ctx->net.deconv.end()) { set_error(ctx, "dlv_op_deconv: unknown layer '%s'", name); return DLV_ERR_ARG; }
    const ConvLayer& Ly = it->second;
    const Level L = make_level(D, H, W, n), Lo = make_level(2 * D, 2 * H, 2 * W, n);
    bf16 *in = nullptr, *out = nullptr;
    int rc;
    if ((rc = temp_level_tensor(ctx, L, Ly.cin / 8, &in))) return rc;
    if ((rc = temp_level_tensor(ctx, Lo, Ly.cout / 8, &out))) { cudaFree(in); return rc; }
    const int nv = D * H * W;
    pack_act_kernel<<<dim3((nv + 255) / 256, n * (Ly.cin / 8)), 256, 0, ctx->stream>>>(x, Ly.cin, to_dev(L), in, Ly.cin / 8);
    ctx->launches++;
    rc = run_conv(ctx, Ly, L, n, in, Ly.cin / 8, nullptr, out, Lo, nullptr);
    if (rc == 0) {
        unpack_act_kernel<<<dim3((nv * 8 + 255) / 256, n * (Ly.cout / 8)), 256, 0, ctx->stream>>>(out, Ly.cout, to_dev(Lo), y, Ly.cout / 8);
        ctx->launches++;
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { set_error(ctx, "dlv_op_deconv: %s", cudaGetErrorString(e)); rc = DLV_ERR_CUDA; }
    }
    cudaFree(in); cudaFree(out);
    return rc;
}

}  // namespace dlv
