// dlv_post.cu - create_nifti_seg on the GPU (inference/inference.py:31-95 of the reference):
//   binaries = (sigmoid(avg_logits) >= threshold) AND binary_erosion(input > 0, iterations=30, border_value=1)
// evaluated per numpy Arrayterator block (z-slabs of `block_planes` planes, inference.py:53).
//
// Iterated erosion with the 6-neighbour cross and border_value=1 equals "L1 distance to the nearest zero voxel
// inside the block > iterations".  The L1 distance transform is separable: an exact 1-D pass along x from a
// per-row zero bitmask, then min-plus sweeps (d = min(d, d_prev + 1)) along y and along z, all capped at
// iterations+1 in uint8.  HBM-bound byte work: every pass is coalesced along x.
#include <algorithm>

#include "dlv_common.cuh"
#include "dlv_internal.h"

namespace dlv {

struct PostGeom {
    int64_t SY, SX;       // strides (in voxels) of the padded slab arrays (volume, avg): plane = SY*SX, row = SX
    int64_t Y, X;         // real (unpadded) in-plane extent = extent of dist / binaries rows
    int64_t nplanes;      // local planes covered by dist
    int cap;              // iterations + 1
};

// ---- pass X: one warp per (plane,row); zero bitmask of the row in shared memory, exact nearest-zero search
__global__ void erode_x_kernel(const uint16_t* __restrict__ vol, PostGeom g, int nwords, uint8_t* __restrict__ dist) {
    extern __shared__ uint32_t sbits[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + wib;
    if (row >= g.nplanes * g.Y) return;
    uint32_t* bits = sbits + wib * nwords;
    const int64_t z = row / g.Y, y = row - z * g.Y;
    const uint16_t* src = vol + (z * g.SY + y) * g.SX;
    for (int w = 0; w < nwords; ++w) {
        const int64_t x = static_cast<int64_t>(w) * 32 + lane;
        const bool zero = (x < g.X) && (src[x] == 0);     // outside the array counts as 1 (border_value=1)
        const uint32_t m = __ballot_sync(0xffffffffu, zero);
        if (lane == 0) bits[w] = m;
    }
    __syncwarp();
    uint8_t* dst = dist + row * g.X;
    for (int64_t x = lane; x < g.X; x += 32) {
        const int wi = static_cast<int>(x >> 5), b = static_cast<int>(x & 31);
        int d = g.cap;
        // nearest zero at or left of x
        uint32_t m = bits[wi] & (0xFFFFFFFFu >> (31 - b));
        for (int k = 0;; ++k) {
            if (m) { d = min(d, b + 32 * k - (31 - __clz(m))); break; }
            if (b + 1 + 32 * k >= d || wi - k - 1 < 0) break;
            m = bits[wi - k - 1];
        }
        // nearest zero right of x
        m = bits[wi] & (0xFFFFFFFEu << b);    // bits > b  (b == 31 -> 0)
        if (b == 31) m = 0;
        for (int k = 0;; ++k) {
            if (m) { d = min(d, (__ffs(m) - 1) + 32 * k - b); break; }
            if (32 * (k + 1) - b >= d || wi + k + 1 >= nwords) break;
            m = bits[wi + k + 1];
        }
        dst[x] = static_cast<uint8_t>(d);
    }
}

// ---- pass Y: forward then backward min-plus sweep down a column, in place.
// The sweep is a serial chain, but its LOADS do not depend on it: kSweepBatch rows are fetched before the chain
// consumes them (a load issued after the previous row's store to the same array is not hoisted by the compiler, which
// left every row paying a full memory round trip).  erode_y4 handles 4 adjacent columns per thread with byte-wise
// SIMD min / saturating add on 32-bit words (X % 4 == 0); erode_y is the scalar form for other widths.
constexpr int kSweepBatch = 8;
__global__ void erode_y4_kernel(uint8_t* __restrict__ dist, PostGeom g) {
    const int64_t x4 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t z = blockIdx.y;
    const int64_t pitch = g.X >> 2;
    if (x4 >= pitch) return;
    uint32_t* col = reinterpret_cast<uint32_t*>(dist + z * g.Y * g.X) + x4;
    const uint32_t cap4 = static_cast<uint32_t>(g.cap) * 0x01010101u, one4 = 0x01010101u;
    uint32_t d = cap4, v[kSweepBatch];
    for (int64_t y0 = 0; y0 < g.Y; y0 += kSweepBatch) {
#pragma unroll
        for (int k = 0; k < kSweepBatch; ++k) v[k] = (y0 + k < g.Y) ? col[(y0 + k) * pitch] : 0u;
#pragma unroll
        for (int k = 0; k < kSweepBatch; ++k) {
            d = __vminu4(v[k], __vminu4(__vaddus4(d, one4), cap4));
            if (y0 + k < g.Y) col[(y0 + k) * pitch] = d;
        }
    }
    d = cap4;
    for (int64_t y0 = g.Y - 1; y0 >= 0; y0 -= kSweepBatch) {
#pragma unroll
        for (int k = 0; k < kSweepBatch; ++k) v[k] = (y0 - k >= 0) ? col[(y0 - k) * pitch] : 0u;
#pragma unroll
        for (int k = 0; k < kSweepBatch; ++k) {
            d = __vminu4(v[k], __vminu4(__vaddus4(d, one4), cap4));
            if (y0 - k >= 0) col[(y0 - k) * pitch] = d;
        }
    }
}
__global__ void erode_y_kernel(uint8_t* __restrict__ dist, PostGeom g) {
    const int64_t x = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t z = blockIdx.y;
    if (x >= g.X) return;
    uint8_t* col = dist + z * g.Y * g.X + x;
    int d = g.cap, v[kSweepBatch];
    for (int64_t y0 = 0; y0 < g.Y; y0 += kSweepBatch) {
#pragma unroll
        for (int k = 0; k < kSweepBatch; ++k) v[k] = (y0 + k < g.Y) ? col[(y0 + k) * g.X] : 0;
#pragma unroll
        for (int k = 0; k < kSweepBatch; ++k) {
            d = min(v[k], min(d + 1, g.cap));
            if (y0 + k < g.Y) col[(y0 + k) * g.X] = static_cast<uint8_t>(d);
        }
    }
    d = g.cap;
    for (int64_t y0 = g.Y - 1; y0 >= 0; y0 -= kSweepBatch) {
#pragma unroll
        for (int k = 0; k < kSweepBatch; ++k) v[k] = (y0 - k >= 0) ? col[(y0 - k) * g.X] : 0;
#pragma unroll
        for (int k = 0; k < kSweepBatch; ++k) {
            d = min(v[k], min(d + 1, g.cap));
            if (y0 - k >= 0) col[(y0 - k) * g.X] = static_cast<uint8_t>(d);
        }
    }
}

// ---- pass Z (per Arrayterator block) fused with sigmoid/threshold and the final AND.
// dist covers local planes [0, nplanes) = global planes [gz0, gz0+nplanes); blocks are
// [k*bp, (k+1)*bp) in GLOBAL plane numbers; blockIdx.z enumerates the blocks intersecting the slab.
// Output planes [oz0, oz1) (global) go to binaries (row pitch X, plane pitch Y*X, first plane = oz0).
__global__ void erode_z_final_kernel(uint8_t* __restrict__ dist, PostGeom g, const float* __restrict__ avg,
                                     int64_t gz0, int64_t Zreal, int64_t bp, int64_t first_block, int64_t oz0, int64_t oz1,
                                     float thr, int iters, uint8_t* __restrict__ bin, float* __restrict__ sig) {
    const int64_t x = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t y = blockIdx.y;
    if (x >= g.X) return;
    const int64_t blk = first_block + blockIdx.z;
    int64_t b0 = blk * bp, b1 = min(Zreal, b0 + bp);            // global block range
    int64_t l0 = max(b0, gz0) - gz0, l1 = min(b1, gz0 + g.nplanes) - gz0;   // local planes of the block present here
    if (l0 >= l1) return;
    const int64_t pstride = g.Y * g.X;
    uint8_t* col = dist + y * g.X + x;
    int d = g.cap, v[kSweepBatch];
    for (int64_t z0 = l0; z0 < l1; z0 += kSweepBatch) {
#pragma unroll
        for (int k = 0; k < kSweepBatch; ++k) v[k] = (z0 + k < l1) ? col[(z0 + k) * pstride] : 0;
#pragma unroll
        for (int k = 0; k < kSweepBatch; ++k) {
            d = min(v[k], min(d + 1, g.cap));
            if (z0 + k < l1) col[(z0 + k) * pstride] = static_cast<uint8_t>(d);
        }
    }
    d = g.cap;
    for (int64_t z0 = l1 - 1; z0 >= l0; z0 -= kSweepBatch) {
        float a[kSweepBatch];
#pragma unroll
        for (int k = 0; k < kSweepBatch; ++k) {
            const int64_t z = z0 - k, gz = gz0 + z;
            v[k] = (z >= l0) ? col[z * pstride] : 0;
            a[k] = (z >= l0 && gz >= oz0 && gz < oz1) ? avg[(z * g.SY + y) * g.SX + x] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kSweepBatch; ++k) {
            const int64_t z = z0 - k, gz = gz0 + z;
            if (z < l0) break;
            d = min(v[k], min(d + 1, g.cap));
            if (gz >= oz0 && gz < oz1) {
                // the reference keeps averaged logits in fp16 (inference.py:242,295) and applies an fp32 sigmoid (:65-68)
                const float ah = __half2float(__float2half_rn(a[k]));
                const float s = 1.f / (1.f + expf(-ah));
                const int64_t o = ((gz - oz0) * g.Y + y) * g.X + x;
                if (sig) sig[o] = s;
                bin[o] = static_cast<uint8_t>((s >= thr) && (d > iters));
            }
        }
    }
}

// Generalised form used by dlv_op_finalise (whole volume: gz0 = 0, output = all planes) and by the slab driver.
int post_finalise_slab(Ctx* ctx, const float* avg, const uint16_t* vol, int64_t SY, int64_t SX, int64_t nplanes, int64_t gz0,
                       const int64_t sr[3], float thr, int iters, int64_t block_planes, int64_t oz0, int64_t oz1,
                       uint8_t* bin, float* sig) {
    if (iters < 0 || iters > 254) { set_error(ctx, "erosion iterations %d outside [0,254]", iters); return DLV_ERR_ARG; }
    PostGeom g;
    g.SY = SY; g.SX = SX; g.Y = sr[1]; g.X = sr[2]; g.nplanes = nplanes; g.cap = iters + 1;
    if (nplanes <= 0 || g.Y <= 0 || g.X <= 0) return 0;
    uint8_t* dist = nullptr;
    DLV_CUDA_OK(ctx, scratch_get(ctx, kScratchDist, static_cast<size_t>(nplanes) * g.Y * g.X, reinterpret_cast<void**>(&dist)));
    const int nwords = static_cast<int>((g.X + 31) / 32);
    const int wpb = 8;
    const size_t smem = static_cast<size_t>(wpb) * nwords * 4;
    const int64_t rows = nplanes * g.Y;
    erode_x_kernel<<<static_cast<unsigned>((rows + wpb - 1) / wpb), wpb * 32, smem, ctx->stream>>>(vol, g, nwords, dist);
    if ((g.X & 3) == 0)     // dist comes from the pool (256 B aligned) and rows are X bytes: every row is word aligned
        erode_y4_kernel<<<dim3(static_cast<unsigned>((g.X / 4 + 63) / 64), static_cast<unsigned>(nplanes)), 64, 0, ctx->stream>>>(dist, g);
    else
        erode_y_kernel<<<dim3(static_cast<unsigned>((g.X + 127) / 128), static_cast<unsigned>(nplanes)), 128, 0, ctx->stream>>>(dist, g);
    const int64_t Zreal = sr[0];
    const int64_t bp = block_planes > 0 ? block_planes : Zreal;
    const int64_t first_block = gz0 / bp;
    const int64_t last_block = (std::min<int64_t>(Zreal, gz0 + nplanes) - 1) / bp;
    dim3 grid(static_cast<unsigned>((g.X + 127) / 128), static_cast<unsigned>(g.Y), static_cast<unsigned>(last_block - first_block + 1));
    erode_z_final_kernel<<<grid, 128, 0, ctx->stream>>>(dist, g, avg, gz0, Zreal, bp, first_block, oz0, oz1, thr, iters, bin, sig);
    ctx->launches += 3;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { set_error(ctx, "finalise kernels failed: %s", cudaGetErrorString(e)); return DLV_ERR_CUDA; }
    return 0;
}

int post_finalise(Ctx* ctx, const float* avg, const uint16_t* vol, const int64_t sp[3], const int64_t sr[3], float thr,
                  int iters, int64_t block_planes, uint8_t* bin, float* sig) {
    return post_finalise_slab(ctx, avg, vol, sp[1], sp[2], sr[0], 0, sr, thr, iters, block_planes, 0, sr[0], bin, sig);
}

}  // namespace dlv
