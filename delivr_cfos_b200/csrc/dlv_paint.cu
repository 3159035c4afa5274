// dlv_paint.cu - blob painter.
//
// Replaces the per-cell Python loops that colour every blob through its bounding box:
//   blob_highlighter.py:107-124 (R/G/B uint8 volumes), :143-151 (region-id uint16 volume),
//   blob_depthmap.py:198-207 (depth uint16 volume):
//       for k in order:  out_c[box_k] = mask[box_k] * value[k][c]
// Every assignment rewrites the whole box, so the result is
//       out_c[v] = mask[v] * value[K(v)][c],  K(v) = the LAST box in paint order that contains v   (0 if none)
// - including the reference's "a long blob re-colours its neighbours" behaviour where boxes overlap.
//
// Two HBM-bound byte passes per z-chunk:
//   paint_owner_*    box k -> atomicMax(owner[v], k + 1) over its foreground voxels (one warp per small box, the
//                    whole grid per large box), which makes the result independent of scheduling;
//   paint_resolve    owner -> value lookup -> coalesced, vectorised store of the nch output volumes; owner entries
//                    are reset on the way, so the scratch is cleared once, not once per chunk.
//                    (Measured alternative, round 2: outputs cleared by memsets + a second walk over the boxes that
//                    paints the voxels whose owner entry names the box - no pass over the background at all - was
//                    SLOWER on cfg3, 23.7 vs 17.5 ms per call: the box walk pays an integer division and scattered
//                    byte accesses per box voxel twice.  profiles/r02_y_bench_cfg3*.json)
// Algorithmic bytes per voxel: 1 (mask) + nch * elem_bytes (outputs); the owner scratch (4 B, touched only around
// foreground) is overhead counted against the achieved fraction.
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <thread>
#include <vector>

#include "dlv_internal.h"

namespace dlv {

struct PaintGeom {
    int64_t Y, X;      // in-plane extent
    int64_t z0, z1;    // global planes of the chunk [z0, z1)
};

constexpr int64_t kPaintBigBox = 1 << 15;   // voxels (after clipping to the chunk) above which a box is painted by the whole grid

__device__ __forceinline__ bool clip_box(const int32_t* __restrict__ b, const PaintGeom& g, int64_t Z, int64_t& za, int64_t& zb,
                                         int64_t& ya, int64_t& yb, int64_t& xa, int64_t& xb) {
    // numpy slice semantics for non-negative bounds: clipped at the array end, empty when start >= stop
    za = max(static_cast<int64_t>(b[0]), g.z0); zb = min(min(static_cast<int64_t>(b[1]), Z), g.z1);
    ya = b[2]; yb = min(static_cast<int64_t>(b[3]), g.Y);
    xa = b[4]; xb = min(static_cast<int64_t>(b[5]), g.X);
    return za < zb && ya < yb && xa < xb;
}

__global__ void paint_owner_small_kernel(const uint8_t* __restrict__ mask, const int32_t* __restrict__ boxes, int64_t n, int64_t Z,
                                         PaintGeom g, uint32_t* __restrict__ owner, uint32_t* __restrict__ big, uint32_t* __restrict__ nbig) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t k = warp0; k < n; k += nwarps) {
        int64_t za, zb, ya, yb, xa, xb;
        if (!clip_box(boxes + k * 6, g, Z, za, zb, ya, yb, xa, xb)) continue;
        const int64_t bx = xb - xa, by = yb - ya, vol = (zb - za) * by * bx;
        if (vol > kPaintBigBox) {
            if (lane == 0) big[atomicAdd(nbig, 1u)] = static_cast<uint32_t>(k);
            continue;
        }
        const int ibx = static_cast<int>(bx), iby = static_cast<int>(by), ivol = static_cast<int>(vol);
        for (int i = lane; i < ivol; i += 32) {
            const int r = i / ibx, dx = i - r * ibx;
            const int dz = r / iby, dy = r - dz * iby;
            const int64_t v = ((za + dz - g.z0) * g.Y + (ya + dy)) * g.X + (xa + dx);     // chunk-local voxel index
            if (mask[v]) atomicMax(owner + v, static_cast<uint32_t>(k + 1));
        }
    }
}

__global__ void paint_owner_big_kernel(const uint8_t* __restrict__ mask, const int32_t* __restrict__ boxes, int64_t Z, PaintGeom g,
                                       uint32_t* __restrict__ owner, const uint32_t* __restrict__ big, const uint32_t* __restrict__ nbig) {
    const uint32_t nb = *nbig;
    const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t nt = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (uint32_t j = 0; j < nb; ++j) {
        const int64_t k = big[j];
        int64_t za, zb, ya, yb, xa, xb;
        clip_box(boxes + k * 6, g, Z, za, zb, ya, yb, xa, xb);
        const int64_t bx = xb - xa, by = yb - ya, vol = (zb - za) * by * bx;
        for (int64_t i = t0; i < vol; i += nt) {
            const int64_t r = i / bx, dx = i - r * bx;
            const int64_t dz = r / by, dy = r - dz * by;
            const int64_t v = ((za + dz - g.z0) * g.Y + (ya + dy)) * g.X + (xa + dx);
            if (mask[v]) atomicMax(owner + v, static_cast<uint32_t>(k + 1));
        }
    }
}

struct PaintOut {
    void* p[3];
};

// Resolve.  Each lane takes 16 consecutive voxels: one 16 B mask load and 16 B (uint8) / 2 x 16 B (uint16) streaming
// ZERO stores per channel - 98 % of a blob volume is background and this is all that happens there.  Lanes whose 16
// voxels contain foreground are then served two at a time by the two half-warps, one voxel per
// lane: owner lookup + reset, value lookup, element store over the zeros (ordered by __syncwarp).  Doing the 16 x nch
// conditional look-ups inside the lane that owns the group made nearly every warp walk ~600 predicated instructions
// (1.5-1.8 TB/s, instruction-bound).  `n16` whole groups when every pointer of the chunk is 16 B aligned (vec),
// then a scalar tail.
template <typename T>
__global__ void paint_resolve_kernel(const uint8_t* __restrict__ mask, uint32_t* __restrict__ owner, int64_t nvox,
                                     const uint16_t* __restrict__ values, int nch, PaintOut out, int vec) {
    const int64_t n16 = vec ? (nvox >> 4) : 0;
    const int lane = threadIdx.x & 31;
    const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t nt = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t qb = t0 - lane; qb < n16; qb += nt) {       // warp-uniform trip count
        const int64_t q = qb + lane;
        uint4 m4 = make_uint4(0u, 0u, 0u, 0u);
        if (q < n16) {
            m4 = __ldcs(reinterpret_cast<const uint4*>(mask) + q);
            const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
            for (int c = 0; c < nch; ++c) {
                uint4* o = reinterpret_cast<uint4*>(out.p[c]) + q * sizeof(T);
                __stcs(o, z4);
                if (sizeof(T) == 2) __stcs(o + 1, z4);
            }
        }
        unsigned pending = __ballot_sync(0xffffffffu, (m4.x | m4.y | m4.z | m4.w) != 0u);
        if (!pending) continue;
        __syncwarp();                                        // the zero stores above precede the element stores below
        while (pending) {
            // Two groups per round, one per half-warp, and kRounds rounds per pass: the owner loads of a pass are all
            // issued before the first of them is used.  A round is a dependent chain (owner load -> value load ->
            // store) of DRAM / L2 round trips, so the number of chains walked one after the other - not the work in
            // them - is what the foreground costs (one group per round: 8.5 ms on cfg3, 2.9 TB/s).
            constexpr int kRounds = 4;
            uint32_t o[kRounds], mbv[kRounds];
            int64_t vv[kRounds];
#pragma unroll
            for (int r = 0; r < kRounds; ++r) {
                o[r] = 0u; mbv[r] = 0u; vv[r] = 0;
                if (!pending) continue;                              // warp-uniform
                const int s0 = __ffs(pending) - 1;
                pending &= pending - 1;
                const int s1 = pending ? __ffs(pending) - 1 : -1;
                if (s1 >= 0) pending &= pending - 1;
                const int src = (lane < 16) ? s0 : s1;
                const int from = src < 0 ? lane : src;
                const uint32_t mx = __shfl_sync(0xffffffffu, m4.x, from), my = __shfl_sync(0xffffffffu, m4.y, from);
                const uint32_t mz = __shfl_sync(0xffffffffu, m4.z, from), mw = __shfl_sync(0xffffffffu, m4.w, from);
                if (src >= 0) {
                    const int l16 = lane & 15;
                    const uint32_t word = (l16 & 8) ? ((l16 & 4) ? mw : mz) : ((l16 & 4) ? my : mx);
                    mbv[r] = (word >> (8 * (l16 & 3))) & 0xFFu;
                    vv[r] = ((qb + src) << 4) + l16;
                    if (mbv[r]) o[r] = owner[vv[r]];
                }
            }
#pragma unroll
            for (int r = 0; r < kRounds; ++r) {
                if (o[r]) {
                    owner[vv[r]] = 0u;
                    for (int c = 0; c < nch; ++c)
                        static_cast<T*>(out.p[c])[vv[r]] = static_cast<T>(mbv[r] * values[static_cast<int64_t>(o[r] - 1) * nch + c]);
                }
            }
        }
    }
    for (int64_t v = (n16 << 4) + t0; v < nvox; v += nt) {
        const uint32_t m = mask[v];
        uint32_t o = 0;
        if (m) { o = owner[v]; if (o) owner[v] = 0; }
        for (int c = 0; c < nch; ++c)
            static_cast<T*>(out.p[c])[v] = o ? static_cast<T>(m * values[static_cast<int64_t>(o - 1) * nch + c]) : T(0);
    }
}

// ---- exact Euclidean distance transform of the DOWN-SAMPLED mask (blob_depthmap.py:174-181; a volume ~100x smaller
// than the segmentation, so a plain separable min-plus sweep is enough).  Matches
//     distance_transform_edt(np.pad(stack, 1), sampling)[1:-1, 1:-1, 1:-1]     (the reference's ndimage call)
// value for value: squared offsets are accumulated in the order z, y, x (ndimage sums ((ft - idx) * sampling)^2 over
// the axes in that order) and the zero padding appears as two virtual background positions (-1 and L) on every line.
__global__ void edt_pass_kernel(const double* __restrict__ in, const uint8_t* __restrict__ nonzero, double* __restrict__ out,
                                int64_t n, int64_t L, int64_t stride, double s, int take_sqrt) {
    const int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int64_t i = (v / stride) % L, base = v - i * stride;
    const double dl = static_cast<double>(i + 1) * s, dr = static_cast<double>(L - i) * s;
    double best = fmin(dl * dl, dr * dr);
    for (int64_t j = 0; j < L; ++j) {
        const double g = in ? in[base + j * stride] : (nonzero[base + j * stride] ? INFINITY : 0.0);
        if (g < best) {
            const double d = static_cast<double>(i - j) * s;
            best = fmin(best, g + d * d);
        }
    }
    out[v] = take_sqrt ? sqrt(best) : best;
}

static bool is_dev(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

int paint_boxes(Ctx* ctx, const void* mask_any, const int64_t shape[3], const int64_t* boxes_host, const int64_t* values_host,
                int64_t n, int nch, int elem_bytes, void* const* out_any, int64_t chunk_voxels) {
    const int64_t Z = shape[0], Y = shape[1], X = shape[2];
    // DLV_TRACE=1: host wall clock per phase on stderr (adds a stream synchronisation at every mark)
    static const bool trace = getenv("DLV_TRACE") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!trace) return;
        cudaStreamSynchronize(ctx->stream);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[dlv_paint] %-22s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };
    if (Z < 0 || Y < 0 || X < 0 || n < 0 || nch < 1 || nch > 3 || (elem_bytes != 1 && elem_bytes != 2)) {
        set_error(ctx, "dlv_paint_boxes: bad shape / channel count / element size");
        return DLV_ERR_ARG;
    }
    if (n >= 0xFFFFFFFFll) { set_error(ctx, "dlv_paint_boxes: more than 2^32 - 2 boxes"); return DLV_ERR_UNSUPPORTED; }
    const int64_t plane = Y * X;
    if (Z == 0 || plane == 0) return 0;
    const bool mask_dev = is_dev(mask_any);
    bool out_dev[3] = {false, false, false};
    for (int c = 0; c < nch; ++c) {
        if (!out_any[c]) { set_error(ctx, "dlv_paint_boxes: null output %d", c); return DLV_ERR_ARG; }
        out_dev[c] = is_dev(out_any[c]);
    }
    // Box list and values are narrowed (slice bounds clipped to int32 - anything larger is beyond the array anyway -
    // values modulo the output width, which is what numpy's cast of the int64 product does) straight into one pinned
    // block from the library's pool, on a few host threads: the upload then runs at PCIe speed instead of through the
    // driver's pageable staging (71 MB of int64 for the 1.3 M boxes of cfg3).
    const size_t nbox = static_cast<size_t>(std::max<int64_t>(n, 1));
    const size_t stage_bytes = nbox * 24 + nbox * nch * 2;
    size_t stage_cap = 0;
    uint8_t* stage = static_cast<uint8_t*>(pinned_take(stage_bytes, &stage_cap));
    if (!stage) { set_error(ctx, "dlv_paint_boxes: pinned staging allocation failed (%zu bytes)", stage_bytes); return DLV_ERR_CUDA; }
    int32_t* hbox = reinterpret_cast<int32_t*>(stage);
    uint16_t* hval = reinterpret_cast<uint16_t*>(stage + nbox * 24);
    {
        const uint64_t vmask = elem_bytes == 1 ? 0xFFu : 0xFFFFu;
        const int nthr = n > 200000 ? static_cast<int>(std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()))) : 1;
        std::vector<int64_t> bad(nthr, -1);
        auto work = [&](int t) {
            const int64_t i0 = n * t / nthr, i1 = n * (t + 1) / nthr;
            for (int64_t i = i0; i < i1; ++i) {
                for (int k = 0; k < 6; ++k) {
                    const int64_t v = boxes_host[6 * i + k];
                    if (v < 0 && bad[t] < 0) bad[t] = i;
                    hbox[6 * i + k] = static_cast<int32_t>(std::min<int64_t>(std::max<int64_t>(v, 0), INT32_MAX));
                }
                for (int c = 0; c < nch; ++c) hval[i * nch + c] = static_cast<uint16_t>(static_cast<uint64_t>(values_host[i * nch + c]) & vmask);
            }
        };
        if (nthr == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (int t = 0; t < nthr; ++t) th.emplace_back(work, t);
            for (auto& x : th) x.join();
        }
        for (int t = 0; t < nthr; ++t)
            if (bad[t] >= 0) {
                pinned_give(stage);
                set_error(ctx, "dlv_paint_boxes: negative slice bound in box %lld", static_cast<long long>(bad[t]));
                return DLV_ERR_ARG;
            }
    }

    mark("narrow to pinned");
    if (chunk_voxels <= 0) {
        // every z-chunk scans the whole box list, so the default is as few chunks as memory allows: the 4 B/voxel owner
        // scratch may take a quarter of the free device memory (cfg3: one chunk, 16.8 GB)
        size_t free_b = 0, total_b = 0;
        chunk_voxels = 1ll << 28;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) chunk_voxels = std::max<int64_t>(chunk_voxels, static_cast<int64_t>(free_b / 16));
        else cudaGetLastError();
    }
    const int64_t cz = std::max<int64_t>(1, std::min(Z, chunk_voxels / plane));
    const int64_t cvox = cz * plane;
    int32_t* boxes = nullptr; uint16_t* values = nullptr; uint32_t *owner = nullptr, *big = nullptr, *nbig = nullptr;
    uint8_t* mask_buf = nullptr; void* out_buf[3] = {nullptr, nullptr, nullptr};
    auto cleanup = [&]() {
        cudaStreamSynchronize(ctx->stream);          // the pinned block must not return to the pool while a copy reads it
        pinned_give(stage);
        dfree(ctx, boxes); dfree(ctx, values); dfree(ctx, big); dfree(ctx, nbig); dfree(ctx, mask_buf);
        for (int c = 0; c < 3; ++c) dfree(ctx, out_buf[c]);
    };
#define PAINT_OK(expr)                                                                                                   \
    do {                                                                                                                 \
        cudaError_t _e = (expr);                                                                                         \
        if (_e != cudaSuccess) { set_error(ctx, "dlv_paint_boxes: %s failed: %s", #expr, cudaGetErrorString(_e)); cleanup(); return DLV_ERR_CUDA; } \
    } while (0)
    PAINT_OK(dmalloc(ctx, &boxes, nbox * 24));
    PAINT_OK(dmalloc(ctx, &values, nbox * nch * 2));
    // The owner scratch is kept by the context between calls: paint_resolve leaves it all-zero, so only a new (or larger)
    // scratch - or one left behind by a failed call - has to be cleared (16.8 GB for cfg3: 3-4 ms per call otherwise).
    if (ctx->paint_owner_cap < static_cast<size_t>(cvox)) {
        dfree(ctx, ctx->paint_owner);
        ctx->paint_owner = nullptr; ctx->paint_owner_cap = 0;
        PAINT_OK(dmalloc(ctx, &ctx->paint_owner, static_cast<size_t>(cvox) * 4));
        ctx->paint_owner_cap = static_cast<size_t>(cvox);
        ctx->paint_owner_clean = false;
    }
    owner = ctx->paint_owner;
    if (!ctx->paint_owner_clean) PAINT_OK(cudaMemsetAsync(owner, 0, ctx->paint_owner_cap * 4, ctx->stream));
    ctx->paint_owner_clean = false;
    PAINT_OK(dmalloc(ctx, &big, static_cast<size_t>(std::max<int64_t>(n, 1)) * 4));
    PAINT_OK(dmalloc(ctx, &nbig, 4));
    if (!mask_dev) PAINT_OK(dmalloc(ctx, &mask_buf, static_cast<size_t>(cvox)));
    for (int c = 0; c < nch; ++c)
        if (!out_dev[c]) PAINT_OK(dmalloc(ctx, &out_buf[c], static_cast<size_t>(cvox) * elem_bytes));
    PAINT_OK(cudaMemcpyAsync(boxes, hbox, nbox * 24, cudaMemcpyHostToDevice, ctx->stream));
    PAINT_OK(cudaMemcpyAsync(values, hval, nbox * nch * 2, cudaMemcpyHostToDevice, ctx->stream));

    mark("alloc + upload");
    const int grid = ctx->num_sms * 8;
    for (int64_t z0 = 0; z0 < Z; z0 += cz) {
        const int64_t z1 = std::min(Z, z0 + cz), nv = (z1 - z0) * plane;
        const uint8_t* m = static_cast<const uint8_t*>(mask_any) + z0 * plane;
        if (!mask_dev) {
            PAINT_OK(cudaMemcpyAsync(mask_buf, m, static_cast<size_t>(nv), cudaMemcpyHostToDevice, ctx->stream));
            m = mask_buf;
        }
        PaintGeom g{Y, X, z0, z1};
        PaintOut po{{nullptr, nullptr, nullptr}};
        for (int c = 0; c < nch; ++c)
            po.p[c] = out_dev[c] ? static_cast<void*>(static_cast<uint8_t*>(out_any[c]) + z0 * plane * elem_bytes) : out_buf[c];
        if (n > 0) {
            PAINT_OK(cudaMemsetAsync(nbig, 0, 4, ctx->stream));
            paint_owner_small_kernel<<<grid, 256, 0, ctx->stream>>>(m, boxes, n, Z, g, owner, big, nbig);
            paint_owner_big_kernel<<<grid, 256, 0, ctx->stream>>>(m, boxes, Z, g, owner, big, nbig);
            ctx->launches += 2;
        }
        mark("owner kernels");
        int vec = (reinterpret_cast<uintptr_t>(m) & 15u) == 0;       // the owner scratch is always aligned
        for (int c = 0; c < nch; ++c) vec = vec && (reinterpret_cast<uintptr_t>(po.p[c]) & 15u) == 0;
        if (elem_bytes == 1) paint_resolve_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>(m, owner, nv, values, nch, po, vec);
        else paint_resolve_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>(m, owner, nv, values, nch, po, vec);
        ctx->launches += 1;
        mark("resolve kernel");
        for (int c = 0; c < nch; ++c)
            if (!out_dev[c])
                PAINT_OK(cudaMemcpyAsync(static_cast<uint8_t*>(out_any[c]) + z0 * plane * elem_bytes, out_buf[c],
                                         static_cast<size_t>(nv) * elem_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        if (!mask_dev) PAINT_OK(cudaStreamSynchronize(ctx->stream));      // mask_buf / out_buf are re-used by the next chunk
    }
    PAINT_OK(cudaGetLastError());
    PAINT_OK(cudaStreamSynchronize(ctx->stream));
#undef PAINT_OK
    ctx->paint_owner_clean = true;                   // every entry that was set has been reset by paint_resolve
    cleanup();
    mark("release");
    return 0;
}


int edt_run(Ctx* ctx, const void* nonzero_any, const int64_t shape[3], const double sampling[3], void* dist_any) {
    const int64_t Z = shape[0], Y = shape[1], X = shape[2], n = Z * Y * X;
    if (Z < 0 || Y < 0 || X < 0) { set_error(ctx, "dlv_edt: negative shape"); return DLV_ERR_ARG; }
    if (n == 0) return 0;
    uint8_t* nz = nullptr; double *a = nullptr, *b = nullptr;
    const bool in_dev = is_dev(nonzero_any), out_dev = is_dev(dist_any);
    auto cleanup = [&]() { if (!in_dev) dfree(ctx, nz); dfree(ctx, a); if (!out_dev) dfree(ctx, b); };
#define EDT_OK(expr)                                                                                                     \
    do {                                                                                                                 \
        cudaError_t _e = (expr);                                                                                         \
        if (_e != cudaSuccess) { set_error(ctx, "dlv_edt: %s failed: %s", #expr, cudaGetErrorString(_e)); cleanup(); return DLV_ERR_CUDA; } \
    } while (0)
    if (in_dev) nz = const_cast<uint8_t*>(static_cast<const uint8_t*>(nonzero_any));
    else {
        EDT_OK(dmalloc(ctx, &nz, static_cast<size_t>(n)));
        EDT_OK(cudaMemcpyAsync(nz, nonzero_any, static_cast<size_t>(n), cudaMemcpyHostToDevice, ctx->stream));
    }
    EDT_OK(dmalloc(ctx, &a, static_cast<size_t>(n) * 8));
    if (out_dev) b = static_cast<double*>(dist_any); else EDT_OK(dmalloc(ctx, &b, static_cast<size_t>(n) * 8));
    const unsigned grid = static_cast<unsigned>((n + 255) / 256);
    edt_pass_kernel<<<grid, 256, 0, ctx->stream>>>(nullptr, nz, b, n, Z, Y * X, sampling[0], 0);
    edt_pass_kernel<<<grid, 256, 0, ctx->stream>>>(b, nz, a, n, Y, X, sampling[1], 0);
    edt_pass_kernel<<<grid, 256, 0, ctx->stream>>>(a, nz, b, n, X, 1, sampling[2], 1);
    ctx->launches += 3;
    EDT_OK(cudaGetLastError());
    if (!out_dev) EDT_OK(cudaMemcpyAsync(dist_any, b, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    EDT_OK(cudaStreamSynchronize(ctx->stream));
#undef EDT_OK
    cleanup();
    return 0;
}

}  // namespace dlv

extern "C" int dlv_paint_boxes(dlv_ctx* c, const void* mask_any, const int64_t shape[3], const int64_t* boxes_host,
                               const int64_t* values_host, int64_t n, int nch, int elem_bytes, void* const* out_any,
                               int64_t chunk_voxels) {
    dlv::Ctx* ctx = reinterpret_cast<dlv::Ctx*>(c);
    if (!ctx) return DLV_ERR_ARG;
    if (!mask_any || !shape || !out_any || (n > 0 && (!boxes_host || !values_host))) {
        dlv::set_error(ctx, "dlv_paint_boxes: null argument");
        return DLV_ERR_ARG;
    }
    cudaSetDevice(ctx->device);
    return dlv::paint_boxes(ctx, mask_any, shape, boxes_host, values_host, n, nch, elem_bytes, out_any, chunk_voxels);
}

extern "C" int dlv_edt(dlv_ctx* c, const void* nonzero_any, const int64_t shape[3], const double sampling[3], void* dist_out_any) {
    dlv::Ctx* ctx = reinterpret_cast<dlv::Ctx*>(c);
    if (!ctx) return DLV_ERR_ARG;
    if (!nonzero_any || !shape || !sampling || !dist_out_any) { dlv::set_error(ctx, "dlv_edt: null argument"); return DLV_ERR_ARG; }
    cudaSetDevice(ctx->device);
    return dlv::edt_run(ctx, nonzero_any, shape, sampling, dist_out_any);
}
