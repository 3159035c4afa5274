// dlv_segment.cu - sliding-window driver: window enumeration, skip rule, batched U-Net passes with the
// fixed-point overlap blend, analytic averaging, then binarisation (dlv_post.cu).
//
// Replaces the compute of run_inference (inference/inference.py:229-329) and sliding_window_inference
// (inference/sliding_window_inferer.py:102-251) of the reference.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <vector>

#include "dlv_common.cuh"
#include "dlv_internal.h"

namespace dlv {

// ------------------------------------------------------------------- window grid (host)
// _get_scan_interval (sliding_window_inferer.py:255-276)
static int scan_interval(int64_t image, int roi, float overlap) {
    if (roi == image) return roi;
    const int iv = static_cast<int>(roi * (1.0 - static_cast<double>(overlap)));
    return iv > 0 ? iv : 1;
}
// per-dimension start list of MONAI 1.2.0 dense_patch_slices (sliding_window_inferer.py:143)
std::vector<int> window_starts(int64_t image, int roi, float overlap) {
    const int iv = scan_interval(image, roi, overlap);
    const int64_t num = (image + iv - 1) / iv;
    int scan_num = 1;
    for (int64_t d = 0; d < num; ++d)
        if (d * iv + roi >= image) { scan_num = static_cast<int>(d) + 1; break; }
    std::vector<int> s;
    for (int i = 0; i < scan_num; ++i) {
        int64_t st = static_cast<int64_t>(i) * iv;
        st -= std::max<int64_t>(st + roi - image, 0);
        s.push_back(static_cast<int>(st));
    }
    return s;
}

// MONAI compute_importance_map(mode="gaussian", sigma_scale=0.125): separable exp(-x^2 / (2 (0.125 n)^2)),
// normalised by its maximum; zeros clamped to the smallest non-zero value.  Optional mode (no reference oracle:
// the reference hard-codes mode='constant', sliding_window_inferer.py:148).
static std::vector<float> gaussian_1d(int n) {
    std::vector<float> w(n);
    const double sigma = 0.125 * n;
    for (int i = 0; i < n; ++i) {
        const double x = i - (n - 1) / 2.0;
        w[i] = static_cast<float>(std::exp(-(x * x) / (2.0 * sigma * sigma)));
    }
    return w;
}

// ------------------------------------------------------------------- averaging (inference.py:285-299)
// avg = (sum_active w*logit + sum_skipped w*(-1000)) / sum w   over all windows covering the voxel and all passes.
// Cover ranges per dimension come from small lookup tables; the skipped-window term is evaluated from the
// window-grid `active` flags instead of being blended, so empty windows cost no HBM traffic.
struct AvgArgs {
    int64_t PZ, PY, PX;        // region extent (planes of this slab, padded in-plane extent)
    int64_t gz0;               // global z of local plane 0
    const int32_t *lo_z, *hi_z, *lo_y, *hi_y, *lo_x, *hi_x;   // per-coordinate covering window index ranges (global coords)
    const int32_t *sz, *sy, *sx;                              // window starts
    int ny, nx;
    const int32_t* active;     // [nz][ny][nx]
    BlendDev bw;               // gaussian 1-D weights + per-axis normalisers (nz indexed by GLOBAL z here), or all null
    int passes;
};

// `acc` and `avg` are the SAME buffer in the library's calls (the average overwrites the sums in place, element by
// element, by the thread that read it): no __restrict__ on them.
__global__ void average_kernel(const int32_t* acc, float* avg, AvgArgs a) {
    const int64_t x = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t y = blockIdx.y, z = blockIdx.z;
    if (x >= a.PX) return;
    const int64_t gz = a.gz0 + z;
    float wsum = 0.f, wskip = 0.f;
    // the same weight expression as final_blend_kernel: (w_axis * normaliser_axis) per axis, then the product
    for (int iz = a.lo_z[gz]; iz <= a.hi_z[gz]; ++iz) {
        const float fz = a.bw.wz ? a.bw.wz[gz - a.sz[iz]] * a.bw.nz[gz] : 1.f;
        for (int iy = a.lo_y[y]; iy <= a.hi_y[y]; ++iy) {
            const float fy = a.bw.wz ? a.bw.wy[y - a.sy[iy]] * a.bw.ny[y] : 1.f;
            for (int ix = a.lo_x[x]; ix <= a.hi_x[x]; ++ix) {
                const float w = fz * fy * (a.bw.wz ? a.bw.wx[x - a.sx[ix]] * a.bw.nx[x] : 1.f);
                wsum += w;
                if (!a.active[(static_cast<int64_t>(iz) * a.ny + iy) * a.nx + ix]) wskip += w;
            }
        }
    }
    const int64_t i = (z * a.PY + y) * a.PX + x;
    const float s = static_cast<float>(acc[i]) * (1.f / kAccScale) + kSkipLogit * wskip * a.passes;
    avg[i] = s / (wsum * a.passes);     // acc and avg may alias (same element, same thread)
}

// Constant blend (what the reference computes): the number of windows covering a voxel, and how many of them were
// skipped, only change where a window starts or ends.  Each axis is cut into cells of constant (lo, hi); a small
// table holds (count, skipped) per 3-D cell and the volume pass is a pure stream: 16 B in, 16 B out per thread.
struct CellArgs {
    const int32_t *lo_z, *hi_z, *lo_y, *hi_y, *lo_x, *hi_x;   // per CELL
    int ncz, ncy, ncx, ny, nx;
    const int32_t* active;
};
__global__ void cell_table_kernel(CellArgs c, float2* __restrict__ table) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.ncz * c.ncy * c.ncx) return;
    const int cx = i % c.ncx, cy = (i / c.ncx) % c.ncy, cz = i / (c.ncx * c.ncy);
    float wsum = 0.f, wskip = 0.f;
    for (int iz = c.lo_z[cz]; iz <= c.hi_z[cz]; ++iz)
        for (int iy = c.lo_y[cy]; iy <= c.hi_y[cy]; ++iy)
            for (int ix = c.lo_x[cx]; ix <= c.hi_x[cx]; ++ix) {
                wsum += 1.f;
                if (!c.active[(static_cast<int64_t>(iz) * c.ny + iy) * c.nx + ix]) wskip += 1.f;
            }
    table[i] = make_float2(wsum, wskip);
}
__global__ void average_const_kernel(const int4* acc, float4* avg /* same buffer, see average_kernel */, int64_t PY, int64_t PX4, int64_t gz0,
                                     const int32_t* __restrict__ cid_z, const int32_t* __restrict__ cid_y,
                                     const int4* __restrict__ cid_x4, int ncy, int ncx, const float2* __restrict__ table, int passes) {
    const int64_t x4 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t y = blockIdx.y, z = blockIdx.z;
    if (x4 >= PX4) return;
    const float2* row = table + (static_cast<int64_t>(cid_z[gz0 + z]) * ncy + cid_y[y]) * ncx;
    const int4 cx = cid_x4[x4];
    const int64_t i = (z * PY + y) * PX4 + x4;
    const int4 v = acc[i];
    const float2 t0 = row[cx.x], t1 = row[cx.y], t2 = row[cx.z], t3 = row[cx.w];
    float4 o;      // same expression as average_kernel: (acc / 2^12 + (-1000) * skipped * passes) / (count * passes)
    o.x = (static_cast<float>(v.x) * (1.f / kAccScale) + kSkipLogit * t0.y * passes) / (t0.x * passes);
    o.y = (static_cast<float>(v.y) * (1.f / kAccScale) + kSkipLogit * t1.y * passes) / (t1.x * passes);
    o.z = (static_cast<float>(v.z) * (1.f / kAccScale) + kSkipLogit * t2.y * passes) / (t2.x * passes);
    o.w = (static_cast<float>(v.w) * (1.f / kAccScale) + kSkipLogit * t3.y * passes) / (t3.x * passes);
    avg[i] = o;      // acc and avg may alias (same 16 bytes, same thread)
}

struct DevBuf {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    ~DevBuf() { if (p) cudaFreeAsync(p, s); }
    template <class T> T* as() { return static_cast<T*>(p); }
};
struct Borrowed {              // context-owned grow-only scratch (scratch_get): same accessors as DevBuf, nothing to free
    void* p = nullptr;
    template <class T> T* as() { return static_cast<T*>(p); }
};
static int dev_alloc(Ctx* ctx, DevBuf& b, size_t bytes) {
    b.s = ctx->stream;
    DLV_CUDA_OK(ctx, dmalloc(ctx, &b.p, bytes));
    return 0;
}
template <class T>
static int dev_upload(Ctx* ctx, DevBuf& b, const std::vector<T>& v) {
    int rc = dev_alloc(ctx, b, v.size() * sizeof(T));
    if (rc) return rc;
    DLV_CUDA_OK(ctx, cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
static bool is_device_ptr(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

static void cover_tables(const std::vector<int>& starts, int roi, int64_t dim, std::vector<int32_t>& lo, std::vector<int32_t>& hi) {
    lo.assign(dim, 0); hi.assign(dim, -1);
    const int n = static_cast<int>(starts.size());
    for (int64_t g = 0; g < dim; ++g) {
        int l = n, h = -1;
        for (int i = 0; i < n; ++i)
            if (starts[i] <= g && g < starts[i] + roi) { l = std::min(l, i); h = std::max(h, i); }
        lo[g] = l; hi[g] = h;
    }
}

// ------------------------------------------------------------------- reusable stages (also exported per slab)
struct BlendWeights {
    DevBuf z, y, x, nz, ny, nx;
    BlendDev dev;              // nz indexed by GLOBAL z (callers offset it to their slab)
};
// reciprocal of the largest weight any covering window gives coordinate g (windows start at `starts`, extent roi)
static std::vector<float> axis_normaliser(const std::vector<int>& starts, int roi, int64_t dim, const std::vector<float>& w) {
    std::vector<float> r(dim, 1.f);
    for (int64_t g = 0; g < dim; ++g) {
        float m = 0.f;
        for (int s : starts)
            if (s <= g && g < s + roi) m = std::max(m, w[g - s]);
        r[g] = m > 0.f ? 1.f / m : 1.f;
    }
    return r;
}
static int blend_weights(Ctx* ctx, int blend_mode, const int32_t roi[3], const int64_t* shape_pad, float overlap, BlendWeights& w) {
    if (blend_mode == 0) return 0;
    if (blend_mode != 1) { set_error(ctx, "blend_mode must be 0 (constant) or 1 (gaussian)"); return DLV_ERR_ARG; }
    if (!shape_pad) { set_error(ctx, "the gaussian blend needs the window grid (padded shape, overlap, first plane of the slab)"); return DLV_ERR_ARG; }
    int rc;
    const std::vector<float> gz = gaussian_1d(roi[0]), gy = gaussian_1d(roi[1]), gx = gaussian_1d(roi[2]);
    if ((rc = dev_upload(ctx, w.z, gz)) || (rc = dev_upload(ctx, w.y, gy)) || (rc = dev_upload(ctx, w.x, gx))) return rc;
    if ((rc = dev_upload(ctx, w.nz, axis_normaliser(window_starts(shape_pad[0], roi[0], overlap), roi[0], shape_pad[0], gz))) ||
        (rc = dev_upload(ctx, w.ny, axis_normaliser(window_starts(shape_pad[1], roi[1], overlap), roi[1], shape_pad[1], gy))) ||
        (rc = dev_upload(ctx, w.nx, axis_normaliser(window_starts(shape_pad[2], roi[2], overlap), roi[2], shape_pad[2], gx))))
        return rc;
    w.dev.wz = w.z.as<float>(); w.dev.wy = w.y.as<float>(); w.dev.wx = w.x.as<float>();
    w.dev.nz = w.nz.as<float>(); w.dev.ny = w.ny.as<float>(); w.dev.nx = w.nx.as<float>();
    return 0;
}

// Library default for the windows per launch sequence: 128 windows of the reference's 96 x 96 x 64 (larger batches
// fill the 148 SMs on the low-resolution layers and shorten the share of every kernel's ramp-down tail: cfg2 0.449 /
// 0.464 / 0.466 / 0.470 Gvoxels/s at 32 / 64 / 96 / 128 windows), scaled inversely
// with the window volume and kept below half of the free device memory (~480 B of activations per window voxel).
static int default_window_batch(const int32_t roi[3]) {
    const int64_t vox = static_cast<int64_t>(roi[0]) * roi[1] * roi[2];
    int64_t b = std::max<int64_t>(1, std::min<int64_t>(512, (128LL * 96 * 96 * 64) / std::max<int64_t>(vox, 1)));
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
        b = std::max<int64_t>(1, std::min<int64_t>(b, static_cast<int64_t>(free_b / 2) / (480 * vox)));
    else
        cudaGetLastError();
    return static_cast<int>(b);
}

// run the scheduled windows (origins local to `slab`) and blend them into acc (int32, same extent as slab)
int seg_accumulate(Ctx* ctx, const uint16_t* slab, int64_t SY, int64_t SX, const std::vector<WindowDesc>& sched,
                   const int32_t roi[3], int batch, int blend_mode, int32_t* acc, const dlv_blend_geom* geom) {
    if (!ctx->net.loaded) { set_error(ctx, "call dlv_load_weights first"); return DLV_ERR_STATE; }
    if (sched.empty()) return 0;
    batch = batch > 0 ? batch : default_window_batch(roi);
    int rc = engine_prepare(ctx, roi, batch);
    if (rc) return rc;
    batch = engine_batch_capacity(ctx);
    BlendWeights bw;
    if ((rc = blend_weights(ctx, blend_mode, roi, geom ? geom->shape_pad : nullptr, geom ? geom->overlap : 0.f, bw))) return rc;
    if (blend_mode && (geom->gz0 < 0 || geom->gz0 >= geom->shape_pad[0] || SY != geom->shape_pad[1] || SX != geom->shape_pad[2])) {
        set_error(ctx, "gaussian blend: slab origin / in-plane extent inconsistent with the padded shape");
        return DLV_ERR_ARG;
    }
    BlendDev bdev = bw.dev;
    if (blend_mode) bdev.nz += geom->gz0;          // slab-local z indexing inside the window loop
    DevBuf d_sched;
    if ((rc = dev_upload(ctx, d_sched, sched))) return rc;
    for (size_t off = 0; off < sched.size() && rc == 0; off += batch) {
        const int n = static_cast<int>(std::min<size_t>(batch, sched.size() - off));
        rc = engine_run_batch(ctx, slab, SY, SX, d_sched.as<WindowDesc>() + off, n, acc, bdev, nullptr);
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);   // d_sched / weights are freed on return
    if (rc == 0 && e != cudaSuccess) { set_error(ctx, "accumulate: %s", cudaGetErrorString(e)); rc = DLV_ERR_CUDA; }
    return rc;
}

// int32 blend sums -> fp32 averaged logits, in place, for `nplanes` planes starting at global plane gz0
int seg_average(Ctx* ctx, int32_t* acc, int64_t nplanes, int64_t gz0, const int64_t shape_pad[3], const int32_t roi[3],
                float overlap, const int32_t* active_host, int passes, int blend_mode) {
    const int64_t PZ = shape_pad[0], PY = shape_pad[1], PX = shape_pad[2];
    if (nplanes <= 0) return 0;
    if (gz0 < 0 || gz0 + nplanes > PZ) { set_error(ctx, "average: plane range outside the padded volume"); return DLV_ERR_ARG; }
    const std::vector<int> sz = window_starts(PZ, roi[0], overlap), sy = window_starts(PY, roi[1], overlap),
                           sx = window_starts(PX, roi[2], overlap);
    const int64_t nwin = static_cast<int64_t>(sz.size()) * sy.size() * sx.size();
    BlendWeights bw;
    int rc;
    if ((rc = blend_weights(ctx, blend_mode, roi, shape_pad, overlap, bw))) return rc;
    DevBuf t_loz, t_hiz, t_loy, t_hiy, t_lox, t_hix, t_sz, t_sy, t_sx, d_active;
    std::vector<int32_t> lo, hi;
    cover_tables(sz, roi[0], PZ, lo, hi);
    if ((rc = dev_upload(ctx, t_loz, lo)) || (rc = dev_upload(ctx, t_hiz, hi))) return rc;
    cover_tables(sy, roi[1], PY, lo, hi);
    if ((rc = dev_upload(ctx, t_loy, lo)) || (rc = dev_upload(ctx, t_hiy, hi))) return rc;
    cover_tables(sx, roi[2], PX, lo, hi);
    if ((rc = dev_upload(ctx, t_lox, lo)) || (rc = dev_upload(ctx, t_hix, hi))) return rc;
    std::vector<int32_t> s32(sz.begin(), sz.end());
    if ((rc = dev_upload(ctx, t_sz, s32))) return rc;
    s32.assign(sy.begin(), sy.end());
    if ((rc = dev_upload(ctx, t_sy, s32))) return rc;
    s32.assign(sx.begin(), sx.end());
    if ((rc = dev_upload(ctx, t_sx, s32))) return rc;
    std::vector<int32_t> act(active_host, active_host + nwin);
    if ((rc = dev_upload(ctx, d_active, act))) return rc;
    if (blend_mode == 0 && (PX & 3) == 0 && (reinterpret_cast<uintptr_t>(acc) & 15u) == 0) {
        // cells of constant window cover per axis
        auto cells = [](const std::vector<int32_t>& lo_v, const std::vector<int32_t>& hi_v, std::vector<int32_t>& cid,
                        std::vector<int32_t>& clo, std::vector<int32_t>& chi) {
            cid.resize(lo_v.size()); clo.clear(); chi.clear();
            for (size_t g = 0; g < lo_v.size(); ++g) {
                if (g == 0 || lo_v[g] != lo_v[g - 1] || hi_v[g] != hi_v[g - 1]) { clo.push_back(lo_v[g]); chi.push_back(hi_v[g]); }
                cid[g] = static_cast<int32_t>(clo.size()) - 1;
            }
        };
        std::vector<int32_t> lz, hz, ly, hy, lx, hx, cidz, cidy, cidx, cloz, chiz, cloy, chiy, clox, chix;
        cover_tables(sz, roi[0], PZ, lz, hz); cells(lz, hz, cidz, cloz, chiz);
        cover_tables(sy, roi[1], PY, ly, hy); cells(ly, hy, cidy, cloy, chiy);
        cover_tables(sx, roi[2], PX, lx, hx); cells(lx, hx, cidx, clox, chix);
        DevBuf d_cidz, d_cidy, d_cidx, d_cloz, d_chiz, d_cloy, d_chiy, d_clox, d_chix, d_table;
        if ((rc = dev_upload(ctx, d_cidz, cidz)) || (rc = dev_upload(ctx, d_cidy, cidy)) || (rc = dev_upload(ctx, d_cidx, cidx)) ||
            (rc = dev_upload(ctx, d_cloz, cloz)) || (rc = dev_upload(ctx, d_chiz, chiz)) || (rc = dev_upload(ctx, d_cloy, cloy)) ||
            (rc = dev_upload(ctx, d_chiy, chiy)) || (rc = dev_upload(ctx, d_clox, clox)) || (rc = dev_upload(ctx, d_chix, chix)))
            return rc;
        CellArgs c;
        c.lo_z = d_cloz.as<int32_t>(); c.hi_z = d_chiz.as<int32_t>(); c.lo_y = d_cloy.as<int32_t>(); c.hi_y = d_chiy.as<int32_t>();
        c.lo_x = d_clox.as<int32_t>(); c.hi_x = d_chix.as<int32_t>();
        c.ncz = static_cast<int>(cloz.size()); c.ncy = static_cast<int>(cloy.size()); c.ncx = static_cast<int>(clox.size());
        c.ny = static_cast<int>(sy.size()); c.nx = static_cast<int>(sx.size()); c.active = d_active.as<int32_t>();
        const int64_t ncell = static_cast<int64_t>(c.ncz) * c.ncy * c.ncx;
        if (ncell < (1ll << 31)) {
            if ((rc = dev_alloc(ctx, d_table, static_cast<size_t>(ncell) * sizeof(float2)))) return rc;
            cell_table_kernel<<<static_cast<unsigned>((ncell + 255) / 256), 256, 0, ctx->stream>>>(c, d_table.as<float2>());
            dim3 grid(static_cast<unsigned>((PX / 4 + 127) / 128), static_cast<unsigned>(PY), static_cast<unsigned>(nplanes));
            average_const_kernel<<<grid, 128, 0, ctx->stream>>>(reinterpret_cast<const int4*>(acc), reinterpret_cast<float4*>(acc), PY, PX / 4,
                                                                gz0, d_cidz.as<int32_t>(), d_cidy.as<int32_t>(),
                                                                reinterpret_cast<const int4*>(d_cidx.as<int32_t>()), c.ncy, c.ncx,
                                                                d_table.as<float2>(), passes);
            ctx->launches += 2;
            cudaError_t e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);   // lookup tables are freed on return
            if (e != cudaSuccess) { set_error(ctx, "average kernel: %s", cudaGetErrorString(e)); return DLV_ERR_CUDA; }
            return 0;
        }
    }
    AvgArgs a;
    a.PZ = nplanes; a.PY = PY; a.PX = PX; a.gz0 = gz0;
    a.lo_z = t_loz.as<int32_t>(); a.hi_z = t_hiz.as<int32_t>();
    a.lo_y = t_loy.as<int32_t>(); a.hi_y = t_hiy.as<int32_t>();
    a.lo_x = t_lox.as<int32_t>(); a.hi_x = t_hix.as<int32_t>();
    a.sz = t_sz.as<int32_t>(); a.sy = t_sy.as<int32_t>(); a.sx = t_sx.as<int32_t>();
    a.ny = static_cast<int>(sy.size()); a.nx = static_cast<int>(sx.size()); a.active = d_active.as<int32_t>();
    a.bw = bw.dev; a.passes = passes;
    dim3 grid(static_cast<unsigned>((PX + 255) / 256), static_cast<unsigned>(PY), static_cast<unsigned>(nplanes));
    average_kernel<<<grid, 256, 0, ctx->stream>>>(acc, reinterpret_cast<float*>(acc), a);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);   // lookup tables are freed on return
    if (e != cudaSuccess) { set_error(ctx, "average kernel: %s", cudaGetErrorString(e)); return DLV_ERR_CUDA; }
    return 0;
}

// The 13-pass test-time-augmentation plan of inference.py:265-279: one plain pass, then 4 x {plain, flip z, flip y}
// (the reference adds unseeded N(0, U(0,1e-3)) noise to raw intensities >= 1 in the 12 extra passes; that is below
// the resolution of the arithmetic and not reproducible, so the passes are evaluated noise-free).  Noise-free, the 13
// passes are 5 x plain + 4 x flip z + 4 x flip y; every pass is deterministic and the blend is an integer sum, so
// each distinct pass is evaluated once and added `repeat` times - bit-identical to running all 13 (tested).
static const int kTtaFlips[3] = {0, 1, 2};
static const int kTtaRepeat[3] = {5, 4, 4};

int segment_run(Ctx* ctx, const void* volume_any, const dlv_seg_params* P, void* binaries_any, void* avg_any, void* sig_any,
                dlv_seg_stats* st_out) {
    if (!ctx->net.loaded) { set_error(ctx, "dlv_segment: call dlv_load_weights first"); return DLV_ERR_STATE; }
    const int64_t PZ = P->shape_pad[0], PY = P->shape_pad[1], PX = P->shape_pad[2];
    const int64_t Z = P->shape_real[0], Y = P->shape_real[1], X = P->shape_real[2];
    for (int i = 0; i < 3; ++i) {
        if (P->roi[i] <= 0 || P->shape_pad[i] < P->roi[i] || P->shape_real[i] <= 0 || P->shape_real[i] > P->shape_pad[i]) {
            set_error(ctx, "dlv_segment: inconsistent shapes (dim %d: real %lld, padded %lld, roi %d)", i,
                      (long long)P->shape_real[i], (long long)P->shape_pad[i], P->roi[i]);
            return DLV_ERR_ARG;
        }
    }
    if (P->overlap < 0.f || P->overlap >= 1.f) { set_error(ctx, "overlap must be >= 0 and < 1."); return DLV_ERR_ARG; }
    const int batch = P->window_batch > 0 ? P->window_batch : default_window_batch(P->roi);
    // DLV_TRACE=1: host wall clock per phase on stderr (adds a stream synchronisation at every mark)
    static const bool trace = getenv("DLV_TRACE") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!trace) return;
        cudaStreamSynchronize(ctx->stream);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[dlv_segment] %-22s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };
    int rc = engine_prepare(ctx, P->roi, batch);
    if (rc) return rc;
    mark("engine_prepare");

    const size_t nvox_pad = static_cast<size_t>(PZ) * PY * PX;
    const size_t nvox = static_cast<size_t>(Z) * Y * X;

    // ---- window grid, z-major / x fastest like dense_patch_slices
    const std::vector<int> sz = window_starts(PZ, P->roi[0], P->overlap), sy = window_starts(PY, P->roi[1], P->overlap),
                           sx = window_starts(PX, P->roi[2], P->overlap);
    const int nz = sz.size(), ny = sy.size(), nx = sx.size();
    const int64_t nwin = static_cast<int64_t>(nz) * ny * nx, per_layer = static_cast<int64_t>(ny) * nx;
    std::vector<int32_t> origins(nwin * 3);
    for (int iz = 0, w = 0; iz < nz; ++iz)
        for (int iy = 0; iy < ny; ++iy)
            for (int ix = 0; ix < nx; ++ix, ++w) { origins[3 * w] = sz[iz]; origins[3 * w + 1] = sy[iy]; origins[3 * w + 2] = sx[ix]; }
    const int passes = P->tta ? 13 : 1;
    int single_flip = 0;
    if (!P->tta && P->flip_dim) {
        if (P->flip_dim < 2 || P->flip_dim > 4) { set_error(ctx, "flip_dim must be 0, 2 (z), 3 (y) or 4 (x)"); return DLV_ERR_ARG; }
        single_flip = P->flip_dim - 1;
    }
    const int distinct = P->tta ? 3 : 1;
    auto push_window = [&](std::vector<WindowDesc>& out, int64_t w, int ps) {
        out.push_back(WindowDesc{origins[3 * w], origins[3 * w + 1], origins[3 * w + 2],
                                 P->tta ? (kTtaFlips[ps] | ((kTtaRepeat[ps] - 1) << 8)) : single_flip});
    };

    // ---- input slab on the device.  A host volume is uploaded in z chunks on a second stream - chunk k ends where
    // window z-layer k ends - and layer k's skip scan and window batches start as soon as their chunk has landed, so
    // everything but the first chunk of the upload hides behind the network (pinned host memory; a pageable volume
    // is staged by the driver and overlaps only with batches that are already enqueued).
    DevBuf d_orig, d_active;
    Borrowed slab_own, d_acc, bin_own;
    const uint16_t* slab = static_cast<const uint16_t*>(volume_any);
    const bool host_volume = !is_device_ptr(volume_any);
    if ((rc = dev_upload(ctx, d_orig, origins))) return rc;
    if ((rc = dev_alloc(ctx, d_active, nwin * sizeof(int32_t)))) return rc;
    std::vector<int32_t> active(nwin, 1);
    DLV_CUDA_OK(ctx, scratch_get(ctx, kScratchAcc, nvox_pad * 4, &d_acc.p));
    DLV_CUDA_OK(ctx, cudaMemsetAsync(d_acc.p, 0, nvox_pad * 4, ctx->stream));
    struct Events {           // destroyed on every return path
        std::vector<cudaEvent_t> e;
        explicit Events(int n) : e(n, nullptr) { for (auto& x : e) cudaEventCreateWithFlags(&x, cudaEventDisableTiming); }
        ~Events() { for (auto& x : e) if (x) cudaEventDestroy(x); }
    };
    struct TimingEvents {
        cudaEvent_t e[3] = {nullptr, nullptr, nullptr};
        TimingEvents() { for (auto& x : e) cudaEventCreate(&x); }
        ~TimingEvents() { for (auto& x : e) if (x) cudaEventDestroy(x); }
    } tev;
    cudaEvent_t e0 = tev.e[0], e1 = tev.e[1], e2 = tev.e[2];
    dlv_blend_geom geom;
    geom.shape_pad[0] = PZ; geom.shape_pad[1] = PY; geom.shape_pad[2] = PX; geom.overlap = P->overlap; geom.gz0 = 0;
    const int64_t launches0 = ctx->launches;
    ctx->conv_ms = 0.0;
    int64_t nactive = 0;
    cudaEventRecord(e0, ctx->stream);
    if (!host_volume) {
        if (P->skip_empty) {
            if ((rc = windows_active(ctx, slab, PY, PX, d_orig.as<int32_t>(), static_cast<int>(nwin), P->roi, d_active.as<int32_t>()))) return rc;
            DLV_CUDA_OK(ctx, cudaMemcpyAsync(active.data(), d_active.p, nwin * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
            DLV_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        }
        mark("skip-rule scan");
        // ---- schedule: passes x active windows
        std::vector<WindowDesc> sched;
        for (int64_t w = 0; w < nwin; ++w) nactive += active[w] != 0;
        sched.reserve(nactive * distinct);
        for (int ps = 0; ps < distinct; ++ps)
            for (int64_t w = 0; w < nwin; ++w)
                if (active[w]) push_window(sched, w, ps);
        rc = seg_accumulate(ctx, slab, PY, PX, sched, P->roi, batch, P->blend_mode, d_acc.as<int32_t>(), &geom);
    } else {
        DLV_CUDA_OK(ctx, scratch_get(ctx, kScratchSlab, nvox_pad * 2, &slab_own.p));
        slab = slab_own.as<uint16_t>();
        Events landed(nz);
        // the allocation (stream-ordered on ctx->stream) must precede the first copy on the copy stream
        cudaEvent_t alloc_done = nullptr;
        cudaEventCreateWithFlags(&alloc_done, cudaEventDisableTiming);
        cudaEventRecord(alloc_done, ctx->stream);
        cudaStreamWaitEvent(ctx->copy_stream, alloc_done, 0);
        cudaEventDestroy(alloc_done);
        const size_t plane_bytes = static_cast<size_t>(PY) * PX * 2;
        int64_t copied = 0;                                     // planes already enqueued
        auto upload_layer = [&](int k) -> int {                 // chunk k: up to the end of window layer k (the last one takes the rest)
            const int64_t end = (k == nz - 1) ? PZ : std::min<int64_t>(PZ, static_cast<int64_t>(sz[k]) + P->roi[0]);
            if (end > copied) {
                DLV_CUDA_OK(ctx, cudaMemcpyAsync(slab_own.as<uint8_t>() + copied * plane_bytes, static_cast<const uint8_t*>(volume_any) + copied * plane_bytes,
                                                 static_cast<size_t>(end - copied) * plane_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
                copied = end;
            }
            DLV_CUDA_OK(ctx, cudaEventRecord(landed.e[k], ctx->copy_stream));
            return 0;
        };
        if ((rc = upload_layer(0))) return rc;
        std::vector<WindowDesc> pending;                        // active windows whose planes have landed, not yet run
        const size_t bcap = static_cast<size_t>(engine_batch_capacity(ctx));
        for (int k = 0; k < nz && rc == 0; ++k) {
            if (k + 1 < nz && (rc = upload_layer(k + 1))) return rc;
            DLV_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, landed.e[k], 0));
            const int64_t w0 = static_cast<int64_t>(k) * per_layer;
            if (P->skip_empty) {
                if ((rc = windows_active(ctx, slab, PY, PX, d_orig.as<int32_t>() + 3 * w0, static_cast<int>(per_layer), P->roi,
                                         d_active.as<int32_t>() + w0))) return rc;
                DLV_CUDA_OK(ctx, cudaMemcpyAsync(active.data() + w0, d_active.as<int32_t>() + w0, per_layer * sizeof(int32_t),
                                                 cudaMemcpyDeviceToHost, ctx->stream));
                DLV_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
            }
            for (int ps = 0; ps < distinct; ++ps)
                for (int64_t w = w0; w < w0 + per_layer; ++w)
                    if (active[w]) { push_window(pending, w, ps); nactive += ps == 0; }
            // whole batches now; what is left over joins the next layer's windows (window order does not matter:
            // the blend is an integer sum)
            const size_t run = (k == nz - 1) ? pending.size() : (bcap ? pending.size() / bcap * bcap : pending.size());
            if (run) {
                std::vector<WindowDesc> now(pending.begin(), pending.begin() + run);
                pending.erase(pending.begin(), pending.begin() + run);
                rc = seg_accumulate(ctx, slab, PY, PX, now, P->roi, batch, P->blend_mode, d_acc.as<int32_t>(), &geom);
            }
        }
        cudaStreamSynchronize(ctx->copy_stream);
        mark("upload + skip-rule scan");
    }
    cudaEventRecord(e1, ctx->stream);
    mark("u-net passes");

    // ---- average (in place: int32 sums -> fp32 logits)
    if (rc == 0) rc = seg_average(ctx, d_acc.as<int32_t>(), PZ, 0, P->shape_pad, P->roi, P->overlap, active.data(), passes, P->blend_mode);

    mark("average");
    // ---- binarise + eroded-mask gate
    DevBuf sig_own;
    uint8_t* bin = static_cast<uint8_t*>(binaries_any);
    float* sig = static_cast<float*>(sig_any);
    const bool bin_host = !is_device_ptr(binaries_any);
    const bool sig_host = sig_any && !is_device_ptr(sig_any);
    if (rc == 0 && bin_host) { DLV_CUDA_OK(ctx, scratch_get(ctx, kScratchBin, nvox, &bin_own.p)); bin = bin_own.as<uint8_t>(); }
    if (rc == 0 && sig_host) { if ((rc = dev_alloc(ctx, sig_own, nvox * 4))) return rc; sig = sig_own.as<float>(); }
    if (rc == 0)
        rc = post_finalise(ctx, d_acc.as<float>(), slab, P->shape_pad, P->shape_real, P->threshold, P->erosion_iters,
                           P->erosion_block_planes, bin, sig);
    cudaEventRecord(e2, ctx->stream);
    mark("finalise");
    if (rc == 0) {
        if (bin_host) cudaMemcpyAsync(binaries_any, bin, nvox, cudaMemcpyDeviceToHost, ctx->stream);
        if (sig_host) cudaMemcpyAsync(sig_any, sig, nvox * 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (avg_any)
            cudaMemcpyAsync(avg_any, d_acc.p, nvox_pad * 4, is_device_ptr(avg_any) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream);
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (rc == 0 && e != cudaSuccess) { set_error(ctx, "dlv_segment: %s", cudaGetErrorString(e)); rc = DLV_ERR_CUDA; }
    mark("outputs");
    if (rc == 0 && st_out) {
        float ms = 0.f;
        st_out->windows_total = nwin;
        st_out->windows_active = nactive;
        st_out->passes = passes;
        st_out->kernel_launches = ctx->launches - launches0;
        cudaEventElapsedTime(&ms, e0, e1); st_out->ms_unet = ms;
        cudaEventElapsedTime(&ms, e1, e2); st_out->ms_finalise = ms;
        st_out->ms_conv = ctx->conv_ms;
    }
    return rc;
}

}  // namespace dlv
