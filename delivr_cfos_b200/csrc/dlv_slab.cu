// dlv_slab.cu - slab-level entry points for z-sharded runs (one process per GPU, NCCL between them):
// window grid, per-slab accumulate / average / finalise, boundary label pairs and relabelling for the
// cross-slab component merge.  The host side that strings these together is delivr_cfos_b200/slabs.py.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "dlv_common.cuh"
#include "dlv_internal.h"

namespace dlv {

std::vector<int> window_starts(int64_t image, int roi, float overlap);
int seg_accumulate(Ctx* ctx, const uint16_t* slab, int64_t SY, int64_t SX, const std::vector<WindowDesc>& sched,
                   const int32_t roi[3], int batch, int blend_mode, int32_t* acc, const dlv_blend_geom* geom);
int seg_average(Ctx* ctx, int32_t* acc, int64_t nplanes, int64_t gz0, const int64_t shape_pad[3], const int32_t roi[3],
                float overlap, const int32_t* active_host, int passes, int blend_mode);

// 26-adjacent foreground pairs across a slab boundary: `lo` is the last plane of the lower slab, `hi` the first
// plane of the upper slab (local labels, 0 = background).  One thread per voxel of `hi`; a pair is skipped when
// the same (dy,dx) offset produced the identical pair one voxel to the left (cheap run-level de-duplication).
__global__ void boundary_pairs_kernel(const uint32_t* __restrict__ lo, const uint32_t* __restrict__ hi, int64_t Y, int64_t X,
                                      uint32_t* __restrict__ pairs, unsigned long long cap, unsigned long long* __restrict__ count) {
    const int64_t x = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t y = blockIdx.y;
    if (x >= X) return;
    const uint32_t b = hi[y * X + x];
    if (!b) return;
    const uint32_t bprev = (x > 0) ? hi[y * X + x - 1] : 0u;
    for (int dy = -1; dy <= 1; ++dy) {
        const int64_t yy = y + dy;
        if (yy < 0 || yy >= Y) continue;
        for (int dx = -1; dx <= 1; ++dx) {
            const int64_t xx = x + dx;
            if (xx < 0 || xx >= X) continue;
            const uint32_t a = lo[yy * X + xx];
            if (!a) continue;
            if (bprev == b && xx > 0 && lo[yy * X + xx - 1] == a) continue;
            const unsigned long long i = atomicAdd(count, 1ull);
            if (i < cap) { pairs[2 * i] = a; pairs[2 * i + 1] = b; }
        }
    }
}

// `head` = elements in front of the first 16-byte boundary (a sub-slab of a volume whose planes are not a multiple of
// 4 voxels starts anywhere): they and the tail are handled one by one, the aligned middle four at a time.
__global__ void relabel_kernel(uint32_t* __restrict__ labels, int64_t n, const uint32_t* __restrict__ map, int head) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t i = head + t * 4;
    if (t == 0)
        for (int64_t j = 0; j < head && j < n; ++j) { const uint32_t l = labels[j]; if (l) labels[j] = map[l]; }
    if (i + 3 < n) {
        uint4 v = *reinterpret_cast<uint4*>(labels + i);
        if (v.x | v.y | v.z | v.w) {
            v.x = v.x ? map[v.x] : 0u; v.y = v.y ? map[v.y] : 0u; v.z = v.z ? map[v.z] : 0u; v.w = v.w ? map[v.w] : 0u;
            *reinterpret_cast<uint4*>(labels + i) = v;
        }
    } else {
        for (int64_t j = i; j < n; ++j) { const uint32_t l = labels[j]; if (l) labels[j] = map[l]; }
    }
}

}  // namespace dlv

using dlv::Ctx;
static Ctx* C(dlv_ctx* c) { return reinterpret_cast<Ctx*>(c); }

extern "C" {

int dlv_window_grid(const int64_t shape_pad[3], const int32_t roi[3], float overlap, int32_t counts_out[3], int32_t* starts_out) {
    if (!shape_pad || !roi || !counts_out) return DLV_ERR_ARG;
    int off = 0;
    for (int d = 0; d < 3; ++d) {
        if (roi[d] <= 0 || shape_pad[d] < roi[d] || overlap < 0.f || overlap >= 1.f) return DLV_ERR_ARG;
        const std::vector<int> s = dlv::window_starts(shape_pad[d], roi[d], overlap);
        counts_out[d] = static_cast<int32_t>(s.size());
        if (starts_out) for (size_t i = 0; i < s.size(); ++i) starts_out[off + i] = s[i];
        off += static_cast<int>(s.size());
    }
    return DLV_OK;
}

int dlv_windows_active(dlv_ctx* c, const uint16_t* slab_dev, int64_t SY, int64_t SX, const int32_t* origins_host, int n,
                       const int32_t roi[3], int32_t* active_host) {
    Ctx* ctx = C(c);
    if (!ctx) return DLV_ERR_ARG;
    if (n <= 0) return DLV_OK;
    cudaSetDevice(ctx->device);
    int32_t *d_o = nullptr, *d_a = nullptr;
    DLV_CUDA_OK(ctx, dlv::dmalloc(ctx, &d_o, sizeof(int32_t) * 3 * n));
    DLV_CUDA_OK(ctx, dlv::dmalloc(ctx, &d_a, sizeof(int32_t) * n));
    cudaMemcpyAsync(d_o, origins_host, sizeof(int32_t) * 3 * n, cudaMemcpyHostToDevice, ctx->stream);
    int rc = dlv::windows_active(ctx, slab_dev, SY, SX, d_o, n, roi, d_a);
    cudaMemcpyAsync(active_host, d_a, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    dlv::dfree(ctx, d_o); dlv::dfree(ctx, d_a);
    if (rc == 0 && e != cudaSuccess) { dlv::set_error(ctx, "dlv_windows_active: %s", cudaGetErrorString(e)); rc = DLV_ERR_CUDA; }
    return rc;
}

int dlv_seg_accumulate(dlv_ctx* c, const uint16_t* slab_dev, int64_t SY, int64_t SX, const int32_t* windows_host, int n,
                       const int32_t roi[3], int window_batch, int blend_mode, const dlv_blend_geom* geom_or_null, int32_t* acc_dev) {
    Ctx* ctx = C(c);
    if (!ctx) return DLV_ERR_ARG;
    if (!slab_dev || !roi || !acc_dev || (n > 0 && !windows_host)) { dlv::set_error(ctx, "dlv_seg_accumulate: null argument"); return DLV_ERR_ARG; }
    cudaSetDevice(ctx->device);
    std::vector<dlv::WindowDesc> sched(n > 0 ? n : 0);
    for (int i = 0; i < n; ++i) {
        const int32_t f = windows_host[4 * i + 3] & 0xFF, rep1 = windows_host[4 * i + 3] >> 8;
        if ((f != 0 && (f < 2 || f > 4)) || rep1 < 0 || rep1 > 255) {
            dlv::set_error(ctx, "dlv_seg_accumulate: flip_dim must be 0, 2, 3 or 4 and repeat 1..256");
            return DLV_ERR_ARG;
        }
        sched[i] = dlv::WindowDesc{windows_host[4 * i], windows_host[4 * i + 1], windows_host[4 * i + 2], (f ? f - 1 : 0) | (rep1 << 8)};
    }
    if (blend_mode != 0 && !geom_or_null) { dlv::set_error(ctx, "dlv_seg_accumulate: the gaussian blend needs dlv_blend_geom"); return DLV_ERR_ARG; }
    return dlv::seg_accumulate(ctx, slab_dev, SY, SX, sched, roi, window_batch, blend_mode, acc_dev, geom_or_null);
}

int dlv_seg_average(dlv_ctx* c, int32_t* acc_dev_inout, int64_t nplanes, int64_t gz0, const int64_t shape_pad[3],
                    const int32_t roi[3], float overlap, const int32_t* active_host, int passes, int blend_mode) {
    Ctx* ctx = C(c);
    if (!ctx) return DLV_ERR_ARG;
    if (!acc_dev_inout || !shape_pad || !roi || !active_host || passes < 1) { dlv::set_error(ctx, "dlv_seg_average: bad argument"); return DLV_ERR_ARG; }
    cudaSetDevice(ctx->device);
    return dlv::seg_average(ctx, acc_dev_inout, nplanes, gz0, shape_pad, roi, overlap, active_host, passes, blend_mode);
}

int dlv_op_finalise_slab(dlv_ctx* c, const float* avg_dev, const uint16_t* volume_dev, int64_t SY, int64_t SX, int64_t nplanes,
                         int64_t gz0, const int64_t shape_real[3], float threshold, int erosion_iters, int64_t erosion_block_planes,
                         int64_t oz0, int64_t oz1, uint8_t* binaries_dev, float* sigmoid_dev_or_null) {
    Ctx* ctx = C(c);
    if (!ctx) return DLV_ERR_ARG;
    cudaSetDevice(ctx->device);
    // only real planes take part in the erosion (outside the array counts as 1)
    const int64_t np_real = std::min<int64_t>(nplanes, shape_real[0] - gz0);
    return dlv::post_finalise_slab(ctx, avg_dev, volume_dev, SY, SX, np_real, gz0, shape_real, threshold, erosion_iters,
                                   erosion_block_planes, oz0, oz1, binaries_dev, sigmoid_dev_or_null);
}

int dlv_ccl_boundary_pairs(dlv_ctx* c, const uint32_t* labels_lo_plane_dev, const uint32_t* labels_hi_plane_dev, int64_t Y, int64_t X,
                           uint32_t* pairs_dev, int64_t cap, int64_t* count_host_out) {
    Ctx* ctx = C(c);
    if (!ctx) return DLV_ERR_ARG;
    if (!labels_lo_plane_dev || !labels_hi_plane_dev || !pairs_dev || !count_host_out || cap < 0) { dlv::set_error(ctx, "dlv_ccl_boundary_pairs: bad argument"); return DLV_ERR_ARG; }
    cudaSetDevice(ctx->device);
    unsigned long long* cnt = nullptr;
    DLV_CUDA_OK(ctx, dlv::dmalloc(ctx, &cnt, 8));
    cudaMemsetAsync(cnt, 0, 8, ctx->stream);
    if (Y > 0 && X > 0) {
        dim3 grid(static_cast<unsigned>((X + 127) / 128), static_cast<unsigned>(Y));
        dlv::boundary_pairs_kernel<<<grid, 128, 0, ctx->stream>>>(labels_lo_plane_dev, labels_hi_plane_dev, Y, X, pairs_dev,
                                                                 static_cast<unsigned long long>(cap), cnt);
        ctx->launches++;
    }
    unsigned long long h = 0;
    cudaMemcpyAsync(&h, cnt, 8, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    dlv::dfree(ctx, cnt);
    if (e != cudaSuccess) { dlv::set_error(ctx, "dlv_ccl_boundary_pairs: %s", cudaGetErrorString(e)); return DLV_ERR_CUDA; }
    *count_host_out = static_cast<int64_t>(h);
    return DLV_OK;
}

int dlv_relabel(dlv_ctx* c, uint32_t* labels_dev, int64_t n, const uint32_t* map_dev, int64_t nmap) {
    Ctx* ctx = C(c);
    if (!ctx) return DLV_ERR_ARG;
    if (!labels_dev || !map_dev || n < 0 || nmap < 1) { dlv::set_error(ctx, "dlv_relabel: bad argument"); return DLV_ERR_ARG; }
    cudaSetDevice(ctx->device);
    if (n > 0) {
        const int head = static_cast<int>(((16 - (reinterpret_cast<uintptr_t>(labels_dev) & 15u)) & 15u) / 4);
        const int64_t nthreads = (n + 3) / 4 + 1;
        dlv::relabel_kernel<<<static_cast<unsigned>((nthreads + 255) / 256), 256, 0, ctx->stream>>>(labels_dev, n, map_dev, head);
        ctx->launches++;
    }
    DLV_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return DLV_OK;
}

/* Host-only: exact merge of per-slab statistics tables into the global table (see include/delivr_b200.h). */
int dlv_table_merge(int64_t n_global, int ntables, const int64_t* rows, const uint32_t* const* luts,
                    const uint64_t* const* counts, const uint64_t* const* sums, const int64_t* const* bbox,
                    const int64_t* z_offsets, const int64_t shape[3], uint64_t* counts_out, uint64_t* sums_out,
                    int64_t* bbox_out, double* centroids_out) {
    if (n_global < 0 || ntables < 0 || !shape || !counts_out || !sums_out || !bbox_out || !centroids_out) return DLV_ERR_ARG;
    const int64_t R = n_global + 1;
    // Every global row is owned by ONE host thread (rows split into contiguous ranges): it initialises the rows of its
    // range, scans all label maps (4 B per component) and accumulates the slab rows that map into the range, in table
    // order then row order - the same order as a sequential merge, and the sums are integers anyway.  A whole brain on 8
    // ranks is 2.5 M rows of 104 B on every rank: a single thread spent ~0.3 s per step here, half of it in first-touch
    // page faults of the output arrays.
    const int nthr = R > 200000 ? static_cast<int>(std::min<unsigned>(4u, std::max(1u, std::thread::hardware_concurrency()))) : 1;
    std::vector<int> bad(static_cast<size_t>(nthr), 0);
    auto work = [&](int k) {
        const int64_t g0 = R * k / nthr, g1 = R * (k + 1) / nthr;
        for (int64_t g = g0; g < g1; ++g) {
            counts_out[g] = 0;
            sums_out[3 * g] = sums_out[3 * g + 1] = sums_out[3 * g + 2] = 0;
            int64_t* b = bbox_out + 6 * g;
            b[0] = shape[0]; b[1] = -1; b[2] = shape[1]; b[3] = -1; b[4] = shape[2]; b[5] = -1;
        }
        for (int t = 0; t < ntables; ++t) {
            if (!luts[t] || !counts[t] || !sums[t] || !bbox[t]) continue;          /* a rank without planes */
            const uint64_t z0 = static_cast<uint64_t>(z_offsets[t]);
            const uint32_t* lut = luts[t];
            for (int64_t l = 0; l < rows[t]; ++l) {
                const int64_t g = lut[l];
                if (g >= R) { bad[k] = 1; continue; }
                if (g < g0 || g >= g1) continue;
                const uint64_t c = counts[t][l];
                counts_out[g] += c;
                sums_out[3 * g] += sums[t][3 * l] + c * z0;
                sums_out[3 * g + 1] += sums[t][3 * l + 1];
                sums_out[3 * g + 2] += sums[t][3 * l + 2];
                const int64_t* b = bbox[t] + 6 * l;
                if (b[1] < 0) continue;                                              /* row without voxels: neutral box */
                int64_t* o = bbox_out + 6 * g;
                o[0] = std::min(o[0], b[0] + z_offsets[t]); o[1] = std::max(o[1], b[1] + z_offsets[t]);
                o[2] = std::min(o[2], b[2]); o[3] = std::max(o[3], b[3]);
                o[4] = std::min(o[4], b[4]); o[5] = std::max(o[5], b[5]);
            }
        }
        for (int64_t g = g0; g < g1; ++g) {
            const double c = static_cast<double>(counts_out[g]);                    /* 0 / 0 -> NaN like numpy */
            for (int i = 0; i < 3; ++i) centroids_out[3 * g + i] = static_cast<double>(sums_out[3 * g + i]) / c;
        }
    };
    if (nthr == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int k = 0; k < nthr; ++k) th.emplace_back(work, k);
        for (auto& x : th) x.join();
    }
    for (int k = 0; k < nthr; ++k)
        if (bad[k]) return DLV_ERR_ARG;
    return DLV_OK;
}

/* ---- CSV text of the per-cell table (see include/delivr_b200.h) */
namespace {
// Python's repr(float) (float_repr_style 'short'): shortest digits that round-trip, fixed notation with at least one
// decimal unless the decimal point falls at or before 1e-4 or beyond 16 digits, then d[.ddd]e+XX / e-XX.
char* py_float_repr(double x, char* out) {
    if (x != x) { memcpy(out, "nan", 3); return out + 3; }
    if (x == 0.0) { if (std::signbit(x)) *out++ = '-'; memcpy(out, "0.0", 3); return out + 3; }
    if (std::isinf(x)) { if (x < 0) *out++ = '-'; memcpy(out, "inf", 3); return out + 3; }
    char sci[40];
    const auto r = std::to_chars(sci, sci + sizeof(sci), x, std::chars_format::scientific);   // shortest: d[.ddd]e[+-]XX
    const char* p = sci;
    if (*p == '-') { *out++ = '-'; ++p; }
    char digits[24];
    int nd = 0;
    for (; p < r.ptr && *p != 'e'; ++p)
        if (*p != '.') digits[nd++] = *p;
    int e10 = 0;
    if (p < r.ptr) {
        ++p;
        const bool neg = (*p == '-');
        if (*p == '+' || *p == '-') ++p;
        for (; p < r.ptr; ++p) e10 = e10 * 10 + (*p - '0');
        if (neg) e10 = -e10;
    }
    const int decpt = e10 + 1;                       // value = 0.d1d2... x 10^decpt
    if (decpt <= -4 || decpt > 16) {
        *out++ = digits[0];
        if (nd > 1) { *out++ = '.'; memcpy(out, digits + 1, nd - 1); out += nd - 1; }
        *out++ = 'e';
        int e = decpt - 1;
        *out++ = e < 0 ? '-' : '+';
        if (e < 0) e = -e;
        if (e < 10) *out++ = '0';
        const auto re = std::to_chars(out, out + 8, e);
        return re.ptr;
    }
    if (decpt <= 0) {
        *out++ = '0'; *out++ = '.';
        for (int i = 0; i < -decpt; ++i) *out++ = '0';
        memcpy(out, digits, nd);
        return out + nd;
    }
    if (decpt >= nd) {
        memcpy(out, digits, nd); out += nd;
        for (int i = 0; i < decpt - nd; ++i) *out++ = '0';
        *out++ = '.'; *out++ = '0';
        return out;
    }
    memcpy(out, digits, decpt); out += decpt;
    *out++ = '.';
    memcpy(out, digits + decpt, nd - decpt);
    return out + (nd - decpt);
}
}  // namespace

int64_t dlv_table_csv(const double* centroids, const uint64_t* voxel_counts, int64_t n, char* buf, int64_t cap) {
    if (!centroids || !voxel_counts || n < 0 || cap < 0 || (cap > 0 && !buf)) return DLV_ERR_ARG;
    static const char header[] = ",Blob,Coords,Size\n";
    const int64_t nrows = n > 1 ? n - 1 : 0;         // labels 1 .. n-1
    const int nthr = nrows > 50000 ? static_cast<int>(std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()))) : 1;
    std::vector<std::string> part(static_cast<size_t>(nthr));
    auto work = [&](int k) {
        const int64_t i0 = 1 + nrows * k / nthr, i1 = 1 + nrows * (k + 1) / nthr;
        std::string& s = part[k];
        s.reserve(static_cast<size_t>(i1 - i0) * 72);
        char line[192];
        for (int64_t i = i0; i < i1; ++i) {
            char* o = line;
            *o++ = '0'; *o++ = ',';
            o = std::to_chars(o, o + 20, static_cast<long long>(i)).ptr;
            *o++ = ','; *o++ = '"'; *o++ = '[';
            for (int a = 0; a < 3; ++a) {
                if (a) { *o++ = ','; *o++ = ' '; }
                o = py_float_repr(centroids[3 * i + a], o);
            }
            *o++ = ']'; *o++ = '"'; *o++ = ',';
            o = std::to_chars(o, o + 24, static_cast<unsigned long long>(voxel_counts[i])).ptr;
            *o++ = '\n';
            s.append(line, static_cast<size_t>(o - line));
        }
    };
    if (nthr == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int k = 0; k < nthr; ++k) th.emplace_back(work, k);
        for (auto& x : th) x.join();
    }
    int64_t total = static_cast<int64_t>(sizeof(header) - 1);
    for (const auto& s : part) total += static_cast<int64_t>(s.size());
    if (total <= cap) {
        char* o = buf;
        memcpy(o, header, sizeof(header) - 1); o += sizeof(header) - 1;
        for (const auto& s : part) { memcpy(o, s.data(), s.size()); o += s.size(); }
    }
    return total;
}

/* Host-only: global component numbering from per-slab counts and seam pairs (see include/delivr_b200.h). */
int dlv_resolve_labels(int nslabs, const int64_t* counts, const uint32_t* const* pairs, const int64_t* npairs,
                       uint32_t* const* luts_out, int64_t* n_global_out) {
    if (nslabs < 0 || !counts || !pairs || !npairs || !luts_out || !n_global_out) return DLV_ERR_ARG;
    std::vector<int64_t> off(static_cast<size_t>(nslabs) + 1, 0);
    for (int r = 0; r < nslabs; ++r) {
        if (counts[r] < 0) return DLV_ERR_ARG;
        off[r + 1] = off[r] + counts[r];
    }
    const int64_t total = off[nslabs];
    if (total >= 0xFFFFFFFFll) return DLV_ERR_UNSUPPORTED;
    // node id of (slab r, local label l >= 1) = off[r] + l; the root of a set is its smallest id = the component's
    // member in the lowest slab with the lowest local label = its first voxel in raster order
    std::vector<uint32_t> parent(static_cast<size_t>(total) + 1);
    for (int64_t i = 0; i <= total; ++i) parent[i] = static_cast<uint32_t>(i);
    auto find = [&](uint32_t a) {
        while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; }
        return a;
    };
    for (int r = 1; r < nslabs; ++r) {
        if (!pairs[r] || npairs[r] <= 0) continue;
        for (int64_t k = 0; k < npairs[r]; ++k) {
            const int64_t lo = pairs[r][2 * k], hi = pairs[r][2 * k + 1];
            if (lo == 0 || hi == 0) continue;                                  // background: no adjacency
            if (lo > counts[r - 1] || hi > counts[r]) return DLV_ERR_ARG;
            const uint32_t a = find(static_cast<uint32_t>(off[r - 1] + lo)), b = find(static_cast<uint32_t>(off[r] + hi));
            if (a < b) parent[b] = a; else if (b < a) parent[a] = b;
        }
    }
    std::vector<uint32_t> label(static_cast<size_t>(total) + 1, 0u);
    uint32_t n = 0;
    for (int64_t i = 1; i <= total; ++i) {
        const uint32_t root = find(static_cast<uint32_t>(i));
        label[i] = (root == i) ? ++n : label[root];          // root < i: already numbered
    }
    for (int r = 0; r < nslabs; ++r) {
        if (!luts_out[r]) return DLV_ERR_ARG;
        luts_out[r][0] = 0u;
        for (int64_t l = 1; l <= counts[r]; ++l) luts_out[r][l] = label[off[r] + l];
    }
    *n_global_out = n;
    return DLV_OK;
}

}  // extern "C"
