// dlv_tiff.cu - raw TIFF planes -> device-resident, masked, window-padded uint16 volume.
//
// Replaces the masked_nifti.npy producer loop of the reference (downsample/downsample_and_mask.py:398-414:
// cv2.imread(plane, -1) -> threshold / mask multiply -> write into the zero-padded (1,1,Zp,Yp,Xp) .npy) and
// get_real_size (downsample_and_mask.py:25-30).  The reference writes that array to disk (2 B/voxel) and
// run_inference reads it back (inference/inference.py:232); here every plane is decoded on host threads into pinned
// memory and a kernel reads it straight over PCIe (zero copy), applies the mask rule and writes the padded slab row.
//
// The TIFF reader is a from-scratch baseline decoder (classic TIFF, little/big endian, strips, 8/16-bit unsigned
// grayscale, compression none / LZW / Deflate / PackBits, horizontal predictor) - the subset cv2 / tifffile / Fiji
// write for light-sheet planes.  Anything else is an error, never a silent wrong read.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "dlv_common.cuh"
#include "dlv_internal.h"

namespace dlv {

struct TiffInfo {
    int64_t width = 0, height = 0;
    int bits = 0, compression = 1, predictor = 1, samples = 1, planar = 1, sample_format = 1, photometric = 1;
    int64_t rows_per_strip = -1;
    bool big_endian = false, tiled = false;
    std::vector<uint64_t> offsets, counts;
};

struct Reader {
    const uint8_t* d;
    size_t n;
    bool be;
    bool ok(size_t off, size_t len) const { return off <= n && len <= n - off; }
    uint16_t u16(size_t o) const { return be ? static_cast<uint16_t>((d[o] << 8) | d[o + 1]) : static_cast<uint16_t>(d[o] | (d[o + 1] << 8)); }
    uint32_t u32(size_t o) const {
        return be ? (static_cast<uint32_t>(d[o]) << 24) | (d[o + 1] << 16) | (d[o + 2] << 8) | d[o + 3]
                  : (static_cast<uint32_t>(d[o + 3]) << 24) | (d[o + 2] << 16) | (d[o + 1] << 8) | d[o];
    }
};

static bool read_file(const char* path, std::vector<uint8_t>& buf, std::string& err) {
    FILE* f = fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz < 8) { fclose(f); err = std::string(path) + ": not a TIFF file (too short)"; return false; }
    buf.resize(static_cast<size_t>(sz));
    const size_t got = fread(buf.data(), 1, buf.size(), f);
    fclose(f);
    if (got != buf.size()) { err = std::string(path) + ": short read"; return false; }
    return true;
}

// values of one IFD entry (types BYTE 1, SHORT 3, LONG 4) as integers
static bool entry_values(const Reader& r, size_t e, std::vector<uint64_t>& out) {
    const uint16_t type = r.u16(e + 2);
    const uint32_t count = r.u32(e + 4);
    const size_t esz = type == 1 ? 1 : type == 3 ? 2 : type == 4 ? 4 : 0;
    if (!esz || count == 0) return false;
    size_t off = e + 8;
    if (static_cast<uint64_t>(esz) * count > 4) off = r.u32(e + 8);
    if (!r.ok(off, esz * static_cast<size_t>(count))) return false;
    out.resize(count);
    for (uint32_t i = 0; i < count; ++i)
        out[i] = esz == 1 ? r.d[off + i] : esz == 2 ? r.u16(off + 2 * i) : r.u32(off + 4 * static_cast<size_t>(i));
    return true;
}

static bool parse_tiff(const std::vector<uint8_t>& buf, TiffInfo& t, std::string& err) {
    Reader r{buf.data(), buf.size(), false};
    if (buf[0] == 'I' && buf[1] == 'I') r.be = false;
    else if (buf[0] == 'M' && buf[1] == 'M') r.be = true;
    else { err = "not a TIFF file (byte-order mark)"; return false; }
    t.big_endian = r.be;
    const uint16_t magic = r.u16(2);
    if (magic == 43) { err = "BigTIFF is not supported"; return false; }
    if (magic != 42) { err = "not a TIFF file (magic)"; return false; }
    const size_t ifd = r.u32(4);
    if (!r.ok(ifd, 2)) { err = "IFD outside the file"; return false; }
    const int n = r.u16(ifd);
    if (!r.ok(ifd + 2, static_cast<size_t>(n) * 12)) { err = "IFD outside the file"; return false; }
    std::vector<uint64_t> v;
    for (int i = 0; i < n; ++i) {
        const size_t e = ifd + 2 + static_cast<size_t>(i) * 12;
        const uint16_t tag = r.u16(e);
        switch (tag) {
            case 256: if (entry_values(r, e, v)) t.width = static_cast<int64_t>(v[0]); break;
            case 257: if (entry_values(r, e, v)) t.height = static_cast<int64_t>(v[0]); break;
            case 258: if (entry_values(r, e, v)) t.bits = static_cast<int>(v[0]); break;
            case 259: if (entry_values(r, e, v)) t.compression = static_cast<int>(v[0]); break;
            case 262: if (entry_values(r, e, v)) t.photometric = static_cast<int>(v[0]); break;
            case 273: if (!entry_values(r, e, t.offsets)) { err = "bad StripOffsets"; return false; } break;
            case 277: if (entry_values(r, e, v)) t.samples = static_cast<int>(v[0]); break;
            case 278: if (entry_values(r, e, v)) t.rows_per_strip = static_cast<int64_t>(v[0]); break;
            case 279: if (!entry_values(r, e, t.counts)) { err = "bad StripByteCounts"; return false; } break;
            case 284: if (entry_values(r, e, v)) t.planar = static_cast<int>(v[0]); break;
            case 317: if (entry_values(r, e, v)) t.predictor = static_cast<int>(v[0]); break;
            case 322: case 323: case 324: case 325: t.tiled = true; break;
            case 339: if (entry_values(r, e, v)) t.sample_format = static_cast<int>(v[0]); break;
            default: break;
        }
    }
    if (t.width <= 0 || t.height <= 0) { err = "missing ImageWidth / ImageLength"; return false; }
    if (t.tiled) { err = "tiled TIFF is not supported"; return false; }
    if (t.samples != 1) { err = "only single-channel planes are supported"; return false; }
    if (t.bits != 8 && t.bits != 16) { err = "only 8- and 16-bit samples are supported"; return false; }
    if (t.sample_format != 1) { err = "only unsigned integer samples are supported"; return false; }
    if (t.predictor != 1 && t.predictor != 2) { err = "unsupported predictor"; return false; }
    if (t.compression != 1 && t.compression != 5 && t.compression != 8 && t.compression != 32946 && t.compression != 32773) {
        err = "unsupported compression " + std::to_string(t.compression); return false;
    }
    if (t.rows_per_strip <= 0 || t.rows_per_strip > t.height) t.rows_per_strip = t.height;
    const int64_t nstrips = (t.height + t.rows_per_strip - 1) / t.rows_per_strip;
    if (static_cast<int64_t>(t.offsets.size()) != nstrips || static_cast<int64_t>(t.counts.size()) != nstrips) {
        err = "strip table does not match the image height"; return false;
    }
    for (int64_t s = 0; s < nstrips; ++s)
        if (!r.ok(t.offsets[s], t.counts[s])) { err = "strip outside the file"; return false; }
    return true;
}

// TIFF-flavoured LZW (MSB-first codes, 9..12 bits, "early change"); -> bytes written
static size_t lzw_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    static const int kClear = 256, kEoi = 257;
    std::vector<uint16_t> prefix(4096);
    std::vector<uint8_t> suffix(4096), first(4096);
    std::vector<uint32_t> length(4096);
    for (int i = 0; i < 256; ++i) { prefix[i] = 0; suffix[i] = static_cast<uint8_t>(i); first[i] = static_cast<uint8_t>(i); length[i] = 1; }
    int nbits = 9, next = 258, prev = -1;
    uint64_t acc = 0;
    int have = 0;
    size_t ip = 0, op = 0;
    while (true) {
        while (have < nbits && ip < n) { acc = (acc << 8) | src[ip++]; have += 8; }
        if (have < nbits) break;
        const int code = static_cast<int>((acc >> (have - nbits)) & ((1u << nbits) - 1u));
        have -= nbits;
        if (code == kEoi) break;
        if (code == kClear) { nbits = 9; next = 258; prev = -1; continue; }
        uint32_t len;
        if (prev < 0) {
            if (code >= 256) break;                        // corrupt stream
            if (op < cap) dst[op] = static_cast<uint8_t>(code);
            ++op; prev = code;
            continue;
        }
        if (code < next) {
            len = length[code];
            {
                int c = code;
                for (uint32_t k = len; k-- > 0;) { if (op + k < cap) dst[op + k] = suffix[c]; c = prefix[c]; }
            }
            if (next < 4096) {
                prefix[next] = static_cast<uint16_t>(prev); suffix[next] = first[code]; first[next] = first[prev];
                length[next] = length[prev] + 1; ++next;
            }
        } else if (code == next && next < 4096) {
            prefix[next] = static_cast<uint16_t>(prev); suffix[next] = first[prev]; first[next] = first[prev];
            length[next] = length[prev] + 1; ++next;
            len = length[code];
            {
                int c = code;
                for (uint32_t k = len; k-- > 0;) { if (op + k < cap) dst[op + k] = suffix[c]; c = prefix[c]; }
            }
        } else {
            break;                                          // corrupt stream
        }
        op += len;
        if (op >= cap) break;
        if (next >= (1 << nbits) - 1 && nbits < 12) ++nbits;
        prev = code;
    }
    return std::min(op, cap);
}

static size_t packbits_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    size_t ip = 0, op = 0;
    while (ip < n && op < cap) {
        const int8_t h = static_cast<int8_t>(src[ip++]);
        if (h >= 0) {
            const size_t c = std::min<size_t>(static_cast<size_t>(h) + 1, std::min(n - ip, cap - op));
            memcpy(dst + op, src + ip, c); ip += static_cast<size_t>(h) + 1; op += c;
        } else if (h != -128) {
            if (ip >= n) break;
            const size_t c = std::min<size_t>(static_cast<size_t>(1 - h), cap - op);
            memset(dst + op, src[ip++], c); op += c;
        }
    }
    return op;
}

// Decodes IFD 0 of `path` into out[height][width] uint16 (8-bit samples are widened, like .astype(np.uint16)).
static bool tiff_read_u16(const char* path, uint16_t* out, int64_t height, int64_t width, std::string& err) {
    std::vector<uint8_t> buf;
    if (!read_file(path, buf, err)) return false;
    TiffInfo t;
    if (!parse_tiff(buf, t, err)) { err = std::string(path) + ": " + err; return false; }
    if (t.height != height || t.width != width) {
        err = std::string(path) + ": plane is " + std::to_string(t.height) + "x" + std::to_string(t.width) + ", expected " +
              std::to_string(height) + "x" + std::to_string(width);
        return false;
    }
    const size_t bps = t.bits / 8, rowbytes = static_cast<size_t>(width) * bps;
    std::vector<uint8_t> strip;
    for (size_t s = 0; s < t.offsets.size(); ++s) {
        const int64_t r0 = static_cast<int64_t>(s) * t.rows_per_strip;
        const int64_t nr = std::min<int64_t>(t.rows_per_strip, height - r0);
        const size_t want = static_cast<size_t>(nr) * rowbytes;
        const uint8_t* src = buf.data() + t.offsets[s];
        const size_t sn = t.counts[s];
        const uint8_t* raw = nullptr;
        if (t.compression == 1) {
            if (sn < want) { err = std::string(path) + ": truncated strip"; return false; }
            raw = src;
        } else {
            strip.resize(want);
            size_t got = 0;
            if (t.compression == 5) got = lzw_decode(src, sn, strip.data(), want);
            else if (t.compression == 32773) got = packbits_decode(src, sn, strip.data(), want);
            else {
                uLongf dl = static_cast<uLongf>(want);
                const int zr = uncompress(strip.data(), &dl, src, static_cast<uLong>(sn));
                got = (zr == Z_OK || zr == Z_BUF_ERROR) ? static_cast<size_t>(dl) : 0;
            }
            if (got != want) { err = std::string(path) + ": strip " + std::to_string(s) + " decoded to " + std::to_string(got) + " of " + std::to_string(want) + " bytes"; return false; }
            raw = strip.data();
        }
        for (int64_t y = 0; y < nr; ++y) {
            uint16_t* o = out + (r0 + y) * width;
            const uint8_t* p = raw + static_cast<size_t>(y) * rowbytes;
            if (bps == 2) {
                if (t.big_endian) for (int64_t x = 0; x < width; ++x) o[x] = static_cast<uint16_t>((p[2 * x] << 8) | p[2 * x + 1]);
                else memcpy(o, p, rowbytes);
                if (t.predictor == 2) for (int64_t x = 1; x < width; ++x) o[x] = static_cast<uint16_t>(o[x] + o[x - 1]);
            } else {
                uint8_t run = 0;
                for (int64_t x = 0; x < width; ++x) {
                    const uint8_t v = (t.predictor == 2 && x > 0) ? static_cast<uint8_t>(p[x] + run) : p[x];
                    run = v; o[x] = v;
                }
            }
        }
    }
    return true;
}

// ------------------------------------------------------------------- device side
// One decoded plane (pinned host memory, read over PCIe) -> one padded slab plane.  Mask rule of
// downsample_and_mask.py:405-411: `img *= mask[z]` (uint16 wrap-around multiply) or `img[img < threshold] = 0`.
__global__ void ingest_plane_kernel(const uint16_t* __restrict__ plane, int64_t Y, int64_t X, int threshold,
                                    const uint8_t* __restrict__ mask, uint16_t* __restrict__ out, int64_t SY, int64_t SX) {
    const int64_t x0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
    const int64_t y = blockIdx.y;
    if (x0 >= SX) return;
    uint16_t v[8];
    const bool row = y < Y;
    const bool full = row && x0 + 8 <= X && (X % 8 == 0);
    if (full) {
        const uint4 u = *reinterpret_cast<const uint4*>(plane + y * X + x0);
        memcpy(v, &u, 16);
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (row && x0 + j < X) ? plane[y * X + x0 + j] : static_cast<uint16_t>(0);
    }
    if (row) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (mask) { if (x0 + j < X) v[j] = static_cast<uint16_t>(v[j] * mask[y * X + x0 + j]); }
            else if (threshold >= 0 && static_cast<int>(v[j]) < threshold) v[j] = 0;
        }
    }
    uint16_t* o = out + y * SX + x0;
    if (x0 + 8 <= SX && (SX % 8 == 0)) {
        uint4 u;
        memcpy(&u, v, 16);
        *reinterpret_cast<uint4*>(o) = u;
    } else {
        for (int j = 0; j < 8 && x0 + j < SX; ++j) o[j] = v[j];
    }
}

int tiff_load_planes(Ctx* ctx, const char* const* paths, int n, int64_t Y, int64_t X, int threshold, const uint8_t* mask_dev,
                     uint16_t* slab_dev, int64_t SY, int64_t SX, int nthreads) {
    if (n <= 0) return 0;
    if (SY < Y || SX < X) { set_error(ctx, "dlv_load_tiff_planes: slab plane %lldx%lld smaller than the image %lldx%lld", (long long)SY, (long long)SX, (long long)Y, (long long)X); return DLV_ERR_ARG; }
    nthreads = std::max(1, std::min(nthreads > 0 ? nthreads : static_cast<int>(std::thread::hardware_concurrency()), 64));
    nthreads = std::min(nthreads, n);
    const size_t plane_bytes = static_cast<size_t>(Y) * X * 2;
    uint16_t* pinned[2] = {nullptr, nullptr};
    cudaEvent_t freed[2] = {nullptr, nullptr};
    int rc = 0;
    std::string err;
    for (int h = 0; h < 2; ++h) {
        if (cudaHostAlloc(reinterpret_cast<void**>(&pinned[h]), plane_bytes * nthreads, cudaHostAllocDefault) != cudaSuccess ||
            cudaEventCreateWithFlags(&freed[h], cudaEventDisableTiming) != cudaSuccess) {
            set_error(ctx, "dlv_load_tiff_planes: pinned staging allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = DLV_ERR_CUDA;
        }
    }
    const dim3 grid(static_cast<unsigned>((SX + 8 * 128 - 1) / (8 * 128)), static_cast<unsigned>(SY));
    for (int g0 = 0, gi = 0; g0 < n && rc == 0; g0 += nthreads, ++gi) {
        const int h = gi & 1, cnt = std::min(nthreads, n - g0);
        if (gi >= 2) cudaEventSynchronize(freed[h]);            // the kernels that read this half have finished
        std::vector<std::string> errs(cnt);
        std::vector<std::thread> pool;
        std::atomic<int> bad{0};
        for (int k = 0; k < cnt; ++k)
            pool.emplace_back([&, k]() {
                if (!tiff_read_u16(paths[g0 + k], pinned[h] + static_cast<size_t>(k) * Y * X, Y, X, errs[k])) bad++;
            });
        for (auto& t : pool) t.join();
        if (bad) {
            for (const auto& e : errs) if (!e.empty()) { err = e; break; }
            set_error(ctx, "dlv_load_tiff_planes: %s", err.c_str());
            rc = DLV_ERR_ARG;
            break;
        }
        for (int k = 0; k < cnt; ++k) {
            const int64_t z = g0 + k;
            ingest_plane_kernel<<<grid, 128, 0, ctx->stream>>>(pinned[h] + static_cast<size_t>(k) * Y * X, Y, X, threshold,
                                                              mask_dev ? mask_dev + static_cast<size_t>(z) * Y * X : nullptr,
                                                              slab_dev + static_cast<size_t>(z) * SY * SX, SY, SX);
            ctx->launches++;
        }
        if (cudaGetLastError() != cudaSuccess) { set_error(ctx, "dlv_load_tiff_planes: kernel launch failed"); rc = DLV_ERR_CUDA; break; }
        cudaEventRecord(freed[h], ctx->stream);
    }
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (rc == 0 && e != cudaSuccess) { set_error(ctx, "dlv_load_tiff_planes: %s", cudaGetErrorString(e)); rc = DLV_ERR_CUDA; }
    for (int h = 0; h < 2; ++h) { if (pinned[h]) cudaFreeHost(pinned[h]); if (freed[h]) cudaEventDestroy(freed[h]); }
    return rc;
}

}  // namespace dlv

// ------------------------------------------------------------------- C ABI
static thread_local std::string g_tiff_error;

extern "C" {

const char* dlv_tiff_last_error(void) { return g_tiff_error.c_str(); }

int dlv_tiff_info(const char* path, int64_t* height, int64_t* width, int32_t* bits, int32_t* compression) {
    if (!path) return DLV_ERR_ARG;
    std::vector<uint8_t> buf;
    dlv::TiffInfo t;
    std::string err;
    if (!dlv::read_file(path, buf, err) || !dlv::parse_tiff(buf, t, err)) { g_tiff_error = err; return DLV_ERR_ARG; }
    if (height) *height = t.height;
    if (width) *width = t.width;
    if (bits) *bits = t.bits;
    if (compression) *compression = t.compression;
    return DLV_OK;
}

int dlv_tiff_read_u16(const char* path, uint16_t* out_host, int64_t height, int64_t width) {
    if (!path || !out_host) return DLV_ERR_ARG;
    std::string err;
    if (!dlv::tiff_read_u16(path, out_host, height, width, err)) { g_tiff_error = err; return DLV_ERR_ARG; }
    return DLV_OK;
}

int dlv_load_tiff_planes(dlv_ctx* c, const char* const* paths, int n, int64_t Y, int64_t X, int32_t threshold,
                         const uint8_t* mask_dev_or_null, uint16_t* slab_dev, int64_t SY, int64_t SX, int nthreads) {
    dlv::Ctx* ctx = reinterpret_cast<dlv::Ctx*>(c);
    if (!ctx) return DLV_ERR_ARG;
    if (!paths || !slab_dev || n < 0 || Y <= 0 || X <= 0) { dlv::set_error(ctx, "dlv_load_tiff_planes: bad argument"); return DLV_ERR_ARG; }
    cudaSetDevice(ctx->device);
    return dlv::tiff_load_planes(ctx, paths, n, Y, X, threshold, mask_dev_or_null, slab_dev, SY, SX, nthreads);
}

}  // extern "C"
