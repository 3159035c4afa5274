// dlv_tiff_write.cu - host-only TIFF plane writer for the painter's outputs.
//
// Replaces the per-plane `tifffile.imwrite(path, plane, compression='lzw')` loops of the reference
// (blob_highlighter.py:127-133, :158-161; blob_depthmap.py:209-213): one file per z plane, 8- or 16-bit
// unsigned grayscale.  The reference writes the planes one after the other on one core; for a whole brain that is
// 3 x 1500 planes of 16 Mpixel.  Here the planes are LZW-compressed on all host threads.
//
// Format: classic little-endian TIFF, one IFD, strips of ~256 KB, compression none (1) / LZW (5, TIFF flavour:
// MSB-first codes of 9..12 bits, "early change", ClearCode first, table reset when full - what libtiff writes) /
// Deflate (8, zlib), no predictor.  Pixels - not bytes - are the contract: any baseline reader (libtiff via OpenCV,
// tifffile, Fiji, this library's own reader) returns the array that was written.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "dlv_internal.h"

namespace dlv {

// ---- TIFF LZW encoder (mirrors libtiff's code-width schedule: the width grows when the next free entry exceeds
// 2^bits - 1, the table is cleared at 4094 entries, one more entry is accounted for before EndOfInformation)
class LzwEncoder {
public:
    explicit LzwEncoder(std::vector<uint8_t>& out) : out_(out), hash_(kHashSize) {}

    void encode(const uint8_t* src, size_t n) {
        reset_table();
        put(kClear);
        if (n == 0) { put(kEoi); flush(); return; }
        int ent = src[0];
        for (size_t i = 1; i < n; ++i) {
            const int c = src[i];
            const uint32_t key = (static_cast<uint32_t>(ent) << 8) | static_cast<uint32_t>(c);
            uint32_t h = (key * 2654435761u) >> (32 - kHashBits);
            bool found = false;
            while (hash_[h].code >= 0) {
                if (hash_[h].key == key) { ent = hash_[h].code; found = true; break; }
                h = (h + 1) & (kHashSize - 1);
            }
            if (found) continue;
            put(ent);
            ent = c;
            hash_[h].key = key;
            hash_[h].code = free_++;
            if (free_ == kCodeMax - 1) {              // table full: clear and start over
                put(kClear);
                reset_table();
            } else if (free_ > maxcode_) {
                ++nbits_;
                maxcode_ = (1 << nbits_) - 1;
            }
        }
        put(ent);
        ++free_;                                       // the entry the decoder adds after the last code (libtiff's LZWPostEncode)
        if (free_ == kCodeMax - 1) {
            put(kClear);
            reset_table();
        } else if (free_ > maxcode_ && nbits_ < 12) {
            ++nbits_;
            maxcode_ = (1 << nbits_) - 1;
        }
        put(kEoi);
        flush();
    }

private:
    static const int kClear = 256, kEoi = 257, kFirst = 258, kCodeMax = 4095;
    static const int kHashBits = 13, kHashSize = 1 << kHashBits;
    struct Slot { uint32_t key; int code; };

    void reset_table() {
        for (auto& s : hash_) s.code = -1;
        free_ = kFirst;
        nbits_ = 9;
        maxcode_ = 511;
    }
    void put(int code) {
        acc_ = (acc_ << nbits_) | static_cast<uint64_t>(code);
        have_ += nbits_;
        while (have_ >= 8) { out_.push_back(static_cast<uint8_t>(acc_ >> (have_ - 8))); have_ -= 8; }
    }
    void flush() {
        if (have_ > 0) { out_.push_back(static_cast<uint8_t>(acc_ << (8 - have_))); have_ = 0; }
        acc_ = 0;
    }

    std::vector<uint8_t>& out_;
    std::vector<Slot> hash_;
    int free_ = kFirst, nbits_ = 9, maxcode_ = 511;
    uint64_t acc_ = 0;
    int have_ = 0;
};

static void put16(std::vector<uint8_t>& b, uint16_t v) { b.push_back(v & 0xFF); b.push_back(v >> 8); }
static void put32(std::vector<uint8_t>& b, uint32_t v) { for (int k = 0; k < 4; ++k) b.push_back((v >> (8 * k)) & 0xFF); }
static void ifd_entry(std::vector<uint8_t>& b, uint16_t tag, uint16_t type, uint32_t count, uint32_t value) {
    put16(b, tag); put16(b, type); put32(b, count);
    if (type == 3 && count == 1) { put16(b, static_cast<uint16_t>(value)); put16(b, 0); } else put32(b, value);
}

bool tiff_write_plane(const char* path, const uint8_t* data, int64_t height, int64_t width, int bits, int compression, std::string& err) {
    if (height <= 0 || width <= 0 || (bits != 8 && bits != 16) || (compression != 1 && compression != 5 && compression != 8)) {
        err = "tiff write: unsupported geometry / sample width / compression";
        return false;
    }
    const int64_t row_bytes = width * (bits / 8);
    const int64_t rps = std::max<int64_t>(1, std::min<int64_t>(height, (256 * 1024) / row_bytes));
    const int64_t nstrips = (height + rps - 1) / rps;
    if (static_cast<uint64_t>(height) * row_bytes > 0xF0000000ull) { err = "tiff write: plane exceeds classic TIFF's 4 GB"; return false; }
    std::vector<uint8_t> file;
    file.reserve(static_cast<size_t>(height * row_bytes / 4 + 4096));
    file.push_back('I'); file.push_back('I'); put16(file, 42); put32(file, 0);          // IFD offset patched below
    std::vector<uint32_t> offs(nstrips), counts(nstrips);
    std::vector<uint8_t> tmp;
    for (int64_t s = 0; s < nstrips; ++s) {
        const int64_t r0 = s * rps, r1 = std::min(height, r0 + rps);
        const uint8_t* src = data + r0 * row_bytes;
        const size_t n = static_cast<size_t>((r1 - r0) * row_bytes);
        offs[s] = static_cast<uint32_t>(file.size());
        if (compression == 1) {
            file.insert(file.end(), src, src + n);
        } else if (compression == 5) {
            LzwEncoder enc(file);
            enc.encode(src, n);
        } else {
            uLongf cap = compressBound(static_cast<uLong>(n));
            tmp.resize(cap);
            if (compress2(tmp.data(), &cap, src, static_cast<uLong>(n), 6) != Z_OK) { err = "tiff write: zlib failed"; return false; }
            file.insert(file.end(), tmp.begin(), tmp.begin() + cap);
        }
        counts[s] = static_cast<uint32_t>(file.size() - offs[s]);
        if (file.size() > 0xF0000000ull) { err = "tiff write: file exceeds classic TIFF's 4 GB"; return false; }
    }
    if (file.size() & 1) file.push_back(0);
    uint32_t off_arr = 0, cnt_arr = 0;
    if (nstrips > 1) {
        off_arr = static_cast<uint32_t>(file.size());
        for (auto v : offs) put32(file, v);
        cnt_arr = static_cast<uint32_t>(file.size());
        for (auto v : counts) put32(file, v);
    }
    const uint32_t ifd = static_cast<uint32_t>(file.size());
    for (int k = 0; k < 4; ++k) file[4 + k] = (ifd >> (8 * k)) & 0xFF;
    put16(file, 10);                                                                    // entries, ascending tags
    ifd_entry(file, 256, 4, 1, static_cast<uint32_t>(width));
    ifd_entry(file, 257, 4, 1, static_cast<uint32_t>(height));
    ifd_entry(file, 258, 3, 1, static_cast<uint32_t>(bits));
    ifd_entry(file, 259, 3, 1, static_cast<uint32_t>(compression));
    ifd_entry(file, 262, 3, 1, 1);                                                      // BlackIsZero
    ifd_entry(file, 273, 4, static_cast<uint32_t>(nstrips), nstrips > 1 ? off_arr : offs[0]);
    ifd_entry(file, 277, 3, 1, 1);
    ifd_entry(file, 278, 4, 1, static_cast<uint32_t>(rps));
    ifd_entry(file, 279, 4, static_cast<uint32_t>(nstrips), nstrips > 1 ? cnt_arr : counts[0]);
    ifd_entry(file, 339, 3, 1, 1);                                                      // unsigned integer samples
    put32(file, 0);                                                                     // no further IFD
    FILE* f = fopen(path, "wb");
    if (!f) { err = std::string("tiff write: cannot open ") + path; return false; }
    const bool ok = fwrite(file.data(), 1, file.size(), f) == file.size();
    if (fclose(f) != 0 || !ok) { err = std::string("tiff write: short write to ") + path; return false; }
    return true;
}

}  // namespace dlv

static thread_local std::string g_tiffw_error;

extern "C" {

const char* dlv_tiff_write_last_error(void) { return g_tiffw_error.c_str(); }

int dlv_tiff_write_planes(const char* const* paths, int n, const void* volume_host, int64_t height, int64_t width, int32_t bits,
                          int32_t compression, int nthreads) {
    if (!paths || !volume_host || n < 0) { g_tiffw_error = "dlv_tiff_write_planes: null argument"; return DLV_ERR_ARG; }
    if (n == 0) return DLV_OK;
    const int64_t plane_bytes = height * width * (bits / 8);
    int nt = nthreads > 0 ? nthreads : static_cast<int>(std::max(1u, std::thread::hardware_concurrency()));
    nt = std::min(nt, n);
    std::atomic<int> next(0), failed(0);
    std::string first_err;
    std::vector<std::string> errs(nt);
    auto work = [&](int t) {
        for (int i = next.fetch_add(1); i < n && !failed.load(); i = next.fetch_add(1)) {
            if (!dlv::tiff_write_plane(paths[i], static_cast<const uint8_t*>(volume_host) + static_cast<int64_t>(i) * plane_bytes, height, width,
                                       bits, compression, errs[t]))
                failed.store(1);
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    if (failed.load()) {
        for (auto& e : errs) if (!e.empty()) { g_tiffw_error = e; break; }
        return DLV_ERR_ARG;
    }
    return DLV_OK;
}

}  // extern "C"
