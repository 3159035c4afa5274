// dlv_ccl.cu - 26-connected 3-D connected-component labelling + per-component statistics on the GPU.
//
// Replaces cc3d.connected_components(bin_img, return_N=True) and cc3d.statistics(labels) of the reference
// (count_blobs.py:61,64,85; connected-components-3d 3.12.3).  Output contract: labels 1..N numbered by each
// component's first voxel in C-order raster scan; table rows 0..N with exact integer counts / coordinate sums /
// inclusive bounding boxes.
//
// Algorithm (HBM-bound integer work, no tensor cores):
//   P1 init      read the uint8 mask once, emit a 1-bit/voxel bitmask and the label array (0 for background,
//                1 + linear index of the start of the voxel's x-run inside its 32-voxel word otherwise).
//                This is the only pass that touches every label (4 B/voxel write).
//   P2 merge     one thread per bitmask word: for every x-run, union (lock-free atomicMin union-find, root =
//                smallest linear index = first voxel in raster order) with the runs it touches in the 4 raster-
//                predecessor rows (y-1 | z-1,y-1 | z-1,y | z-1,y+1; x-range widened by 1) and the previous word;
//                unions already implied by the predecessor rows' own unions are skipped (ccl_merge_kernel).
//   P3 compress  every run start resolves its root; roots are flagged in a second bitmask.
//   P4 scan      exclusive prefix sum over popcounts of the root bitmask -> rank of every root = final label.
//   P5 relabel   every run looks up its root's rank, rewrites its voxels, and reduces count / sum z,y,x / bbox
//                once per run (closed-form arithmetic series along x) with 64-bit atomics.
#include <limits.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "dlv_common.cuh"
#include "dlv_internal.h"

namespace dlv {

struct CclGeom {
    int64_t Z, Y, X;
    int64_t rows;      // Z*Y
    int W;             // bitmask words per row
};

__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) { return __ldcg(p); }

// labels are 1-based voxel indices; L[l-1] is the parent of l
__device__ __forceinline__ uint32_t uf_find(uint32_t* L, uint32_t l) {
    uint32_t p;
    while ((p = ld_cg_u32(L + (l - 1))) != l) l = p;
    return l;
}
__device__ __forceinline__ void uf_union(uint32_t* L, uint32_t a, uint32_t b) {
    while (true) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a < b) { uint32_t t = a; a = b; b = t; }
        const uint32_t old = atomicMin(L + (a - 1), b);
        if (old == a) return;
        a = old;
    }
}

struct BgBox { int zmin, zmax, ymin, ymax, xmin, xmax; };

// ---- P1: mask -> bitmask + initial labels (+ bounding box of the background for table row 0)
// One warp handles kInitSegs x 128 consecutive voxels per iteration (4 per lane and segment, 16 B label stores); the
// kInitSegs mask loads are issued before any of them is used, so every thread keeps that many loads in flight
// (one load per thread and iteration left the kernel latency-bound at 2.6 TB/s).
constexpr int kInitSegs = 4;      // 8 measured the same (the pass is instruction-bound: IPC 2.3, issue slots 57 % busy, DRAM 17 %; profiles/r02_r_ccl_init_merge_sol.txt)
// `prezeroed`: L was cleared by a memset (full-rate write); the kernel then stores only the 16 B groups that hold
// foreground - on blob masks (a few % foreground) that is < 10 % of the label array instead of all of it.
__global__ void __launch_bounds__(256, 5) ccl_init_kernel(const uint8_t* __restrict__ mask, CclGeom g, uint32_t* __restrict__ bits,
                                uint32_t* __restrict__ L, int* __restrict__ bgbox, int prezeroed) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int segs = (g.W + 3) / 4;                       // 128-voxel segments per row
    const bool vec = (g.X % 4) == 0 && (reinterpret_cast<uintptr_t>(mask) & 3u) == 0 && (reinterpret_cast<uintptr_t>(L) & 15u) == 0;
    const uint32_t uY = static_cast<uint32_t>(g.Y);
    BgBox bb = {INT_MAX, -1, INT_MAX, -1, INT_MAX, -1};
    // rows are dealt to warps; the (z, y) decode is one 32-bit division per row (64-bit divisions per segment made
    // this kernel instruction-bound)
    for (int64_t r = warp_global; r < g.rows; r += nwarps) {
        const int z = static_cast<int>(static_cast<uint32_t>(r) / uY), y = static_cast<int>(static_cast<uint32_t>(r) - static_cast<uint32_t>(z) * uY);
        const int64_t base = r * g.X;
        for (int sg0 = 0; sg0 < segs; sg0 += kInitSegs) {
            uint32_t nibs[kInitSegs];
#pragma unroll
            for (int u = 0; u < kInitSegs; ++u) {
                nibs[u] = 0;
                const int64_t x0 = static_cast<int64_t>(sg0 + u) * 128 + lane * 4;
                if (sg0 + u >= segs) continue;
                if (vec) {
                    if (x0 < g.X) {
                        const uint32_t m = __ldcs(reinterpret_cast<const uint32_t*>(mask + base + x0));
                        nibs[u] = ((m & 0xFFu) ? 1u : 0u) | ((m & 0xFF00u) ? 2u : 0u) | ((m & 0xFF0000u) ? 4u : 0u) | ((m & 0xFF000000u) ? 8u : 0u);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (x0 + j < g.X && mask[base + x0 + j]) nibs[u] |= 1u << j;
                }
            }
#pragma unroll
            for (int u = 0; u < kInitSegs; ++u) {
                const int sg = sg0 + u;
                if (sg >= segs) break;                         // warp-uniform
                const int64_t x0 = static_cast<int64_t>(sg) * 128 + lane * 4;
                if (vec && __ballot_sync(0xffffffffu, nibs[u] != 0u) == 0u) {
                    // 128 background voxels (about half of the segments of a blob mask): zero labels, zero bitmask words
                    // and the background box - none of the run arithmetic below
                    const int wi = sg * 4 + (lane >> 3);
                    if ((lane & 7) == 0 && wi < g.W) bits[r * g.W + wi] = 0u;
                    if (x0 < g.X) {
                        if (!prezeroed) __stcs(reinterpret_cast<uint4*>(L + base + x0), make_uint4(0u, 0u, 0u, 0u));
                        bb.zmin = min(bb.zmin, z); bb.zmax = max(bb.zmax, z);
                        bb.ymin = min(bb.ymin, y); bb.ymax = max(bb.ymax, y);
                        bb.xmin = min(bb.xmin, static_cast<int>(x0)); bb.xmax = max(bb.xmax, static_cast<int>(x0 + 3 < g.X ? x0 + 3 : g.X - 1));
                    }
                    continue;
                }
                // assemble the 32-bit word of this lane's 8-lane group
                uint32_t word = nibs[u] << (4 * (lane & 7));
                word |= __shfl_xor_sync(0xffffffffu, word, 1);
                word |= __shfl_xor_sync(0xffffffffu, word, 2);
                word |= __shfl_xor_sync(0xffffffffu, word, 4);
                const int wi = sg * 4 + (lane >> 3);
                if ((lane & 7) == 0 && wi < g.W) bits[r * g.W + wi] = word;
                if (x0 < g.X) {
                    const uint32_t wbase = static_cast<uint32_t>(base + static_cast<int64_t>(wi) * 32);   // 0-based index of bit 0
                    // label = 1 + index of the first voxel of the x-run inside this word: walk the lane's 4 bits with a
                    // running run start (the highest zero bit below the lane's first bit starts it)
                    const int b0 = 4 * (lane & 7);
                    const uint32_t zeros_below = ~word & ((1u << b0) - 1u);
                    int run = zeros_below ? (32 - __clz(zeros_below)) : 0;
                    uint32_t lab[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool fg = (word >> (b0 + j)) & 1u;
                        lab[j] = fg ? wbase + run + 1u : 0u;
                        run = fg ? run : b0 + j + 1;
                    }
                    // background voxels of this lane (inside the row): x-range from the nibble's zero bits
                    uint32_t bgn = ~nibs[u] & 0xFu;
                    if (x0 + 4 > g.X) bgn &= (1u << (g.X - x0)) - 1u;
                    const int nbg = bgn != 0;
                    const int bx0 = static_cast<int>(x0) + (__ffs(bgn) - 1), bx1 = static_cast<int>(x0) + (31 - __clz(bgn));
                    if (vec) {
                        if (!prezeroed || nibs[u]) __stcs(reinterpret_cast<uint4*>(L + base + x0), make_uint4(lab[0], lab[1], lab[2], lab[3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (x0 + j < g.X) L[base + x0 + j] = lab[j];
                    }
                    if (nbg) {
                        bb.zmin = min(bb.zmin, z); bb.zmax = max(bb.zmax, z);
                        bb.ymin = min(bb.ymin, y); bb.ymax = max(bb.ymax, y);
                        bb.xmin = min(bb.xmin, bx0); bb.xmax = max(bb.xmax, bx1);
                    }
                }
            }
        }
    }
    // warp-reduce the background box, one set of atomics per warp
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        bb.zmin = min(bb.zmin, __shfl_xor_sync(0xffffffffu, bb.zmin, o)); bb.zmax = max(bb.zmax, __shfl_xor_sync(0xffffffffu, bb.zmax, o));
        bb.ymin = min(bb.ymin, __shfl_xor_sync(0xffffffffu, bb.ymin, o)); bb.ymax = max(bb.ymax, __shfl_xor_sync(0xffffffffu, bb.ymax, o));
        bb.xmin = min(bb.xmin, __shfl_xor_sync(0xffffffffu, bb.xmin, o)); bb.xmax = max(bb.xmax, __shfl_xor_sync(0xffffffffu, bb.xmax, o));
    }
    if (lane == 0 && bb.zmax >= 0) {
        atomicMin(bgbox + 0, bb.zmin); atomicMax(bgbox + 1, bb.zmax);
        atomicMin(bgbox + 2, bb.ymin); atomicMax(bgbox + 3, bb.ymax);
        atomicMin(bgbox + 4, bb.xmin); atomicMax(bgbox + 5, bb.xmax);
    }
}

// ---- P1, 1024 voxels per warp and iteration, for rows that are whole 32-voxel words (X % 32 == 0, 16 B aligned buffers:
// cfg2 / cfg3 / cfg4).  Every global access is a fully coalesced 16 B per lane: two loads bring the warp's 1024 mask
// bytes (lane l: voxels 16 l .. 16 l + 15 of each half), eight stores write its 4 KB of labels (store k, lane l: the
// 4 labels of 16 B chunk 32 k + l); the bitmask words travel between the two layouts by shuffle.  ~250 instructions per
// 1024 voxels where the nibble-per-lane form above spends ~800: that form is instruction-bound (IPC 2.3, DRAM 17 %,
// profiles/r02_r_ccl_init_merge_sol.txt), this one is bound by the 4 B/voxel label write.
__device__ __forceinline__ uint32_t nonzero_nibble(uint32_t m) {          // bit j <- byte j of m is non-zero
    return ((__vcmpne4(m, 0u) & 0x08040201u) * 0x01010101u) >> 24;
}
__global__ void __launch_bounds__(256) ccl_init_words_kernel(const uint8_t* __restrict__ mask, CclGeom g, unsigned long long w_magic,
                                                            unsigned long long y_magic, uint32_t* __restrict__ bits,
                                                            uint32_t* __restrict__ L, int* __restrict__ bgbox) {
    const int lane = threadIdx.x & 31;
    const int64_t nwords = g.rows * g.W;
    const int64_t ngroups = (nwords + 31) / 32;                       // groups of 32 words = 1024 voxels
    const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    BgBox bb = {INT_MAX, -1, INT_MAX, -1, INT_MAX, -1};
    for (int64_t grp = warp0; grp < ngroups; grp += nwarps) {
        const int64_t t0 = grp * 32;                                    // first word of the group
        const int nw = (nwords - t0 < 32) ? static_cast<int>(nwords - t0) : 32; // words in this group (< 32 only in the last one)
        // half h of the group = words 16 h .. 16 h + 15; lane l loads bytes 16 l .. of that half: word 16 h + l / 2, half l % 2
        uint32_t wh[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t part = 0u;
            if (16 * h + (lane >> 1) < nw) {
                const uint4 m = __ldcs(reinterpret_cast<const uint4*>(mask + (t0 + 16 * h) * 32) + lane);
                part = (nonzero_nibble(m.x) | (nonzero_nibble(m.y) << 4) | (nonzero_nibble(m.z) << 8) | (nonzero_nibble(m.w) << 12)) << (16 * (lane & 1));
            }
            wh[h] = part | __shfl_xor_sync(0xffffffffu, part, 1);      // lanes 2 j and 2 j + 1 both hold word 16 h + j
        }
        if ((lane & 1) == 0) {
            if ((lane >> 1) < nw) bits[t0 + (lane >> 1)] = wh[0];
            if (16 + (lane >> 1) < nw) bits[t0 + 16 + (lane >> 1)] = wh[1];
        }
        uint4* dst = reinterpret_cast<uint4*>(L + t0 * 32);
        if (__ballot_sync(0xffffffffu, (wh[0] | wh[1]) != 0u) == 0u) {
            // 1024 background voxels
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k * 4 + (lane >> 3) < nw) __stcs(dst + k * 32 + lane, make_uint4(0u, 0u, 0u, 0u));
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                // chunk 32 k + l = 4 voxels of word wi = 4 k + l / 8, bits 4 (l % 8) ..
                const int wi = k * 4 + (lane >> 3);
                const uint32_t word = __shfl_sync(0xffffffffu, (k < 4) ? wh[0] : wh[1], 2 * (wi & 15));
                if (wi < nw) {
                    const uint32_t wbase = static_cast<uint32_t>((t0 + wi) * 32);    // 0-based voxel index of bit 0 (rows are whole words)
                    const int b0 = 4 * (lane & 7);
                    const uint32_t zeros_below = ~word & ((1u << b0) - 1u);
                    int run = zeros_below ? (32 - __clz(zeros_below)) : 0;           // start bit of the x-run in progress (inside this word)
                    uint32_t lab[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool fg = (word >> (b0 + j)) & 1u;
                        lab[j] = fg ? wbase + run + 1u : 0u;
                        run = fg ? run : b0 + j + 1;
                    }
                    __stcs(dst + k * 32 + lane, make_uint4(lab[0], lab[1], lab[2], lab[3]));
                }
            }
        }
        // background box (table row 0), once per word on the even lanes: (z, y) of the row by exact 64-bit magic division
        if ((lane & 1) == 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int wi = 16 * h + (lane >> 1);
                const uint32_t bgw = ~wh[h];
                if (wi < nw && bgw) {
                    const int64_t t = t0 + wi;
                    const uint32_t r = static_cast<uint32_t>(__umul64hi(static_cast<unsigned long long>(t), w_magic));
                    const int w = static_cast<int>(t - static_cast<int64_t>(r) * g.W);
                    const int z = static_cast<int>(__umul64hi(static_cast<unsigned long long>(r), y_magic));
                    const int y = static_cast<int>(r - static_cast<uint32_t>(z) * static_cast<uint32_t>(g.Y));
                    const int bx0 = w * 32 + (__ffs(bgw) - 1), bx1 = w * 32 + (31 - __clz(bgw));
                    bb.zmin = min(bb.zmin, z); bb.zmax = max(bb.zmax, z);
                    bb.ymin = min(bb.ymin, y); bb.ymax = max(bb.ymax, y);
                    bb.xmin = min(bb.xmin, bx0); bb.xmax = max(bb.xmax, bx1);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        bb.zmin = min(bb.zmin, __shfl_xor_sync(0xffffffffu, bb.zmin, o)); bb.zmax = max(bb.zmax, __shfl_xor_sync(0xffffffffu, bb.zmax, o));
        bb.ymin = min(bb.ymin, __shfl_xor_sync(0xffffffffu, bb.ymin, o)); bb.ymax = max(bb.ymax, __shfl_xor_sync(0xffffffffu, bb.ymax, o));
        bb.xmin = min(bb.xmin, __shfl_xor_sync(0xffffffffu, bb.xmin, o)); bb.xmax = max(bb.xmax, __shfl_xor_sync(0xffffffffu, bb.xmax, o));
    }
    if (lane == 0 && bb.zmax >= 0) {
        atomicMin(bgbox + 0, bb.zmin); atomicMax(bgbox + 1, bb.zmax);
        atomicMin(bgbox + 2, bb.ymin); atomicMax(bgbox + 3, bb.ymax);
        atomicMin(bgbox + 4, bb.xmin); atomicMax(bgbox + 5, bb.xmax);
    }
}

// iterate the runs (maximal groups of consecutive set bits) of a 64-bit pattern
#define DLV_FOR_RUNS64(pattern, a, len)                                                    \
    for (unsigned long long _m = (pattern); _m;)                                           \
        for (int a = __ffsll(static_cast<long long>(_m)) - 1,                              \
                 len = (~(_m >> a) == 0ull) ? (64 - a) : (__ffsll(static_cast<long long>(~(_m >> a))) - 1), _once = 1; \
             _once; _once = 0, _m &= ~((len >= 64 ? ~0ull : ((1ull << len) - 1ull)) << a))

// ---- P2: unions between touching runs
// One thread per bitmask word ENUMERATES the (run, touching run) pairs of its word - all neighbour words are loaded
// before any is used - and appends them to a per-block queue in shared memory; then the whole block executes the
// queued unions, one pair per thread.  On blob-like masks only ~15 % of the words are non-zero, so executing the
// unions where they are found leaves most lanes idle behind a few lanes that walk ~20 dependent L2 round trips
// each; through the queue the same round trips run side by side.  Pairs beyond the queue capacity (dense masks)
// are executed in place.  Union-find results do not depend on the order of the unions.
constexpr int kMergeThreads = 256;
constexpr int kMergeWords = 4;          // bitmask words per thread: a block scans 1024 words
constexpr int kMergeQueue = 3072;
// The words are scanned first and the non-zero ones COMPACTED into a per-block list; the enumeration below then runs
// on dense warps (one listed word per thread) instead of on the ~5 lanes per warp that happen to own a non-zero word.
template <bool PRUNE>
__global__ void __launch_bounds__(kMergeThreads) ccl_merge_kernel(CclGeom g, const uint32_t* __restrict__ bits, uint32_t* __restrict__ L) {
    __shared__ uint2 queue[kMergeQueue];
    __shared__ uint16_t wlist[kMergeThreads * kMergeWords];
    __shared__ unsigned qn, wn;
    if (threadIdx.x == 0) { qn = 0; wn = 0; }
    __syncthreads();
    const int64_t nwords = g.rows * g.W;
    const int64_t t0 = static_cast<int64_t>(blockIdx.x) * (kMergeThreads * kMergeWords);
#pragma unroll
    for (int k = 0; k < kMergeWords; ++k) {
        const int li = k * kMergeThreads + threadIdx.x;                 // coalesced across the block
        if (t0 + li < nwords && bits[t0 + li]) wlist[atomicAdd(&wn, 1u)] = static_cast<uint16_t>(li);
    }
    __syncthreads();
    const unsigned nlist = wn;
    for (unsigned li = threadIdx.x; li < nlist; li += kMergeThreads) {
        const int64_t t = t0 + wlist[li];
        const uint32_t cur = bits[t];
        auto push = [&](uint32_t a, uint32_t b) {
            const unsigned i = atomicAdd(&qn, 1u);
            if (i < kMergeQueue) queue[i] = make_uint2(a, b); else uf_union(L, a, b);
        };
        const int64_t r = static_cast<uint32_t>(t) / static_cast<uint32_t>(g.W);      // t < 2^32 (checked in ccl_run)
        const int w = static_cast<int>(t - r * g.W);
        const int64_t z = static_cast<uint32_t>(r) / static_cast<uint32_t>(g.Y), y = r - z * g.Y;
        const uint32_t vbase = static_cast<uint32_t>(r * g.X + static_cast<int64_t>(w) * 32) + 1u;   // label of bit 0
        // raster-predecessor rows: (z, y-1), (z-1, y-1), (z-1, y), (z-1, y+1)
        const int64_t nrows[4] = {r - 1, r - g.Y - 1, r - g.Y, r - g.Y + 1};
        const bool ok[4] = {y > 0, z > 0 && y > 0, z > 0, z > 0 && y + 1 < g.Y};
        uint32_t c[4], p[4], q[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            c[k] = p[k] = q[k] = 0u;
            if (ok[k]) {
                const uint32_t* nb = bits + nrows[k] * g.W;
                c[k] = nb[w];
                if (w > 0) p[k] = nb[w - 1];
                if (w + 1 < g.W) q[k] = nb[w + 1];
            }
        }
        const uint32_t prevw = (w > 0 && (cur & 1u)) ? bits[t - 1] : 0u;
        // same row, previous word
        if (prevw >> 31) push(vbase, vbase - 1u);
        // bit i of comb[k] <-> neighbour-row voxel x = w*32 + i - 1, i in [0, 34); label of neighbour bit i is nbase[k] + i
        unsigned long long comb[4];
        uint32_t nbase[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            comb[k] = (static_cast<unsigned long long>(c[k]) << 1) | (p[k] >> 31) | (static_cast<unsigned long long>(q[k] & 1u) << 33);
            nbase[k] = static_cast<uint32_t>(nrows[k] * g.X + static_cast<int64_t>(w) * 32);
        }
        if (!(comb[0] | comb[1] | comb[2] | comb[3])) continue;
        DLV_FOR_RUNS64(static_cast<unsigned long long>(cur), a, len) {
            // run [a, a+len) of the current word touches neighbour bits [a, a+len+2) of every comb
            const unsigned long long span = ((len + 2 >= 64) ? ~0ull : ((1ull << (len + 2)) - 1ull)) << a;
            if (PRUNE) {
                // Most runs of a blob touch runs in all four predecessor rows, and those runs touch each other: rows
                // (z-1,y-1)|(z-1,y) and (z-1,y)|(z-1,y+1) are neighbours inside plane z-1, row (z,y-1) has (z-1,y-1) and
                // (z-1,y) among ITS predecessor rows.  Two such runs within x +- 1 of each other are united by the word
                // that owns the later of them (an earlier row, by induction over the raster order), so once this run is
                // tied to one of them the other union is implied.  Order: (z-1,y) first - it is adjacent to all three
                // others - then (z,y-1), (z-1,y-1), (z-1,y+1); a run is skipped when it touches an already covered one.
                // The rule is restated and checked against a reference 26-connected labelling on random / blob masks in
                // tests/test_cpu_ccl_prune.py: same partition, 55-66 % fewer unions (merge pass 4.6 -> 2.2 ms on cfg3).
                const unsigned long long c0 = comb[0] & span, c1 = comb[1] & span, c2 = comb[2] & span, c3 = comb[3] & span;
                DLV_FOR_RUNS64(c2, i, ilen) { (void)ilen; push(vbase + a, nbase[2] + i); }
                DLV_FOR_RUNS64(c0, i, ilen) {
                    const unsigned long long rb = ((1ull << ilen) - 1ull) << i;       // ilen <= 34
                    if (!((rb | (rb << 1) | (rb >> 1)) & c2)) push(vbase + a, nbase[0] + i);
                }
                DLV_FOR_RUNS64(c1, i, ilen) {
                    const unsigned long long rb = ((1ull << ilen) - 1ull) << i;
                    if (!((rb | (rb << 1) | (rb >> 1)) & (c2 | c0))) push(vbase + a, nbase[1] + i);
                }
                DLV_FOR_RUNS64(c3, i, ilen) {
                    const unsigned long long rb = ((1ull << ilen) - 1ull) << i;
                    if (!((rb | (rb << 1) | (rb >> 1)) & c2)) push(vbase + a, nbase[3] + i);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    DLV_FOR_RUNS64(comb[k] & span, i, ilen) {
                        (void)ilen;
                        push(vbase + a, nbase[k] + i);
                    }
                }
            }
        }
    }
    __syncthreads();
    const unsigned n = min(qn, static_cast<unsigned>(kMergeQueue));
    for (unsigned i = threadIdx.x; i < n; i += kMergeThreads) uf_union(L, queue[i].x, queue[i].y);
}

// ---- P3: resolve roots per run, flag roots (non-zero words compacted per block like in the merge pass)
__global__ void __launch_bounds__(kMergeThreads) ccl_compress_kernel(CclGeom g, const uint32_t* __restrict__ bits, uint32_t* __restrict__ L,
                                                                    uint32_t* __restrict__ rootbits) {
    __shared__ uint16_t wlist[kMergeThreads * kMergeWords];
    __shared__ unsigned wn;
    if (threadIdx.x == 0) wn = 0;
    __syncthreads();
    const int64_t nwords = g.rows * g.W;
    const int64_t t0 = static_cast<int64_t>(blockIdx.x) * (kMergeThreads * kMergeWords);
#pragma unroll
    for (int k = 0; k < kMergeWords; ++k) {
        const int li = k * kMergeThreads + threadIdx.x;
        if (t0 + li < nwords) {
            if (bits[t0 + li]) wlist[atomicAdd(&wn, 1u)] = static_cast<uint16_t>(li);
            else rootbits[t0 + li] = 0u;
        }
    }
    __syncthreads();
    const unsigned nlist = wn;
    for (unsigned li = threadIdx.x; li < nlist; li += kMergeThreads) {
        const int64_t t = t0 + wlist[li];
        const uint32_t cur = bits[t];
        uint32_t rb = 0;
        const int64_t r = static_cast<uint32_t>(t) / static_cast<uint32_t>(g.W);
        const int w = static_cast<int>(t - r * g.W);
        const int64_t vb0 = r * g.X + static_cast<int64_t>(w) * 32;     // 0-based index of bit 0
        DLV_FOR_RUNS64(static_cast<unsigned long long>(cur), a, len) {
            const uint32_t lbl = static_cast<uint32_t>(vb0 + a) + 1u;
            const uint32_t root = uf_find(L, lbl);
            if (root == lbl) rb |= 1u << a;
            else L[vb0 + a] = root;          // shortcut; only run starts are ever traversed
        }
        rootbits[t] = rb;
    }
}

// ---- P4: exclusive scan of popcounts (three small kernels).  Four root-bitmask words per thread (one 16 B load, one
// 16 B store of the prefixes): with one word per thread the two streaming kernels ran at 1.2 - 1.5 TB/s
// (profiles/r02_v_ccl_traffic_cfg3.txt: 0.45 + 0.65 ms for 0.5 GB each way on cfg3).
constexpr int kScanBlock = 1024;
constexpr int kScanItems = 4;
constexpr int kScanSpan = kScanBlock * kScanItems;      // words per block
__device__ __forceinline__ void scan_load_popc(const uint32_t* __restrict__ rootbits, int64_t nwords, int64_t i0, uint32_t (&pc)[kScanItems]) {
    if (i0 + kScanItems <= nwords) {
        const uint4 w = *reinterpret_cast<const uint4*>(rootbits + i0);      // i0 % 4 == 0, pool allocations are 256 B aligned
        pc[0] = __popc(w.x); pc[1] = __popc(w.y); pc[2] = __popc(w.z); pc[3] = __popc(w.w);
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) pc[k] = (i0 + k < nwords) ? __popc(rootbits[i0 + k]) : 0u;
    }
}
__global__ void scan_block_sums_kernel(const uint32_t* __restrict__ rootbits, int64_t nwords, uint32_t* __restrict__ bsum) {
    __shared__ uint32_t sh[32];
    const int64_t i0 = (static_cast<int64_t>(blockIdx.x) * kScanBlock + threadIdx.x) * kScanItems;
    uint32_t pc[kScanItems];
    scan_load_popc(rootbits, nwords, i0, pc);
    uint32_t v = pc[0] + pc[1] + pc[2] + pc[3];
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = sh[threadIdx.x];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) bsum[blockIdx.x] = v;
    }
}
__device__ __forceinline__ uint32_t block_excl_scan_1024(uint32_t v, uint32_t* sh, uint32_t* total) {
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) sh[wp] = inc;
    __syncthreads();
    if (wp == 0) {
        uint32_t s = sh[lane], si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += n;
        }
        sh[lane] = si - s;
        if (lane == 31) *total = si;
    }
    __syncthreads();
    const uint32_t r = sh[wp] + inc - v;
    __syncthreads();
    return r;
}
__global__ void scan_of_block_sums_kernel(uint32_t* __restrict__ bsum, int64_t nb, uint32_t* __restrict__ n_out) {
    __shared__ uint32_t sh[32];
    __shared__ uint32_t tot;
    uint32_t carry = 0;
    for (int64_t base = 0; base < nb; base += kScanBlock) {
        const int64_t i = base + threadIdx.x;
        const uint32_t v = (i < nb) ? bsum[i] : 0u;
        const uint32_t e = block_excl_scan_1024(v, sh, &tot);
        if (i < nb) bsum[i] = carry + e;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = carry;
}
__global__ void scan_apply_kernel(const uint32_t* __restrict__ rootbits, int64_t nwords, const uint32_t* __restrict__ bsum,
                                  uint32_t* __restrict__ wprefix) {
    __shared__ uint32_t sh[32];
    __shared__ uint32_t tot;
    const int64_t i0 = (static_cast<int64_t>(blockIdx.x) * kScanBlock + threadIdx.x) * kScanItems;
    uint32_t pc[kScanItems];
    scan_load_popc(rootbits, nwords, i0, pc);
    const uint32_t e = bsum[blockIdx.x] + block_excl_scan_1024(pc[0] + pc[1] + pc[2] + pc[3], sh, &tot);
    const uint4 o = make_uint4(e, e + pc[0], e + pc[0] + pc[1], e + pc[0] + pc[1] + pc[2]);
    if (i0 + kScanItems <= nwords) {
        *reinterpret_cast<uint4*>(wprefix + i0) = o;
    } else {
        const uint32_t ov[kScanItems] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int k = 0; k < kScanItems; ++k)
            if (i0 + k < nwords) wprefix[i0 + k] = ov[k];
    }
}

// ---- P5: final labels + statistics, one reduction per x-run
template <bool CHECK>
__global__ void __launch_bounds__(kMergeThreads) ccl_relabel_stats_kernel(CclGeom g, const uint32_t* __restrict__ bits, uint32_t* __restrict__ L,
                                         const uint32_t* __restrict__ rootbits, const uint32_t* __restrict__ wprefix,
                                         unsigned long long* __restrict__ cnt, unsigned long long* __restrict__ sums,
                                         int* __restrict__ bbox) {
    __shared__ uint16_t wlist[kMergeThreads * kMergeWords];
    __shared__ unsigned wn;
    if (threadIdx.x == 0) wn = 0;
    __syncthreads();
    const int64_t nwords = g.rows * g.W;
    const int64_t t0 = static_cast<int64_t>(blockIdx.x) * (kMergeThreads * kMergeWords);
#pragma unroll
    for (int k = 0; k < kMergeWords; ++k) {
        const int li = k * kMergeThreads + threadIdx.x;
        if (t0 + li < nwords && bits[t0 + li]) wlist[atomicAdd(&wn, 1u)] = static_cast<uint16_t>(li);
    }
    __syncthreads();
    const unsigned nlist = wn;
    for (unsigned li = threadIdx.x; li < nlist; li += kMergeThreads) {
        const int64_t t = t0 + wlist[li];
        const uint32_t cur = bits[t];
        const int64_t r = static_cast<uint32_t>(t) / static_cast<uint32_t>(g.W);
        const int w = static_cast<int>(t - r * g.W);
        const int z = static_cast<int>(static_cast<uint32_t>(r) / static_cast<uint32_t>(g.Y)), y = static_cast<int>(r - static_cast<int64_t>(z) * g.Y);
        const int64_t vb0 = r * g.X + static_cast<int64_t>(w) * 32;
        DLV_FOR_RUNS64(static_cast<unsigned long long>(cur), a, len) {
            const uint32_t lbl = static_cast<uint32_t>(vb0 + a) + 1u;
            const uint32_t root = ((rootbits[t] >> a) & 1u) ? lbl : L[vb0 + a];
            const int64_t ri = static_cast<int64_t>(root) - 1;
            const int64_t rr = static_cast<uint32_t>(ri) / static_cast<uint32_t>(g.X);
            const int rx = static_cast<int>(ri - rr * g.X);
            const int64_t rw = rr * g.W + (rx >> 5);
            const uint32_t rank = wprefix[rw] + __popc(rootbits[rw] & ((1u << (rx & 31)) - 1u)) + 1u;
            for (int j = 0; j < len; ++j) L[vb0 + a + j] = rank;
            const int xa = w * 32 + a, xb = xa + len - 1;
            atomicAdd(cnt + rank, static_cast<unsigned long long>(len));
            atomicAdd(sums + 3ull * rank + 0, static_cast<unsigned long long>(z) * len);
            atomicAdd(sums + 3ull * rank + 1, static_cast<unsigned long long>(y) * len);
            atomicAdd(sums + 3ull * rank + 2, static_cast<unsigned long long>(xa + xb) * len / 2ull);
            int* b = bbox + 6ull * rank;
            if (CHECK) {
                // A component has ~20 runs and most of them lie inside the box the others have already spanned: read the
                // row (24 B, one or two sectors, from L2) and send only the reductions that can still change it.  The
                // bounds move monotonically, so a stale read can only cause a redundant atomic, never a missing one.
                const int2 bz = __ldcg(reinterpret_cast<const int2*>(b)), by = __ldcg(reinterpret_cast<const int2*>(b + 2)),
                           bx = __ldcg(reinterpret_cast<const int2*>(b + 4));
                if (z < bz.x) atomicMin(b + 0, z);
                if (z > bz.y) atomicMax(b + 1, z);
                if (y < by.x) atomicMin(b + 2, y);
                if (y > by.y) atomicMax(b + 3, y);
                if (xa < bx.x) atomicMin(b + 4, xa);
                if (xb > bx.y) atomicMax(b + 5, xb);
            } else {
                atomicMin(b + 0, z); atomicMax(b + 1, z);
                atomicMin(b + 2, y); atomicMax(b + 3, y);
                atomicMin(b + 4, xa); atomicMax(b + 5, xb);
            }
        }
    }
}

__global__ void bbox_init_kernel(int* __restrict__ bbox, int64_t rows, int Z, int Y, int X) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    int* b = bbox + 6 * i;
    b[0] = Z; b[1] = -1; b[2] = Y; b[3] = -1; b[4] = X; b[5] = -1;
}

// ---- P6: table rows in their final host layout: bounding boxes widened to int64, centroids = sum / count in fp64
// (IEEE division: bit-identical to the host's, count_blobs.py:85), and the foreground totals that define row 0.
__global__ void ccl_table_finish_kernel(int64_t rows, const unsigned long long* __restrict__ cnt, const unsigned long long* __restrict__ sums,
                                        const int* __restrict__ bbox, long long* __restrict__ bbox64, double* __restrict__ cent,
                                        unsigned long long* __restrict__ totals) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    unsigned long long c = 0, sz = 0, sy = 0, sx = 0;
    if (i < rows) {
#pragma unroll
        for (int k = 0; k < 6; ++k) bbox64[6 * i + k] = bbox[6 * i + k];
        if (i > 0) {
            c = cnt[i]; sz = sums[3 * i]; sy = sums[3 * i + 1]; sx = sums[3 * i + 2];
            const double dc = static_cast<double>(c);
            cent[3 * i] = static_cast<double>(sz) / dc;
            cent[3 * i + 1] = static_cast<double>(sy) / dc;
            cent[3 * i + 2] = static_cast<double>(sx) / dc;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, o); sz += __shfl_xor_sync(0xffffffffu, sz, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o); sx += __shfl_xor_sync(0xffffffffu, sx, o);
    }
    if ((threadIdx.x & 31) == 0 && (c | sz | sy | sx)) {
        atomicAdd(totals + 0, c); atomicAdd(totals + 1, sz); atomicAdd(totals + 2, sy); atomicAdd(totals + 3, sx);
    }
}

// The table is returned in ONE pinned host block (counts | sums | bbox | centroids) taken from a small process-wide
// pool, so that the device-to-host copy runs at PCIe speed and a second call of the same size pays neither page
// faults nor cudaHostAlloc.  dlv_table_free returns the block to the pool.
struct TableBox {
    dlv_table pub;      // must stay the first member: dlv_table* <-> TableBox*
    void* block;        // pinned block (nullptr: the arrays are plain calloc memory)
    size_t cap;
};
struct PinnedBlock { void* p; size_t cap; bool busy; };
static std::mutex g_pool_mu;
static std::vector<PinnedBlock> g_pool;
constexpr size_t kPoolKeep = 4;

void* pinned_take(size_t bytes, size_t* cap_out) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    int best = -1;
    for (size_t i = 0; i < g_pool.size(); ++i)
        if (!g_pool[i].busy && g_pool[i].cap >= bytes && (best < 0 || g_pool[i].cap < g_pool[best].cap)) best = static_cast<int>(i);
    if (best >= 0) { g_pool[best].busy = true; *cap_out = g_pool[best].cap; return g_pool[best].p; }
    for (size_t i = 0; i < g_pool.size(); ++i)          // an idle block that is too small makes room for the new one
        if (!g_pool[i].busy) { cudaFreeHost(g_pool[i].p); g_pool.erase(g_pool.begin() + i); break; }
    void* p = nullptr;
    const size_t cap = bytes + bytes / 4 + 4096;
    if (cudaHostAlloc(&p, cap, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    g_pool.push_back({p, cap, true});
    *cap_out = cap;
    return p;
}
void pinned_give(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    size_t idle = 0;
    for (auto& b : g_pool) idle += b.busy ? 0 : 1;
    for (size_t i = 0; i < g_pool.size(); ++i)
        if (g_pool[i].p == p) {
            if (idle >= kPoolKeep) { cudaFreeHost(p); g_pool.erase(g_pool.begin() + i); }
            else g_pool[i].busy = false;
            return;
        }
}
void table_free(dlv_table* t) {
    if (!t) return;
    TableBox* b = reinterpret_cast<TableBox*>(t);
    if (b->block) {
        pinned_give(b->block);
    } else {
        free(t->voxel_counts); free(t->sums); free(t->bbox); free(t->centroids);
    }
    free(b);
}

static unsigned nblocks(int64_t n, int bs) { return static_cast<unsigned>((n + bs - 1) / bs); }

int ccl_run(Ctx* ctx, const uint8_t* mask, const int64_t shape[3], uint32_t* L, dlv_table** table_out) {
    *table_out = nullptr;
    CclGeom g;
    g.Z = shape[0]; g.Y = shape[1]; g.X = shape[2];
    if (g.Z < 0 || g.Y < 0 || g.X < 0) { set_error(ctx, "dlv_ccl: negative shape"); return DLV_ERR_ARG; }
    const int64_t n = g.Z * g.Y * g.X;
    if (n >= 0xFFFFFFFFLL || g.Z > INT_MAX || g.Y > INT_MAX || g.X > INT_MAX) {
        set_error(ctx, "dlv_ccl: %lld voxels exceed the 32-bit label space of one slab; split along z", (long long)n);
        return DLV_ERR_UNSUPPORTED;
    }
    g.rows = g.Z * g.Y;
    g.W = static_cast<int>((g.X + 31) / 32);
    const int64_t nwords = g.rows * g.W;
    TableBox* box = static_cast<TableBox*>(calloc(1, sizeof(TableBox)));
    dlv_table* T = &box->pub;
    uint32_t N = 0;
    int bg[6] = {static_cast<int>(g.Z), -1, static_cast<int>(g.Y), -1, static_cast<int>(g.X), -1};
    uint32_t *bits = nullptr, *rootbits = nullptr, *wprefix = nullptr, *bsum = nullptr, *n_dev = nullptr;
    int* bg_dev = nullptr;
    unsigned long long *cnt = nullptr, *sums = nullptr;
    int* bbox = nullptr;
    long long* bbox64 = nullptr;
    double* cent = nullptr;
    unsigned long long* totals = nullptr;
    unsigned long long tot[4] = {0, 0, 0, 0};
    int rc = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr, e3 = nullptr;
    float ms_a = 0.f, ms_b = 0.f;
    int64_t launches = 0;
#define CK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_error(ctx, "dlv_ccl: %s: %s", #expr, cudaGetErrorString(_e)); rc = DLV_ERR_CUDA; goto done; } } while (0)
    if (n > 0) {
        const int64_t nb = (nwords + kScanSpan - 1) / kScanSpan;
        CK(dmalloc(ctx, &bits, nwords * 4));
        CK(dmalloc(ctx, &rootbits, nwords * 4));
        CK(dmalloc(ctx, &wprefix, nwords * 4));
        CK(dmalloc(ctx, &bsum, nb * 4));
        CK(dmalloc(ctx, &n_dev, 4));
        CK(dmalloc(ctx, &bg_dev, 6 * sizeof(int)));
        CK(cudaMemcpyAsync(bg_dev, bg, sizeof(bg), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2)); CK(cudaEventCreate(&e3));
        CK(cudaEventRecord(e0, ctx->stream));
        {
            const int64_t total_warps = g.rows;     // one row per warp and iteration
            const unsigned grid = static_cast<unsigned>(std::min<int64_t>((total_warps + 7) / 8, static_cast<int64_t>(ctx->num_sms) * 32));
            // rows of whole 32-voxel words in 16 B aligned buffers: one lane per word; anything else: the general form
            const bool words = (g.X % 32) == 0 && (reinterpret_cast<uintptr_t>(mask) & 15u) == 0 && (reinterpret_cast<uintptr_t>(L) & 15u) == 0;
            if (words) {
                const unsigned long long w_magic = ~0ull / static_cast<unsigned long long>(g.W) + 1ull;     // ceil(2^64 / W) for W > 1
                const unsigned long long y_magic = ~0ull / static_cast<unsigned long long>(g.Y) + 1ull;
                if (g.W > 1 && g.Y > 1) {
                    const unsigned wgrid = static_cast<unsigned>(std::min<int64_t>((nwords + 255) / 256, static_cast<int64_t>(ctx->num_sms) * 16));   // 8 warps x 32 words per block and iteration
                    ccl_init_words_kernel<<<wgrid, 256, 0, ctx->stream>>>(mask, g, w_magic, y_magic, bits, L, bg_dev);
                } else {
                    ccl_init_kernel<<<grid, 256, 0, ctx->stream>>>(mask, g, bits, L, bg_dev, 0);
                }
            } else {
                ccl_init_kernel<<<grid, 256, 0, ctx->stream>>>(mask, g, bits, L, bg_dev, 0);
            }
        }
        if (ctx->ccl_prune) ccl_merge_kernel<true><<<nblocks(nwords, kMergeThreads * kMergeWords), kMergeThreads, 0, ctx->stream>>>(g, bits, L);
        else ccl_merge_kernel<false><<<nblocks(nwords, kMergeThreads * kMergeWords), kMergeThreads, 0, ctx->stream>>>(g, bits, L);
        ccl_compress_kernel<<<nblocks(nwords, kMergeThreads * kMergeWords), kMergeThreads, 0, ctx->stream>>>(g, bits, L, rootbits);
        scan_block_sums_kernel<<<static_cast<unsigned>(nb), kScanBlock, 0, ctx->stream>>>(rootbits, nwords, bsum);
        scan_of_block_sums_kernel<<<1, kScanBlock, 0, ctx->stream>>>(bsum, nb, n_dev);
        scan_apply_kernel<<<static_cast<unsigned>(nb), kScanBlock, 0, ctx->stream>>>(rootbits, nwords, bsum, wprefix);
        launches += 6;
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1, ctx->stream));
        CK(cudaMemcpyAsync(&N, n_dev, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(bg, bg_dev, sizeof(bg), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    {
        const size_t rows = static_cast<size_t>(N) + 1;
        T->n = N;
        if (n > 0) {
            // one pinned block: counts [rows] | sums [rows][3] | bbox [rows][6] | centroids [rows][3]
            const size_t bytes = rows * (8 + 24 + 48 + 24);
            box->block = pinned_take(bytes, &box->cap);
            if (!box->block) { set_error(ctx, "dlv_ccl: pinned host table allocation failed (%zu bytes)", bytes); rc = DLV_ERR_CUDA; goto done; }
            uint8_t* hb = static_cast<uint8_t*>(box->block);
            T->voxel_counts = reinterpret_cast<uint64_t*>(hb);
            T->sums = reinterpret_cast<uint64_t*>(hb + rows * 8);
            T->bbox = reinterpret_cast<int64_t*>(hb + rows * 32);
            T->centroids = reinterpret_cast<double*>(hb + rows * 80);
            CK(dmalloc(ctx, &cnt, rows * 8));
            CK(dmalloc(ctx, &sums, rows * 24));
            CK(dmalloc(ctx, &bbox, rows * 24));
            CK(dmalloc(ctx, &bbox64, rows * 48));
            CK(dmalloc(ctx, &cent, rows * 24));
            CK(dmalloc(ctx, &totals, 32));
            CK(cudaMemsetAsync(cnt, 0, rows * 8, ctx->stream));
            CK(cudaMemsetAsync(sums, 0, rows * 24, ctx->stream));
            CK(cudaMemsetAsync(totals, 0, 32, ctx->stream));
            CK(cudaEventRecord(e2, ctx->stream));
            bbox_init_kernel<<<nblocks(rows, 256), 256, 0, ctx->stream>>>(bbox, rows, static_cast<int>(g.Z), static_cast<int>(g.Y), static_cast<int>(g.X));
            if (ctx->ccl_bbox_check)
                ccl_relabel_stats_kernel<true><<<nblocks(nwords, kMergeThreads * kMergeWords), kMergeThreads, 0, ctx->stream>>>(g, bits, L, rootbits, wprefix, cnt, sums, bbox);
            else
                ccl_relabel_stats_kernel<false><<<nblocks(nwords, kMergeThreads * kMergeWords), kMergeThreads, 0, ctx->stream>>>(g, bits, L, rootbits, wprefix, cnt, sums, bbox);
            ccl_table_finish_kernel<<<nblocks(rows, 256), 256, 0, ctx->stream>>>(rows, cnt, sums, bbox, bbox64, cent, totals);
            launches += 3;
            CK(cudaGetLastError());
            CK(cudaEventRecord(e3, ctx->stream));
            CK(cudaMemcpyAsync(T->voxel_counts, cnt, rows * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(T->sums, sums, rows * 24, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(T->bbox, bbox64, rows * 48, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(T->centroids, cent, rows * 24, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(tot, totals, 32, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            CK(cudaEventElapsedTime(&ms_a, e0, e1));
            CK(cudaEventElapsedTime(&ms_b, e2, e3));
        } else {
            T->voxel_counts = static_cast<uint64_t*>(calloc(rows, sizeof(uint64_t)));
            T->sums = static_cast<uint64_t*>(calloc(rows * 3, sizeof(uint64_t)));
            T->bbox = static_cast<int64_t*>(calloc(rows * 6, sizeof(int64_t)));
            T->centroids = static_cast<double*>(calloc(rows * 3, sizeof(double)));
            if (!T->voxel_counts || !T->sums || !T->bbox || !T->centroids) { set_error(ctx, "dlv_ccl: host table allocation failed"); rc = DLV_ERR_ARG; goto done; }
        }
        // row 0 = background: everything that is not foreground (exact integer identities)
        const uint64_t uz = g.Z, uy = g.Y, ux = g.X;
        T->voxel_counts[0] = static_cast<uint64_t>(n) - tot[0];
        T->sums[0] = (uz ? uy * ux * (uz * (uz - 1) / 2) : 0) - tot[1];
        T->sums[1] = (uy ? uz * ux * (uy * (uy - 1) / 2) : 0) - tot[2];
        T->sums[2] = (ux ? uz * uy * (ux * (ux - 1) / 2) : 0) - tot[3];
        for (int k = 0; k < 6; ++k) T->bbox[k] = bg[k];
        for (int k = 0; k < 3; ++k)     // 0 / 0 -> NaN like numpy's divide
            T->centroids[k] = static_cast<double>(T->sums[k]) / static_cast<double>(T->voxel_counts[0]);
    }
    ctx->ccl_ms = ms_a + ms_b;
    ctx->ccl_launches = launches;
    ctx->launches += launches;
done:
#undef CK
    dfree(ctx, bits); dfree(ctx, rootbits); dfree(ctx, wprefix); dfree(ctx, bsum); dfree(ctx, n_dev); dfree(ctx, bg_dev);
    dfree(ctx, cnt); dfree(ctx, sums); dfree(ctx, bbox); dfree(ctx, bbox64); dfree(ctx, cent); dfree(ctx, totals);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (e2) cudaEventDestroy(e2);
    if (e3) cudaEventDestroy(e3);
    if (rc) { table_free(T); return rc; }
    *table_out = T;
    return 0;
}

}  // namespace dlv
