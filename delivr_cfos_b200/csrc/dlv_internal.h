// dlv_internal.h - host-side structures shared by the translation units of libdelivr_b200.so
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/delivr_b200.h"

namespace dlv {

typedef __nv_bfloat16 bf16;

// Geometry of one U-Net resolution level for a batch of windows.
//
// Activations live in a zero-haloed, channel-chunked, position-linear layout
//     A[chunk = C/8][S positions][8 channels]   (bf16, 16 B per position and chunk)
// position P = guard + win*Vp + (zp*Yp + yp)*Xp + xp with zp in [0,Z+2), yp in [0,Y+2),
// xp in [0,X+1) and Vp = Zp*Yp*Xp rounded up to a multiple of 128; interior voxel (z,y,x) sits at (z+1, y+1, x+1).  Halo positions and the
// guard zones are zero and are never written, so a 3x3x3 zero-padded convolution is a 1-D
// correlation over P with the 27 constant offsets dz*Yp*Xp + dy*Xp + dx (the single x halo
// column separates consecutive rows).
struct Level {
    int Z, Y, X;
    int Zp, Yp, Xp;
    int YpXp, Vp;
    int guard;   // zero positions in front of window 0 (and at least that many + 1024 behind the last)
    int64_t S;   // positions allocated per chunk
};

struct ConvLayer {
    std::string name;
    int cin = 0;       // real input channels
    int cin_pad = 0;   // multiple of 16 (one tcgen05 K step)
    int cout = 0;
    int nblk = 0;      // MMA N per work item (32 or 64)
    int NB = 0;        // N blocks per position group (conv: cout/nblk, deconv: 8*cout/nblk)
    int KB = 0;        // cin_pad / 16
    int ntaps = 27;    // 27 (3x3x3 conv) or 1 (k2s2 transposed conv, 8 sub-positions folded into NB)
    bf16* w = nullptr;         // packed operand tiles [KB][NB][ntaps][2][nblk][8]
    bf16* w_is = nullptr;      // Cout = 32 conv layers: input-stationary packing [KB][9 (ky,kx)][2][96 = (kz 2,1,0) x 32][8]
    float* gamma = nullptr;    // InstanceNorm affine (conv layers)
    float* beta = nullptr;
    float* bias = nullptr;     // transposed-conv bias
};

struct Net {
    std::map<std::string, ConvLayer> conv;    // "conv_0.conv_0" ... "upcat_1.convs.conv_1"
    std::map<std::string, ConvLayer> deconv;  // "upcat_4" ... "upcat_1"
    float* final_w = nullptr;                 // [32]
    float final_b = 0.f;
    bool loaded = false;
};

struct Engine;  // per-(roi, batch) activation buffers, defined in dlv_unet.cu

struct Ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // host -> device volume chunks of dlv_segment, overlapped with the window passes
    char err[1024] = {0};
    int64_t launches = 0;
    Net net;
    Engine* eng = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double conv_ms = 0.0;       // accumulated device time in conv kernels when timing is enabled
    bool time_convs = false;
    double stage_ms[3] = {0.0, 0.0, 0.0};   // same for the other U-Net passes: 0 overlap blend, 1 norm / pool passes, 2 window gather
    double ccl_ms = 0.0;
    int64_t ccl_launches = 0;
    bool use_fused = true;      // Cout = 32 layers on the input-stationary fused kernel (DLV_FUSED=0 selects the per-tap kernel)
    int is_tiles = 0;           // 0 = heuristic; 2 / 4 force the tile count per column (DLV_IS_T)
    int is_tiles_fold = 0;      // tile count of the uint16 first layer (DLV_IS_TF; 0 = the cost model's choice, two tiles on cfg2:
                                // four measured the same throughput, the layer is bounded by the issuing thread's per-plane
                                // bookkeeping and the four epilogue warps, profiles/r02_m_first_layer.txt)
    bool ccl_prune = true;      // CC merge: skip unions implied by the predecessor rows' own unions (DLV_CCL_PRUNE)
    bool ccl_bbox_check = true; // CC statistics: read a component's box before sending min / max reductions to it (DLV_CCL_BBOX_CHECK)
    bool deconv_xstore = true;  // k2s2 deconvs: output halves exchanged by shuffles so that every store covers 512 contiguous bytes (DLV_DECONV_XSTORE)
    int deconv_stages = 0;      // k2s2 deconvs: smem pipeline depth; 0 = by output size (2 or 8), DLV_DECONV_STAGES forces it
    int is_nsub = 1;            // 64 -> 32 layers: planes staged whole (1) or in two half-plane stages (2, DLV_IS_NSUB)
    int is_tiles_xf = 4;        // same for the 32 -> 32 layers that normalise while staging (DLV_IS_TX): four-tile columns
                                // re-transform less halo (RL / R = 1.26 instead of 1.53 at X = 64), which is what bounds them;
                                // measured on cfg2: 3 316 -> 2 747 clk per 256 positions, 0.482 -> 0.492 Gvoxels/s
    // Grow-only buffers for the per-call gigabyte temporaries of dlv_segment / the finalise stage (blend sums, erosion
    // distances, the uploaded volume, device-side binaries).  They used to come from the stream-ordered pool on every
    // call; when small allocations split a cached block the pool grew again mid-step (observed: +0.2-0.4 s in the
    // finalise stage of every step on one box).  Freed by dlv_destroy, or all at once when one of them cannot grow.
    void* scratch[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t scratch_cap[4] = {0, 0, 0, 0};
    uint32_t* paint_owner = nullptr;   // painter scratch (dlv_paint.cu), all-zero between calls when paint_owner_clean
    size_t paint_owner_cap = 0;        // voxels
    bool paint_owner_clean = false;
};

void set_error(Ctx* ctx, const char* fmt, ...);

// Transient device buffers come from the device's stream-ordered pool (release threshold raised in dlv_init, so a
// second call of the same size re-uses the memory instead of paying cudaMalloc / cudaFree for gigabytes).
inline cudaError_t dmalloc(Ctx* ctx, void** p, size_t bytes) { return cudaMallocAsync(p, bytes ? bytes : 1, ctx->stream); }
template <class T> inline cudaError_t dmalloc(Ctx* ctx, T** p, size_t bytes) { return dmalloc(ctx, reinterpret_cast<void**>(p), bytes); }
inline void dfree(Ctx* ctx, void* p) { if (p) cudaFreeAsync(p, ctx->stream); }

enum { kScratchAcc = 0, kScratchDist = 1, kScratchSlab = 2, kScratchBin = 3 };
// -> device pointer of at least `bytes` (contents undefined), valid until the next request for the same slot.  The
// caller orders its use on ctx->stream; growing synchronises the device (cudaFree / cudaMalloc).
inline cudaError_t scratch_get(Ctx* ctx, int slot, size_t bytes, void** out) {
    if (ctx->scratch_cap[slot] < bytes) {
        if (ctx->scratch[slot]) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->scratch[slot]); ctx->scratch[slot] = nullptr; ctx->scratch_cap[slot] = 0; }
        cudaError_t e = cudaMalloc(&ctx->scratch[slot], bytes);
        if (e != cudaSuccess) {            // make room: drop the other slots and try once more
            cudaGetLastError();
            cudaStreamSynchronize(ctx->stream);
            for (int i = 0; i < 4; ++i) { if (ctx->scratch[i]) cudaFree(ctx->scratch[i]); ctx->scratch[i] = nullptr; ctx->scratch_cap[i] = 0; }
            e = cudaMalloc(&ctx->scratch[slot], bytes);
            if (e != cudaSuccess) { ctx->scratch[slot] = nullptr; return e; }
        }
        ctx->scratch_cap[slot] = bytes;
    }
    *out = ctx->scratch[slot];
    return cudaSuccess;
}

// Device time of a stage when dlv_set_conv_timing is on (events on the library stream; serialises it - bench only).
enum { kStageBlend = 0, kStageNorm = 1, kStageGather = 2 };
struct StageTimer {
    Ctx* c; int stage;
    StageTimer(Ctx* ctx, int st) : c(ctx), stage(st) { if (c->time_convs) cudaEventRecord(c->ev0, c->stream); }
    ~StageTimer() {
        if (!c->time_convs) return;
        cudaEventRecord(c->ev1, c->stream);
        cudaEventSynchronize(c->ev1);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->stage_ms[stage] += ms;
    }
};

#define DLV_CUDA_OK(ctx, expr)                                                              \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            dlv::set_error((ctx), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                           __FILE__, __LINE__);                                             \
            return -2;                                                                      \
        }                                                                                   \
    } while (0)

// dlv_unet.cu
struct WindowDesc {
    int32_t oz, oy, ox;   // window origin inside the device-resident slab
    int32_t flip;         // bits 0-7: 0 none, 1 flip z, 2 flip y, 3 flip x (reference flip_dim 2 / 3 / 4);
                          // bits 8-15: repeat - 1 (the window's logits are blended `repeat` times)
};
int net_load(Ctx* ctx, int n, const char* const* names, const float* const* data, const int64_t* numel);
void net_free(Ctx* ctx);
void engine_free(Ctx* ctx);
int engine_prepare(Ctx* ctx, const int32_t roi[3], int batch);
int engine_batch_capacity(Ctx* ctx);
// gather windows described by wd_dev[0..nwin) from a uint16 slab, run the net, blend into acc (fp32, 2 planes sets)
// Gaussian importance weights of the optional blend (device pointers; all null = constant blend): wz/wy/wx[roi] the
// separable window weights, nz/ny/nx the reciprocal of the largest weight any covering window gives a voxel along that
// axis, indexed by the voxel's slab-local coordinate (nz is pre-offset to the slab's first plane).  Scaling every
// contribution of a voxel by the same nz*ny*nx cancels in the average and keeps the fixed-point sums well scaled
// where all covering windows see the voxel at their periphery (raw weights ~1e-5 next to the volume faces).
struct BlendDev {
    const float *wz = nullptr, *wy = nullptr, *wx = nullptr, *nz = nullptr, *ny = nullptr, *nx = nullptr;
};
int engine_run_batch(Ctx* ctx, const uint16_t* slab, int64_t slabY, int64_t slabX, const WindowDesc* wd_dev, int nwin,
                     int32_t* acc, const BlendDev& bw, float* logits_out);
int windows_active(Ctx* ctx, const uint16_t* slab, int64_t slabY, int64_t slabX, const int32_t* origins_dev, int n,
                   const int32_t roi[3], int32_t* active_dev);
int op_conv3d(Ctx* ctx, const char* name, const float* x, int n, int D, int H, int W, float* y, double* stats);
int op_deconv(Ctx* ctx, const char* name, const float* x, int n, int D, int H, int W, float* y);

// dlv_post.cu
int post_finalise(Ctx* ctx, const float* avg, const uint16_t* vol, const int64_t sp[3], const int64_t sr[3], float thr,
                  int iters, int64_t block_planes, uint8_t* bin, float* sig);
int post_finalise_slab(Ctx* ctx, const float* avg, const uint16_t* vol, int64_t SY, int64_t SX, int64_t nplanes, int64_t gz0,
                       const int64_t sr[3], float thr, int iters, int64_t block_planes, int64_t oz0, int64_t oz1,
                       uint8_t* bin, float* sig);

// fixed-point scale of the blend accumulator (logit units of 2^-12; |contribution| clamped to 2000)
constexpr float kAccScale = 4096.f;
constexpr float kAccClamp = 2000.f;
constexpr float kSkipLogit = -1000.f;   // sliding_window_inferer.py:199-200

// dlv_ccl.cu
int ccl_run(Ctx* ctx, const uint8_t* mask_dev, const int64_t shape[3], uint32_t* labels_dev, dlv_table** table_out);
void table_free(dlv_table* t);
// pinned host blocks from the library's small process-wide pool (table rows, painter box lists)
void* pinned_take(size_t bytes, size_t* cap_out);
void pinned_give(void* p);

// dlv_paint.cu
int paint_boxes(Ctx* ctx, const void* mask_any, const int64_t shape[3], const int64_t* boxes_host, const int64_t* values_host,
                int64_t n, int nch, int elem_bytes, void* const* out_any, int64_t chunk_voxels);

}  // namespace dlv
