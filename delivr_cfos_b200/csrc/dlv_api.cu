// dlv_api.cu - context management and the extern "C" surface declared in include/delivr_b200.h
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "dlv_internal.h"

namespace dlv {

void set_error(Ctx* ctx, const char* fmt, ...) {
    if (!ctx) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
    va_end(ap);
}

int segment_run(Ctx* ctx, const void* volume_any, const dlv_seg_params* P, void* binaries_any, void* avg_any, void* sig_any,
                dlv_seg_stats* st_out);

static bool dev_ptr(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

}  // namespace dlv

using dlv::Ctx;

static Ctx* C(dlv_ctx* c) { return reinterpret_cast<Ctx*>(c); }
static const Ctx* C(const dlv_ctx* c) { return reinterpret_cast<const Ctx*>(c); }

extern "C" {

int dlv_abi_version(void) { return DLV_ABI_VERSION; }

int dlv_init(int device, dlv_ctx** out) {
    if (!out) return DLV_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return DLV_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DLV_ERR_CUDA;
    Ctx* ctx = new Ctx();
    *out = reinterpret_cast<dlv_ctx*>(ctx);   // returned even on failure so that dlv_last_error works
    ctx->device = device;
    if (prop.major != 10) {
        dlv::set_error(ctx, "device %d is sm_%d%d; this library contains sm_100a code only (no fallback)", device, prop.major, prop.minor);
        return DLV_ERR_UNSUPPORTED;
    }
    ctx->num_sms = prop.multiProcessorCount;
    if (const char* e = getenv("DLV_FUSED")) ctx->use_fused = atoi(e) != 0;
    if (const char* e = getenv("DLV_IS_T")) ctx->is_tiles = atoi(e);
    if (const char* e = getenv("DLV_IS_TX")) ctx->is_tiles_xf = atoi(e);
    if (const char* e = getenv("DLV_IS_NSUB")) ctx->is_nsub = atoi(e);
    if (const char* e = getenv("DLV_IS_TF")) ctx->is_tiles_fold = atoi(e);
    if (const char* e = getenv("DLV_DECONV_STAGES")) ctx->deconv_stages = atoi(e);
    if (const char* e = getenv("DLV_DECONV_XSTORE")) ctx->deconv_xstore = atoi(e) != 0;
    if (const char* e = getenv("DLV_CCL_BBOX_CHECK")) ctx->ccl_bbox_check = atoi(e) != 0;
    if (const char* e = getenv("DLV_CCL_PRUNE")) ctx->ccl_prune = atoi(e) != 0;
    DLV_CUDA_OK(ctx, cudaSetDevice(device));
    DLV_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    DLV_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    {   // keep freed transient buffers in the pool (they are re-used by the next call instead of returned to the driver)
        cudaMemPool_t pool;
        DLV_CUDA_OK(ctx, cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = ~0ull;
        DLV_CUDA_OK(ctx, cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    DLV_CUDA_OK(ctx, cudaEventCreate(&ctx->ev0));
    DLV_CUDA_OK(ctx, cudaEventCreate(&ctx->ev1));
    return DLV_OK;
}

void dlv_destroy(dlv_ctx* c) {
    Ctx* ctx = C(c);
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    dlv::engine_free(ctx);
    dlv::net_free(ctx);
    if (ctx->paint_owner) cudaFree(ctx->paint_owner);
    for (void* p : ctx->scratch) if (p) cudaFree(p);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* dlv_last_error(const dlv_ctx* c) { return c ? C(c)->err : "null context"; }
int64_t dlv_launch_count(const dlv_ctx* c) { return c ? C(c)->launches : 0; }
void* dlv_stream(dlv_ctx* c) { return c ? static_cast<void*>(C(c)->stream) : nullptr; }

int dlv_synchronize(dlv_ctx* c) {
    Ctx* ctx = C(c);
    if (!ctx) return DLV_ERR_ARG;
    DLV_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return DLV_OK;
}

int dlv_set_conv_timing(dlv_ctx* c, int enable) {
    if (!c) return DLV_ERR_ARG;
    C(c)->time_convs = enable != 0;
    if (enable) { C(c)->conv_ms = 0.0; C(c)->stage_ms[0] = C(c)->stage_ms[1] = C(c)->stage_ms[2] = 0.0; }
    return DLV_OK;
}

int dlv_stage_time_ms(const dlv_ctx* c, int stage, double* ms_out) {
    if (!c || !ms_out || stage < 0 || stage > 3) return DLV_ERR_ARG;
    *ms_out = stage == 0 ? C(c)->conv_ms : C(c)->stage_ms[stage - 1];
    return DLV_OK;
}

int dlv_conv_time_ms(const dlv_ctx* c, double* ms_out) {
    if (!c || !ms_out) return DLV_ERR_ARG;
    *ms_out = C(c)->conv_ms;
    return DLV_OK;
}

int dlv_load_weights(dlv_ctx* c, int n, const char* const* names, const float* const* data_host, const int64_t* numel) {
    Ctx* ctx = C(c);
    if (!ctx || !names || !data_host || !numel) return DLV_ERR_ARG;
    cudaSetDevice(ctx->device);
    return dlv::net_load(ctx, n, names, data_host, numel);
}

int dlv_segment(dlv_ctx* c, const void* volume_any, const dlv_seg_params* params, void* binaries_out_any,
                void* avg_logits_out_any, void* sigmoid_out_any, dlv_seg_stats* stats_out) {
    Ctx* ctx = C(c);
    if (!ctx) return DLV_ERR_ARG;
    if (!volume_any || !params || !binaries_out_any) { dlv::set_error(ctx, "dlv_segment: null argument"); return DLV_ERR_ARG; }
    cudaSetDevice(ctx->device);
    return dlv::segment_run(ctx, volume_any, params, binaries_out_any, avg_logits_out_any, sigmoid_out_any, stats_out);
}

int dlv_ccl(dlv_ctx* c, const void* mask_any, const int64_t shape[3], int connectivity, void* labels_out_any,
            dlv_table** table_out) {
    Ctx* ctx = C(c);
    if (!ctx) return DLV_ERR_ARG;
    if (!mask_any || !shape || !table_out) { dlv::set_error(ctx, "dlv_ccl: null argument"); return DLV_ERR_ARG; }
    if (connectivity != 26) { dlv::set_error(ctx, "dlv_ccl: only connectivity 26 (cc3d's default, count_blobs.py:61) is built"); return DLV_ERR_UNSUPPORTED; }
    cudaSetDevice(ctx->device);
    const size_t n = static_cast<size_t>(shape[0]) * shape[1] * shape[2];
    const uint8_t* mask = static_cast<const uint8_t*>(mask_any);
    void *mask_own = nullptr, *lab_own = nullptr;
    if (!dlv::dev_ptr(mask_any)) {
        DLV_CUDA_OK(ctx, dlv::dmalloc(ctx, &mask_own, n));
        cudaError_t e = cudaMemcpyAsync(mask_own, mask_any, n, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { dlv::dfree(ctx, mask_own); dlv::set_error(ctx, "dlv_ccl: mask upload: %s", cudaGetErrorString(e)); return DLV_ERR_CUDA; }
        mask = static_cast<const uint8_t*>(mask_own);
    }
    uint32_t* labels = static_cast<uint32_t*>(labels_out_any);
    const bool lab_host = labels_out_any && !dlv::dev_ptr(labels_out_any);
    if (!labels_out_any || lab_host) {
        cudaError_t e = dlv::dmalloc(ctx, &lab_own, n * 4);
        if (e != cudaSuccess) { dlv::dfree(ctx, mask_own); dlv::set_error(ctx, "dlv_ccl: label buffer: %s", cudaGetErrorString(e)); return DLV_ERR_CUDA; }
        labels = static_cast<uint32_t*>(lab_own);
    }
    int rc = dlv::ccl_run(ctx, mask, shape, labels, table_out);
    if (rc == 0 && lab_host) {
        cudaError_t e = cudaMemcpyAsync(labels_out_any, labels, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { dlv::set_error(ctx, "dlv_ccl: label copy: %s", cudaGetErrorString(e)); rc = DLV_ERR_CUDA; }
    }
    dlv::dfree(ctx, mask_own);
    dlv::dfree(ctx, lab_own);
    return rc;
}

void dlv_table_free(dlv_table* t) { dlv::table_free(t); }

int dlv_ccl_last_timing(const dlv_ctx* c, double* ms_kernels, int64_t* launches) {
    if (!c) return DLV_ERR_ARG;
    if (ms_kernels) *ms_kernels = C(c)->ccl_ms;
    if (launches) *launches = C(c)->ccl_launches;
    return DLV_OK;
}

int dlv_unet_forward(dlv_ctx* c, const uint16_t* windows_dev, int nwin, const int32_t roi[3], float* logits_dev) {
    Ctx* ctx = C(c);
    if (!ctx) return DLV_ERR_ARG;
    if (!windows_dev || !roi || !logits_dev || nwin < 1) { dlv::set_error(ctx, "dlv_unet_forward: bad argument"); return DLV_ERR_ARG; }
    if (!ctx->net.loaded) { dlv::set_error(ctx, "dlv_unet_forward: call dlv_load_weights first"); return DLV_ERR_STATE; }
    cudaSetDevice(ctx->device);
    int rc = dlv::engine_prepare(ctx, roi, std::max(dlv::engine_batch_capacity(ctx), std::min(nwin, 32)));
    if (rc) return rc;
    const int cap = dlv::engine_batch_capacity(ctx);
    // the window stack is a (nwin*rz, ry, rx) volume whose window w starts at plane w*rz
    std::vector<dlv::WindowDesc> wd(nwin);
    for (int w = 0; w < nwin; ++w) wd[w] = dlv::WindowDesc{w * roi[0], 0, 0, 0};
    dlv::WindowDesc* wd_dev = nullptr;
    DLV_CUDA_OK(ctx, dlv::dmalloc(ctx, &wd_dev, sizeof(dlv::WindowDesc) * nwin));
    cudaMemcpyAsync(wd_dev, wd.data(), sizeof(dlv::WindowDesc) * nwin, cudaMemcpyHostToDevice, ctx->stream);
    const int64_t wvox = static_cast<int64_t>(roi[0]) * roi[1] * roi[2];
    for (int off = 0; off < nwin && rc == 0; off += cap) {
        const int n = std::min(cap, nwin - off);
        rc = dlv::engine_run_batch(ctx, windows_dev, roi[1], roi[2], wd_dev + off, n, nullptr, dlv::BlendDev(), logits_dev + off * wvox);
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    dlv::dfree(ctx, wd_dev);
    if (rc == 0 && e != cudaSuccess) { dlv::set_error(ctx, "dlv_unet_forward: %s", cudaGetErrorString(e)); rc = DLV_ERR_CUDA; }
    return rc;
}

int dlv_op_conv3d(dlv_ctx* c, const char* layer_name, const float* x_dev, int n, int D, int H, int W, float* y_dev,
                  double* stats_dev) {
    Ctx* ctx = C(c);
    if (!ctx) return DLV_ERR_ARG;
    if (!ctx->net.loaded) { dlv::set_error(ctx, "dlv_op_conv3d: call dlv_load_weights first"); return DLV_ERR_STATE; }
    cudaSetDevice(ctx->device);
    return dlv::op_conv3d(ctx, layer_name, x_dev, n, D, H, W, y_dev, stats_dev);
}

int dlv_op_deconv(dlv_ctx* c, const char* upcat_name, const float* x_dev, int n, int D, int H, int W, float* y_dev) {
    Ctx* ctx = C(c);
    if (!ctx) return DLV_ERR_ARG;
    if (!ctx->net.loaded) { dlv::set_error(ctx, "dlv_op_deconv: call dlv_load_weights first"); return DLV_ERR_STATE; }
    cudaSetDevice(ctx->device);
    return dlv::op_deconv(ctx, upcat_name, x_dev, n, D, H, W, y_dev);
}

int dlv_op_finalise(dlv_ctx* c, const float* avg_logits_dev, const uint16_t* volume_dev, const int64_t shape_pad[3],
                    const int64_t shape_real[3], float threshold, int erosion_iters, int64_t erosion_block_planes,
                    uint8_t* binaries_dev, float* sigmoid_dev_or_null) {
    Ctx* ctx = C(c);
    if (!ctx) return DLV_ERR_ARG;
    cudaSetDevice(ctx->device);
    return dlv::post_finalise(ctx, avg_logits_dev, volume_dev, shape_pad, shape_real, threshold, erosion_iters,
                              erosion_block_planes, binaries_dev, sigmoid_dev_or_null);
}

}  // extern "C"
