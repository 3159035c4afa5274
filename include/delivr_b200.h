/*
 * delivr_b200.h - C ABI of libdelivr_b200.so, the B200 (sm_100a) implementation of
 * DELiVR's blob_detection hot path.
 *
 * The reference (erturklab/delivr_cfos) is pure Python; its hot path has no FFI.
 * Each entry point below names the reference code it replaces (file:line under
 * the reference tree); INTEGRATION.md shows the ctypes stub a maintainer binds.
 *
 * Conventions
 *   - every function returns 0 on success, a negative code on failure; the
 *     message is available from dlv_last_error(ctx) (thread-unsafe per ctx);
 *   - one ctx per process/GPU, calls on a ctx are serialised by the caller;
 *   - pointers named *_host are host memory, *_dev device memory on the ctx's
 *     GPU; `void* any` pointers may be either (resolved with
 *     cudaPointerGetAttributes);
 *   - the caller owns every buffer it passes; dlv_table is owned by the
 *     library until dlv_table_free;
 *   - there is NO CPU fallback: dlv_init fails unless the device is sm_100.
 *   - volumes are C-order (Z, Y, X), x fastest.
 */
#ifndef DELIVR_B200_H
#define DELIVR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dlv_ctx dlv_ctx;

#define DLV_OK 0
#define DLV_ERR_ARG (-1)
#define DLV_ERR_CUDA (-2)
#define DLV_ERR_STATE (-3)
#define DLV_ERR_UNSUPPORTED (-4)

#define DLV_ABI_VERSION 4

/* ---- lifecycle -------------------------------------------------------- */
int dlv_abi_version(void);
/* Creates a context on CUDA device `device`. Fails on anything but sm_100. */
int dlv_init(int device, dlv_ctx** out);
void dlv_destroy(dlv_ctx* ctx);
const char* dlv_last_error(const dlv_ctx* ctx);
/* Number of kernels this library launched on the ctx since creation (bench "gpu_launches"). */
int64_t dlv_launch_count(const dlv_ctx* ctx);
/* The CUDA stream all work of this ctx is enqueued on (cudaStream_t as void*). */
void* dlv_stream(dlv_ctx* ctx);
int dlv_synchronize(dlv_ctx* ctx);
/* enable != 0: bracket every convolution launch with CUDA events and accumulate the device time into
 * dlv_seg_stats.ms_conv (serialises the stream; used by bench.py for the roofline line only). */
int dlv_set_conv_timing(dlv_ctx* ctx, int enable);
/* Device time accumulated in the convolution kernels since timing was last enabled / the last dlv_segment started
 * (the slab-level entry points do not reset it). */
int dlv_conv_time_ms(const dlv_ctx* ctx, double* ms_out);
/* Same accounting per stage of the window loop while timing is enabled: stage 0 = tcgen05 convolutions (what
 * dlv_conv_time_ms returns), 1 = final 1x1 conv + overlap blend (sliding_window_inferer.py:232-251), 2 = InstanceNorm /
 * Mish / MaxPool passes that are not fused into a convolution, 3 = uint16 window gather (sliding_window_inferer.py:181-195). */
int dlv_stage_time_ms(const dlv_ctx* ctx, int stage, double* ms_out);

/* ---- network weights --------------------------------------------------
 * Replaces BasicUNet(...) + load_state_dict(checkpoint["state_dict"])
 * (inference/inference.py:190-200,217-222).  `names[i]` are the checkpoint keys
 * (with or without the DataParallel "module." prefix), `data_host[i]` host fp32,
 * `numel[i]` element counts.  All 82 tensors must be present (strict load).
 * Weights are repacked to bf16 tcgen05 operand tiles on the device. */
int dlv_load_weights(dlv_ctx* ctx, int n, const char* const* names, const float* const* data_host,
                     const int64_t* numel);

/* ---- segmentation: sliding-window U-Net + blend + binarise -------------
 * Replaces run_inference's compute (inference/inference.py:229-329):
 * SlidingWindowInferer passes (inference/sliding_window_inferer.py:33-253),
 * block-wise averaging (inference.py:285-299) and create_nifti_seg
 * (inference.py:31-95).
 */
typedef struct dlv_seg_params {
    int64_t shape_pad[3];   /* padded volume (Zp,Yp,Xp): multiples of roi (inference.py:229-231) */
    int64_t shape_real[3];  /* original stack shape (Z,Y,X)                                      */
    int32_t roi[3];         /* window (inference.py:164-168); each a multiple of 16             */
    float overlap;          /* 0.5 in the reference (inference.py:125)                          */
    int32_t tta;            /* 0: one pass; 1: the reference's 13-pass plan (inference.py:265-279),
                               noise-free (sigma <= 1e-3 on intensities >= 1 is below bf16 resolution); the
                               13 passes are 5 plain + 4 flip-z + 4 flip-y, evaluated as 3 weighted passes */
    float threshold;        /* sigmoid threshold, 0.5 (inference.py:120)                        */
    int32_t erosion_iters;  /* 30 (inference.py:82)                                             */
    int64_t erosion_block_planes; /* z-extent of the Arrayterator blocks (inference.py:53); <=0: whole volume */
    int32_t blend_mode;     /* 0 constant (what the reference computes, sliding_window_inferer.py:148);
                               1 gaussian (MONAI importance map, sigma_scale 0.125) - no reference oracle */
    int32_t window_batch;   /* windows per launch; <=0: library default                         */
    int32_t skip_empty;     /* 1: windows whose input max <= 0 get -1000 (sliding_window_inferer.py:198-202,
                               applied per window)                                */
    int32_t flip_dim;       /* only when tta == 0: the single pass flips windows along this dim before the net and
                               back after it (sliding_window_inferer.py:218-219,225-226); 0 none, 2 = z, 3 = y, 4 = x */
} dlv_seg_params;

typedef struct dlv_seg_stats {
    int64_t windows_total;
    int64_t windows_active;
    int64_t passes;
    int64_t kernel_launches;
    double ms_unet;      /* device time in the window loop (gather+convs+norms+blend)        */
    double ms_finalise;  /* average + sigmoid/threshold + erosion                             */
    double ms_conv;      /* device time inside the tcgen05 convolution kernels only          */
} dlv_seg_stats;

/* volume: uint16 (Zp,Yp,Xp) host or device.  binaries_out: uint8 (Z,Y,X) host or device.
 * avg_logits_out (optional): float32 (Zp,Yp,Xp) averaged logits (reference keeps fp16: inference.py:242-246).
 * sigmoid_out (optional): float32 (Z,Y,X) (network_output.npy, inference.py:43,72). */
int dlv_segment(dlv_ctx* ctx, const void* volume_any, const dlv_seg_params* params, void* binaries_out_any,
                void* avg_logits_out_any, void* sigmoid_out_any, dlv_seg_stats* stats_out);

/* ---- connected components + statistics --------------------------------
 * Replaces cc3d.connected_components(bin_img, return_N=True) and
 * cc3d.statistics(labels, no_slice_conversion=True) (count_blobs.py:61,64,85).
 * 26-connectivity; labels 1..N numbered by each component's first voxel in
 * C-order raster scan.  Table rows 0..N (row 0 = background), exact integers;
 * centroid = (double)sum / (double)count, one IEEE fp64 division per coordinate
 * (cc3d's definition; evaluated on the device, bit-identical to a host divide).
 * The arrays live in one pinned host block owned by the library (re-used by
 * later calls once the table is freed).
 */
typedef struct dlv_table {
    int64_t n;              /* number of components N                        */
    uint64_t* voxel_counts; /* [N+1]                                         */
    uint64_t* sums;         /* [N+1][3]  sum of z, y, x                      */
    int64_t* bbox;          /* [N+1][6]  zmin,zmax,ymin,ymax,xmin,xmax incl. */
    double* centroids;      /* [N+1][3]  z, y, x (NaN where count == 0)      */
} dlv_table;

/* mask: uint8 (Z,Y,X), host or device, non-zero = foreground.
 * labels_out (optional): uint32 (Z,Y,X), host or device. */
int dlv_ccl(dlv_ctx* ctx, const void* mask_any, const int64_t shape[3], int connectivity, void* labels_out_any,
            dlv_table** table_out);
void dlv_table_free(dlv_table* t);
/* Device time of the last dlv_ccl call's kernels (ms) and their launch count. */
int dlv_ccl_last_timing(const dlv_ctx* ctx, double* ms_kernels, int64_t* launches);

/* ---- operator-level entry points --------------------------------------
 * The same kernels dlv_segment drives, exposed one stage at a time so the
 * parity tests can check each against the oracle / a torch fp32 reference.
 * All pointers here are DEVICE pointers. */

/* U-Net forward on `nwin` windows already in device memory:
 * windows_dev uint16 [nwin][rz][ry][rx] -> logits_dev float32 [nwin][rz][ry][rx].
 * (predictor(window_data), sliding_window_inferer.py:222) */
int dlv_unet_forward(dlv_ctx* ctx, const uint16_t* windows_dev, int nwin, const int32_t roi[3], float* logits_dev);

/* One 3x3x3 convolution layer of the loaded net on NCDHW fp32 input (device):
 * x_dev [n][cin][D][H][W] -> y_dev [n][cout][D][H][W] = raw conv (no bias, pre-norm, bf16-rounded),
 * stats_dev [n][cout][2] = sum, sum of squares of the fp32 accumulators.  layer_name e.g. "conv_0.conv_1". */
int dlv_op_conv3d(dlv_ctx* ctx, const char* layer_name, const float* x_dev, int n, int D, int H, int W, float* y_dev,
                  double* stats_dev);
/* One k2s2 transposed convolution (+bias) of the loaded net, e.g. "upcat_4": [n][cin][D][H][W] -> [n][cout][2D][2H][2W]. */
int dlv_op_deconv(dlv_ctx* ctx, const char* upcat_name, const float* x_dev, int n, int D, int H, int W, float* y_dev);

/* binarise + eroded-mask gate (create_nifti_seg, inference.py:60-88) on device buffers:
 * avg_logits_dev float32 padded (Zp,Yp,Xp); volume_dev uint16 padded; out uint8 (Z,Y,X). */
int dlv_op_finalise(dlv_ctx* ctx, const float* avg_logits_dev, const uint16_t* volume_dev, const int64_t shape_pad[3],
                    const int64_t shape_real[3], float threshold, int erosion_iters, int64_t erosion_block_planes,
                    uint8_t* binaries_dev, float* sigmoid_dev_or_null);

/* ---- slab-level entry points (z-sharded runs: one process per GPU, see delivr_cfos_b200/slabs.py) ----
 * A slab is a contiguous range of planes of the padded volume held by one GPU.  The stages of dlv_segment
 * are exposed separately so that the host can exchange the blended-logit halo and the boundary labels
 * between neighbouring GPUs (NCCL) in between. */

/* Window grid of sliding_window_inferer.py:140-143 (_get_scan_interval + dense_patch_slices): per-dimension
 * window counts and, if starts_out != NULL, the concatenated start lists (z starts, y starts, x starts). */
int dlv_window_grid(const int64_t shape_pad[3], const int32_t roi[3], float overlap, int32_t counts_out[3], int32_t* starts_out);
/* Skip rule (sliding_window_inferer.py:198): active_host[i] = max over window i > 0.  origins are local to the slab. */
int dlv_windows_active(dlv_ctx* ctx, const uint16_t* slab_dev, int64_t SY, int64_t SX, const int32_t* origins_host, int n,
                       const int32_t roi[3], int32_t* active_host);
/* Gather + U-Net + blend for n scheduled windows.  windows_host[i] = {oz, oy, ox, flip_dim (0|2|3|4) | (repeat-1) << 8},
 * origins local to the slab; the window's logits are added `repeat` times (identical noise-free TTA passes,
 * inference.py:269-279, are evaluated once); acc_dev: int32, same extent as the slab, fixed point 2^-12 logit units
 * (+=, order independent). */
/* Where the slab sits in the window grid - needed by the gaussian blend only (blend_mode 1), whose per-voxel weight
 * normalisation depends on the windows that cover a plane globally; NULL for the constant blend. */
typedef struct dlv_blend_geom {
    int64_t shape_pad[3];   /* padded volume (Zp,Yp,Xp) */
    float overlap;
    int64_t gz0;            /* global plane of the slab's first plane */
} dlv_blend_geom;
int dlv_seg_accumulate(dlv_ctx* ctx, const uint16_t* slab_dev, int64_t SY, int64_t SX, const int32_t* windows_host, int n,
                       const int32_t roi[3], int window_batch, int blend_mode, const dlv_blend_geom* geom_or_null, int32_t* acc_dev);
/* In-place int32 sums -> float32 averaged logits for planes [gz0, gz0+nplanes) of the padded volume
 * (inference.py:285-299); active_host is the whole window grid [nz][ny][nx] (skipped windows contribute -1000). */
int dlv_seg_average(dlv_ctx* ctx, int32_t* acc_dev_inout, int64_t nplanes, int64_t gz0, const int64_t shape_pad[3],
                    const int32_t roi[3], float overlap, const int32_t* active_host, int passes, int blend_mode);
/* create_nifti_seg (inference.py:60-88) on a slab: avg/volume hold planes [gz0, gz0+nplanes) with in-plane strides
 * SY,SX; binaries for global planes [oz0, oz1) are written (first plane = oz0).  The slab must contain every plane
 * within erosion_iters of [oz0, oz1) that lies in the same Arrayterator block. */
int dlv_op_finalise_slab(dlv_ctx* ctx, const float* avg_dev, const uint16_t* volume_dev, int64_t SY, int64_t SX, int64_t nplanes,
                         int64_t gz0, const int64_t shape_real[3], float threshold, int erosion_iters, int64_t erosion_block_planes,
                         int64_t oz0, int64_t oz1, uint8_t* binaries_dev, float* sigmoid_dev_or_null);
/* 26-adjacent label pairs (lo label, hi label) across a slab boundary; pairs_dev uint32[cap][2]; *count may exceed cap
 * (then call again with a larger buffer).  Duplicates are possible. */
int dlv_ccl_boundary_pairs(dlv_ctx* ctx, const uint32_t* labels_lo_plane_dev, const uint32_t* labels_hi_plane_dev, int64_t Y, int64_t X,
                           uint32_t* pairs_dev, int64_t cap, int64_t* count_host_out);
/* labels[i] = map[labels[i]] for non-zero labels (local -> global component numbers). */
int dlv_relabel(dlv_ctx* ctx, uint32_t* labels_dev, int64_t n, const uint32_t* map_dev, int64_t nmap);

/* Host-only (no ctx, no GPU): global component numbering for `nslabs` slabs stacked along z.  counts[r] = N_r local
 * components of slab r; pairs[r] = npairs[r] x {label in slab r-1, label in slab r} of 26-adjacent voxels across
 * the seam below slab r (pairs[0] ignored; duplicates allowed).  Components are numbered by their first voxel in
 * raster order (cc3d's order): slabs in z order, a merged component takes the number of its member in the lowest
 * slab.  luts_out[r] (uint32 [N_r + 1], caller-allocated): local label -> global label, 0 -> 0. */
int dlv_resolve_labels(int nslabs, const int64_t* counts, const uint32_t* const* pairs, const int64_t* npairs,
                       uint32_t* const* luts_out, int64_t* n_global_out);

/* Host-only (no ctx, no GPU): exact merge of `ntables` per-slab statistics tables into the global table rows
 * 0..n_global.  Table t has rows[t] rows (local labels 0..N_t) with local z coordinates; luts[t][l] is the global
 * row of local row l (row 0 -> 0), z_offsets[t] the global plane of the slab's first plane; a NULL luts[t] skips
 * the table.  Integer adds / min / max (associative: any slab partition gives the same table), then one fp64
 * divide per centroid coordinate.  Outputs: counts [n+1], sums [n+1][3], bbox [n+1][6], centroids [n+1][3]. */
int dlv_table_merge(int64_t n_global, int ntables, const int64_t* rows, const uint32_t* const* luts,
                    const uint64_t* const* counts, const uint64_t* const* sums, const int64_t* const* bbox,
                    const int64_t* z_offsets, const int64_t shape[3], uint64_t* counts_out, uint64_t* sums_out,
                    int64_t* bbox_out, double* centroids_out);

/* Host-only (no ctx, no GPU): the per-cell CSV text of count_blobs.py:101-114 - what pandas writes for the DataFrame
 * the reference builds row by row: header ",Blob,Coords,Size", then for every label i = 1..n-1 (the reference's
 * range(1, N): the last component is not listed) the line  0,i,"[z, y, x]",size  with the centroid coordinates as
 * Python float repr (shortest digits that round-trip; exponent form below 1e-4 and from 1e16; "nan" / "inf") and the
 * voxel count in decimal.  centroids [n+1][3], voxel_counts [n+1] (table rows 0..n).  Formats on all host threads
 * (a whole brain has 2.5 M rows = 180 MB of text).  Returns the length of the text in bytes and copies it into buf when
 * it fits cap (no terminating NUL); a larger return value than cap means: call again with that much room.  Negative:
 * bad arguments. */
int64_t dlv_table_csv(const double* centroids, const uint64_t* voxel_counts, int64_t n, char* buf, int64_t cap);

/* ---- raw TIFF planes -> device-resident masked volume ----
 * Replaces the masked_nifti.npy producer loop (downsample/downsample_and_mask.py:398-414: cv2.imread(plane, -1),
 * mask rule, copy into the zero-padded array) and get_real_size (downsample_and_mask.py:25-30), so that the
 * 2 B/voxel intermediate file is never written or re-read.  Reader: classic TIFF, II/MM, strips, one channel,
 * 8/16-bit unsigned, compression none(1) / LZW(5) / Deflate(8, 32946) / PackBits(32773), predictor 1 or 2; anything
 * else fails.  dlv_tiff_info / dlv_tiff_read_u16 are host-only (no ctx, no GPU); their message is
 * dlv_tiff_last_error() (thread-local). */
int dlv_tiff_info(const char* path, int64_t* height, int64_t* width, int32_t* bits, int32_t* compression);
/* out_host[height][width]: IFD 0 decoded, 8-bit samples widened (like numpy .astype(uint16)). */
int dlv_tiff_read_u16(const char* path, uint16_t* out_host, int64_t height, int64_t width);
const char* dlv_tiff_last_error(void);
/* Plane i of the slab <- paths[i] (each Y x X): decoded on `nthreads` host threads (<=0: all cores) into pinned memory,
 * read by the GPU over PCIe, masked and written to slab_dev[i][SY][SX] (SY >= Y, SX >= X; the pad is zero-filled,
 * downsample_and_mask.py:391-396).  Mask rule (:405-411): mask_dev_or_null (uint8 [n][Y][X]) != NULL: v *= mask
 * (uint16 wrap-around); else threshold >= 0: v < threshold -> 0; else: unmasked. */
int dlv_load_tiff_planes(dlv_ctx* ctx, const char* const* paths, int n, int64_t Y, int64_t X, int32_t threshold,
                         const uint8_t* mask_dev_or_null, uint16_t* slab_dev, int64_t SY, int64_t SX, int nthreads);

/* Host-only TIFF plane writer for the painter's outputs: plane i of volume_host (n planes of height x width samples,
 * `bits` = 8 or 16, unsigned) -> paths[i], classic little-endian TIFF, strips, compression 1 (none) / 5 (LZW) /
 * 8 (Deflate), planes compressed on `nthreads` host threads (<= 0: all cores).  Replaces the sequential
 * tifffile.imwrite(path, plane, compression='lzw') loops (blob_highlighter.py:127-133, :158-161;
 * blob_depthmap.py:209-213); any baseline reader returns the pixels that were written.  Message:
 * dlv_tiff_write_last_error() (thread-local). */
int dlv_tiff_write_planes(const char* const* paths, int n, const void* volume_host, int64_t height, int64_t width, int32_t bits,
                          int32_t compression, int nthreads);
const char* dlv_tiff_write_last_error(void);

/* ---- blob painter ----
 * Replaces the per-cell bounding-box colouring loops of blob_highlighter.py:107-124 (R/G/B uint8 volumes),
 * :143-151 (region-id uint16 volume) and blob_depthmap.py:198-207 (depth uint16 volume):
 *     for k in 0..n-1:  out_c[box_k] = mask[box_k] * values[k][c]
 * evaluated as out_c[v] = mask[v] * values[K(v)][c] with K(v) the LAST box in order that contains v (0 where no
 * box does) - the same result, overlapping boxes included.  boxes_host int64 [n][6] = z0,z1,y0,y1,x0,x1 are numpy
 * slice bounds (half-open, clipped at the array end; non-negative; already passed through pad_bb);
 * values_host int64 [n][nch]; the product is truncated to the output width like numpy's cast on assignment.
 * mask_any uint8 (Z,Y,X) and out_any[c] (Z,Y,X) of elem_bytes 1 (uint8) or 2 (uint16): host or device, every
 * voxel of every output is written.  chunk_voxels <= 0: library default z-chunk (bounds the 4 B/voxel scratch). */
int dlv_paint_boxes(dlv_ctx* ctx, const void* mask_any, const int64_t shape[3], const int64_t* boxes_host,
                    const int64_t* values_host, int64_t n, int nch, int elem_bytes, void* const* out_any,
                    int64_t chunk_voxels);

/* Exact Euclidean distance transform of the down-sampled mask stack used to depth-code blobs
 * (blob_depthmap.py:174-181): dist = scipy.ndimage.distance_transform_edt(np.pad(stack, 1), sampling)[1:-1,1:-1,1:-1],
 * i.e. distance (in `sampling` units) of every non-zero voxel to the nearest zero voxel, the volume being surrounded
 * by zeros.  nonzero_any uint8 (Z,Y,X) (stack != 0), dist_out_any float64 (Z,Y,X); host or device. */
int dlv_edt(dlv_ctx* ctx, const void* nonzero_any, const int64_t shape[3], const double sampling[3], void* dist_out_any);

#ifdef __cplusplus
}
#endif
#endif /* DELIVR_B200_H */
