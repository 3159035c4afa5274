"""numpy/torch-CPU restatement of the reference hot path (TEST INFRASTRUCTURE).

Each function cites the reference lines it follows.  The data flow is kept
literal (fp16 running sum, uint8 count map, batch-level skip rule,
Arrayterator blocking) so that it can be compared 1:1 with the unmodified
reference run through ``oracle/shims`` (``make_golden.py``), and then used on
the GPU box - where /root/reference does not exist - as the checker for the
CUDA path.
"""
import math

import numpy as np
import torch

from . import ccl_ref

ARRAYTERATOR_BUF = 1000 ** 3          # inference/inference.py:53,285
EROSION_ITERS = 30                    # inference/inference.py:82
SKIP_VALUE = -1000.0                  # inference/sliding_window_inferer.py:199-200


# ---------------------------------------------------------------- geometry
def padded_shape(shape_real, roi):
    """inference/inference.py:229-231."""
    return tuple(int(np.ceil(d / r) * r) for d, r in zip(shape_real, roi))


def scan_interval(image_size, roi, overlap):
    """inference/sliding_window_inferer.py:255-276."""
    out = []
    for i in range(len(roi)):
        if roi[i] == image_size[i]:
            out.append(int(roi[i]))
        else:
            iv = int(roi[i] * (1 - overlap))
            out.append(iv if iv > 0 else 1)
    return tuple(out)


def window_starts(image_size, roi, interval):
    """Per-dim start lists of MONAI 1.2.0 ``dense_patch_slices`` (sliding_window_inferer.py:143)."""
    starts = []
    for dim in range(len(roi)):
        if interval[dim] == 0:
            num = 1
        else:
            n = int(math.ceil(float(image_size[dim]) / interval[dim]))
            d = next((d for d in range(n) if d * interval[dim] + roi[dim] >= image_size[dim]), None)
            num = d + 1 if d is not None else 1
        s = []
        for idx in range(num):
            st = idx * interval[dim]
            st -= max(st + roi[dim] - image_size[dim], 0)
            s.append(st)
        starts.append(s)
    return starts


def window_list(image_size, roi, overlap):
    """All window origins, first dim slowest (meshgrid ij order)."""
    st = window_starts(image_size, roi, scan_interval(image_size, roi, overlap))
    return [(z, y, x) for z in st[0] for y in st[1] for x in st[2]]


def arrayterator_blocks(shape, buf=ARRAYTERATOR_BUF):
    """Block z-ranges numpy.lib.Arrayterator(buf) yields for a (Z,Y,X) array whose planes fit the buffer.

    inference/inference.py:53,285.  Blocks are full-XY slabs of
    floor(floor(buf/X)/Y) planes (the case Y*X <= buf; the y-split case of
    giant planes is outside what the oracle restates).
    """
    Z, Y, X = shape
    count = buf // X
    if count <= Y:
        raise NotImplementedError("planes larger than the Arrayterator buffer")
    count //= Y
    bz = Z if count > Z else count
    return [(z0, min(Z, z0 + bz)) for z0 in range(0, Z, bz)]


# ---------------------------------------------------------------- sliding window
def sliding_window_pass(volume, roi, overlap, predictor, sw_batch_size, out_sum, count_map,
                        flip_dim=None, threshold=0):
    """One call of ``sliding_window_inference`` (sliding_window_inferer.py:140-251), noise-free.

    volume   (Zp,Yp,Xp) uint16;  out_sum (Zp,Yp,Xp) float16 (+=);  count_map uint8/float16 (+=1)
    predictor: callable (B,1,rz,ry,rx) float32 tensor -> same-shape float32 logits.
    flip_dim uses the reference's 5-D numbering (2 = z, 3 = y, 4 = x).
    """
    wins = window_list(volume.shape, roi, overlap)
    rz, ry, rx = roi
    for g in range(0, len(wins), sw_batch_size):
        batch = wins[g:g + sw_batch_size]
        data = np.stack([volume[z:z + rz, y:y + ry, x:x + rx] for z, y, x in batch]).astype(np.int32)
        if data.max() <= threshold:                                   # :198-202 (whole batch)
            seg = np.full(data.shape, SKIP_VALUE, dtype=np.float16)
        else:
            t = torch.as_tensor(data, dtype=torch.float32)[:, None]  # :207
            if flip_dim is not None:
                t = torch.flip(t, dims=[flip_dim])                    # :218-219
            with torch.no_grad():
                p = predictor(t)
            if flip_dim is not None:
                p = torch.flip(p, dims=[flip_dim])                    # :225-226
            seg = p[:, 0].to(torch.float16).numpy()                   # :229
        for (z, y, x), s in zip(batch, seg):                          # :232-251
            out_sum[z:z + rz, y:y + ry, x:x + rx] += s
            count_map[z:z + rz, y:y + ry, x:x + rx] += 1


TTA_PLAN = [(None,)] + [(None,), (2,), (3,)] * 4                      # inference.py:265-279


def infer_average(volume, roi, overlap, predictor, sw_batch_size, tta=False):
    """run_inference's accumulation + block-wise averaging (inference.py:240-299). -> float16 (Zp,Yp,Xp)."""
    out_sum = np.zeros(volume.shape, dtype=np.float16)
    count = np.zeros(volume.shape, dtype=np.uint8)
    plan = TTA_PLAN if tta else TTA_PLAN[:1]
    for (flip,) in plan:
        sliding_window_pass(volume, roi, overlap, predictor, sw_batch_size, out_sum, count, flip_dim=flip)
    with np.errstate(over="ignore", invalid="ignore"):
        return (torch.as_tensor(out_sum) / torch.as_tensor(count)).numpy()   # fp16 / uint8 -> fp16


# ---------------------------------------------------------------- optional Gaussian blend (not what the reference computes)
def gaussian_importance_map(roi, sigma_scale=0.125):
    """MONAI ``compute_importance_map(mode="gaussian")`` as SURVEY.md section 8(c) restates it: separable
    exp(-x^2 / (2 (sigma_scale n)^2)), x = -(n-1)/2 .. (n-1)/2 per dimension, float32.  The reference itself
    hard-codes mode='constant' (sliding_window_inferer.py:148), so this pins the library's optional blend_mode = 1 to
    the restated formula only: PARITY UNPINNED against MONAI (not installed, not exercised by the reference)."""
    w = None
    for n in roi:
        x = np.arange(n, dtype=np.float32) - np.float32((n - 1) / 2.0)
        g = np.exp(-(x.astype(np.float64) ** 2) / (2.0 * (sigma_scale * n) ** 2)).astype(np.float32)
        w = g if w is None else w[..., None] * g
    return w


def infer_average_weighted(volume, roi, overlap, predictor, weights):
    """Importance-weighted blend: sum_w(weight * logit) / sum_w(weight) over the windows covering a voxel, windows whose
    input is all zero contributing the skip value (per window).  float64 accumulation -> float32 (Zp,Yp,Xp)."""
    num = np.zeros(volume.shape, dtype=np.float64)
    den = np.zeros(volume.shape, dtype=np.float64)
    rz, ry, rx = roi
    w64 = weights.astype(np.float64)
    for (z, y, x) in window_list(volume.shape, roi, overlap):
        data = volume[z:z + rz, y:y + ry, x:x + rx].astype(np.int32)
        if data.max() <= 0:
            seg = np.full(data.shape, SKIP_VALUE, dtype=np.float64)
        else:
            with torch.no_grad():
                seg = predictor(torch.as_tensor(data, dtype=torch.float32)[None, None])[0, 0].numpy().astype(np.float64)
        num[z:z + rz, y:y + ry, x:x + rx] += w64 * seg
        den[z:z + rz, y:y + ry, x:x + rx] += w64
    return (num / den).astype(np.float32)


# ---------------------------------------------------------------- binarise + eroded mask
def create_binaries(avg_logits, volume, shape_real, threshold=0.5, return_sigmoid=False):
    """create_nifti_seg (inference.py:31-95): sigmoid >= thr AND erode30(input > 0), per Arrayterator block."""
    Z, Y, X = shape_real
    out = np.zeros((Z, Y, X), dtype=np.uint8)
    sig_out = np.zeros((Z, Y, X), dtype=np.float32) if return_sigmoid else None
    for z0, z1 in arrayterator_blocks((Z, Y, X)):
        sub = torch.as_tensor(np.ascontiguousarray(avg_logits[z0:z1, :Y, :X]), dtype=torch.float)
        sig = sub.sigmoid().numpy()
        if return_sigmoid:
            sig_out[z0:z1] = sig
        mask = (volume[z0:z1, :Y, :X] > 0).astype(np.uint8)
        mask = ccl_ref.erode6(mask, EROSION_ITERS)
        out[z0:z1] = (sig >= threshold).astype(np.uint8) * mask
    return (out, sig_out) if return_sigmoid else out


# ---------------------------------------------------------------- connected components table
def blob_table(binaries):
    """count_blobs.py:61,85: labels, N and the statistics dict (plus exact integer sums)."""
    labels, n = ccl_ref.connected_components26(binaries)
    return labels, n, ccl_ref.statistics(labels, n)


def csv_text(stats, n):
    """Text of the per-cell CSV (count_blobs.py:101-114): rows for labels 1..N-1, index column always 0."""
    lines = [",Blob,Coords,Size"]
    cent, cnt = stats["centroids"], stats["voxel_counts"]
    for i in range(1, n):
        lines.append(f'0,{i},"{[float(c) for c in cent[i]]}",{int(cnt[i])}')
    return "\n".join(lines) + "\n"


def csv_name(shape_real, brain):
    """count_blobs.py:113."""
    return f"{tuple(int(s) for s in shape_real)}_{brain.replace('.nii.gz', '')}.csv"


# ---------------------------------------------------------------- synthetic inputs
def synth_volume(shape, seed, roi=None):
    """Seeded uint16 test volume: ellipsoid 'brain' (0 outside, >=1 inside) with blob-like cells.

    Fallback generator of SURVEY.md section 8(d) (the shipped cFos patches are not
    available on the GPU box): lognormal background + Gaussian blobs.  If ``roi``
    is given the result is zero-padded at the high end to window multiples, like
    downsample_and_mask.py:391-396 does for masked_nifti.npy.
    """
    rng = np.random.default_rng(seed)
    Z, Y, X = shape
    zz, yy, xx = np.ogrid[:Z, :Y, :X]
    ell = (((zz - (Z - 1) / 2) / (0.45 * Z)) ** 2 + ((yy - (Y - 1) / 2) / (0.45 * Y)) ** 2
           + ((xx - (X - 1) / 2) / (0.45 * X)) ** 2) <= 1.0
    vol = np.exp(rng.normal(7.4, 0.35, size=shape)).astype(np.float32)
    nblob = max(1, int(370 * (Z * Y * X) / 1e6))
    cz = rng.integers(0, Z, nblob); cy = rng.integers(0, Y, nblob); cx = rng.integers(0, X, nblob)
    sg = rng.uniform(1.5, 3.0, nblob); amp = rng.uniform(2000, 30000, nblob)
    r = 8
    g = np.arange(-r, r + 1)
    for i in range(nblob):
        z0, z1 = max(0, cz[i] - r), min(Z, cz[i] + r + 1)
        y0, y1 = max(0, cy[i] - r), min(Y, cy[i] + r + 1)
        x0, x1 = max(0, cx[i] - r), min(X, cx[i] + r + 1)
        gz = np.exp(-0.5 * ((np.arange(z0, z1) - cz[i]) / sg[i]) ** 2)
        gy = np.exp(-0.5 * ((np.arange(y0, y1) - cy[i]) / sg[i]) ** 2)
        gx = np.exp(-0.5 * ((np.arange(x0, x1) - cx[i]) / sg[i]) ** 2)
        vol[z0:z1, y0:y1, x0:x1] += amp[i] * gz[:, None, None] * gy[None, :, None] * gx[None, None, :]
    del g
    vol = np.clip(vol, 1, 65535).astype(np.uint16)
    vol[~ell] = 0
    if roi is not None:
        ps = padded_shape(shape, roi)
        out = np.zeros(ps, dtype=np.uint16)
        out[:Z, :Y, :X] = vol
        return out
    return vol


def synth_mask(shape, seed, kind="blobs", p=0.08):
    """Seeded binary masks for config 3: blob field, or Bernoulli(p)."""
    rng = np.random.default_rng(seed)
    Z, Y, X = shape
    if kind == "bernoulli":
        return (rng.random(shape) < p).astype(np.uint8)
    m = np.zeros(shape, dtype=np.uint8)
    n = max(1, int(370 * (Z * Y * X) / 1e6))
    cz = rng.integers(0, Z, n); cy = rng.integers(0, Y, n); cx = rng.integers(0, X, n)
    offs = [(a, b, c) for a in range(-2, 3) for b in range(-2, 3) for c in range(-2, 3) if a * a + b * b + c * c <= 5]
    for a, b, c in offs:
        z = np.clip(cz + a, 0, Z - 1); y = np.clip(cy + b, 0, Y - 1); x = np.clip(cx + c, 0, X - 1)
        m[z, y, x] = 1
    return m
