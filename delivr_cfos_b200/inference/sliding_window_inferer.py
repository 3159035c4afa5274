"""Drop-in for the reference's ``inference/sliding_window_inferer.py`` (:33-370).

``SlidingWindowInferer`` keeps the reference's constructor and call signature.
The ``network`` argument must be a :class:`DelivrNet` - the handle of the
network whose weights were loaded into the CUDA library - instead of a torch
module: gather, U-Net, blend and count all run on the GPU in ``dlv_segment``.
As in the reference the constructor's ``mode`` is accepted but the blend is
*constant* (sliding_window_inferer.py:148 hard-codes ``mode='constant'``); the
non-reference Gaussian blend is available as ``blend="gaussian"`` on
``run_inference`` only.
"""
import datetime
import math

import numpy as np

from .._lib import Context

__all__ = ["sliding_window_inference", "SlidingWindowInferer", "DelivrNet", "dense_window_starts"]


class DelivrNet:
    """The loaded BasicUNet living on one B200 (replaces the torch module of inference.py:190-222)."""

    def __init__(self, checkpoint_path=None, state_dict=None, device=0):
        self.ctx = Context(device)
        if checkpoint_path is not None:
            self.ctx.load_checkpoint(str(checkpoint_path))
        elif state_dict is not None:
            self.ctx.load_weights(state_dict)
        else:
            raise ValueError("DelivrNet needs a checkpoint path or a state_dict")

    def eval(self):
        return self

    def to(self, *_a, **_k):
        return self


def _get_scan_interval(image_size, roi_size, num_spatial_dims, overlap):
    """sliding_window_inferer.py:255-276."""
    if len(image_size) != num_spatial_dims:
        raise ValueError("image coord different from spatial dims.")
    if len(roi_size) != num_spatial_dims:
        raise ValueError("roi coord different from spatial dims.")
    scan_interval = []
    for i in range(num_spatial_dims):
        if roi_size[i] == image_size[i]:
            scan_interval.append(int(roi_size[i]))
        else:
            interval = int(roi_size[i] * (1 - overlap))
            scan_interval.append(interval if interval > 0 else 1)
    return tuple(scan_interval)


def dense_window_starts(image_size, roi_size, overlap):
    """Per-dimension window starts of MONAI's dense_patch_slices as used at sliding_window_inferer.py:143."""
    interval = _get_scan_interval(image_size, roi_size, len(roi_size), overlap)
    starts = []
    for d in range(len(roi_size)):
        num = int(math.ceil(float(image_size[d]) / interval[d]))
        first = next((k for k in range(num) if k * interval[d] + roi_size[d] >= image_size[d]), None)
        cnt = first + 1 if first is not None else 1
        s = []
        for k in range(cnt):
            st = k * interval[d]
            st -= max(st + roi_size[d] - image_size[d], 0)
            s.append(st)
        starts.append(s)
    return starts


def cover_count(image_size, roi_size, overlap):
    """count_map of one pass (sliding_window_inferer.py:251) - separable, computed analytically. uint8 (Z,Y,X)."""
    starts = dense_window_starts(image_size, roi_size, overlap)
    per_dim = []
    for d in range(3):
        c = np.zeros(image_size[d], dtype=np.int64)
        for s in starts[d]:
            c[s:s + roi_size[d]] += 1
        per_dim.append(c)
    return (per_dim[0][:, None, None] * per_dim[1][None, :, None] * per_dim[2][None, None, :]).astype(np.uint8)


def sliding_window_inference(inputs, roi_size, sw_batch_size, predictor, overlap=0.25, mode="constant",
                             sigma_scale=0.125, padding_mode="constant", cval=0.0, sw_device=None, device=None,
                             SIGMOID=False, output_image=None, count_map=None, tta=None, flip_dim=None,
                             window_data_threshold=0, *args, **kwargs):
    """One sliding-window pass accumulated into ``output_image`` / ``count_map`` (sliding_window_inferer.py:33-253).

    inputs: uint16 array-like (1,1,Zp,Yp,Xp); predictor: DelivrNet; output_image / count_map: torch tensors or
    numpy arrays of the same shape, updated in place (``+=``) exactly like the reference's CPU loop.
    ``tta`` (noise) is accepted and ignored: N(0, U(0,1e-3)) on raw intensities is below the arithmetic's resolution.
    """
    if overlap < 0 or overlap >= 1:
        raise AssertionError("overlap must be >= 0 and < 1.")
    if not isinstance(predictor, DelivrNet):
        raise TypeError("predictor must be a DelivrNet (the CUDA-resident network); torch modules are not run here")
    if window_data_threshold != 0:
        raise NotImplementedError("window_data_threshold other than 0")
    if output_image is None or count_map is None:
        raise ValueError("output_image and count_map are required (the reference accumulates in place)")
    image_size = tuple(int(s) for s in inputs.shape[2:])
    roi = tuple(int(r) if r and r > 0 else image_size[i] for i, r in enumerate(roi_size))   # fall_back_tuple (:111)
    if any(image_size[i] < roi[i] for i in range(3)):
        raise NotImplementedError("volumes smaller than the window (the reference's reflect-pad branch, :119-136)")
    print("Inferring...")
    vol = np.ascontiguousarray(np.asarray(inputs)[0, 0]).astype(np.uint16, copy=False)   # the reference casts per window (:181-195,207)
    avg = np.empty(image_size, dtype=np.float32)
    scratch = np.empty(image_size, dtype=np.uint8)
    predictor.ctx.segment(vol, image_size, image_size, roi, scratch, overlap=overlap, tta=False, flip_dim=flip_dim,
                          erosion_iters=0, window_batch=0, avg_logits_out=avg)
    count = cover_count(image_size, roi, overlap)
    seg_sum = avg * count                                    # the pass's summed logits
    oi = output_image.numpy() if hasattr(output_image, "numpy") else output_image
    cm = count_map.numpy() if hasattr(count_map, "numpy") else count_map
    with np.errstate(over="ignore"):
        oi[0, 0] += seg_sum.astype(oi.dtype)
        cm[0, 0] += count.astype(cm.dtype)
    print(f"{datetime.datetime.now()} : Inference run finished")


class SlidingWindowInferer:
    """Same constructor / call as the reference class (sliding_window_inferer.py:278-370)."""

    def __init__(self, roi_size, sw_batch_size=1, overlap=0.25, mode="constant", sigma_scale=0.125,
                 padding_mode="constant", cval=0.0, sw_device=None, device=None):
        if str(getattr(mode, "value", mode)) not in ("constant", "gaussian"):
            raise ValueError(f"unsupported blend mode {mode!r}")     # BlendMode(mode) at :336
        self.roi_size = roi_size
        self.sw_batch_size = sw_batch_size
        self.overlap = overlap
        self.mode = mode
        self.sigma_scale = sigma_scale
        self.padding_mode = padding_mode
        self.cval = cval
        self.sw_device = sw_device
        self.device = device

    def __call__(self, inputs, network, *args, **kwargs):
        return sliding_window_inference(inputs, self.roi_size, self.sw_batch_size, network, self.overlap, self.mode,
                                        self.sigma_scale, self.padding_mode, self.cval, self.sw_device, self.device,
                                        *args, **kwargs)
