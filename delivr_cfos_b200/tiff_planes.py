"""Raw TIFF planes -> device-resident masked volume, without the ``masked_nifti.npy`` round trip.

Mirrors the part of the reference's ``downsample/downsample_and_mask.py`` that feeds blob_detection:

* ``get_real_size``  (:25-30)   - (Z, Y, X) from the plane count and one plane's shape;
* the masking loop   (:398-414) - ``cv2.imread(plane, -1)``, ``img[img < threshold] = 0`` (simple-threshold mode) or
  ``img *= mask[z]`` (Ilastik mode), copy into the array zero-padded to multiples of the window.

Decoding (C++ host threads) and masking/padding (CUDA) run inside ``dlv_load_tiff_planes``; the result is the uint16
``(Zp, Yp, Xp)`` tensor that ``run_inference(..., volume=...)`` consumes directly.
"""
import math
import os

import numpy as np

from ._lib import tiff_info, tiff_read_u16  # noqa: F401


def list_planes(raw_location):
    """Plane files in z order: ``sorted(x for x in os.listdir(raw_location) if ".tif" in x)`` (:399)."""
    return [os.path.join(raw_location, x) for x in sorted(x for x in os.listdir(raw_location) if ".tif" in x)]


def get_real_size(raw_folder):
    """downsample_and_mask.py:25-30 (z = number of entries containing ".tif"; y, x from a plane's header)."""
    planes = list_planes(raw_folder)
    if not planes:
        raise FileNotFoundError(f"no .tif planes in {raw_folder}")
    y, x, _, _ = tiff_info(planes[0])
    return (len(planes), y, x)


def padded_shape(shape, crop_size):
    """downsample_and_mask.py:391-393 / inference.py:229-231."""
    return tuple(int(math.ceil(d / c) * c) for d, c in zip(shape, crop_size))


def load_masked_volume(ctx, raw_location, crop_size, threshold=None, mask=None, z_range=None, nthreads=0):
    """-> (volume, shape_real): device uint16 tensor (Zp, Yp, Xp) holding planes ``z_range`` (default all) of the stack.

    ``threshold``: simple-threshold mode (settings["mask_detection"]["simple_threshold_value"]); ``mask``: uint8
    (Z, Y, X) array / tensor, Ilastik mode (``img *= mask_us[i]``); neither: unmasked.  With ``z_range=(z0, z1)`` only
    that slab is loaded (z-sharded runs: every rank reads its own planes) and no z padding is added.
    """
    import torch
    planes = list_planes(raw_location)
    Z = len(planes)
    _, Y, X = get_real_size(raw_location)
    z0, z1 = (0, Z) if z_range is None else (int(z_range[0]), int(z_range[1]))
    PZ, PY, PX = padded_shape((Z, Y, X), crop_size)
    nz = (PZ if z_range is None else z1 - z0)
    dev = torch.device("cuda", ctx.device)
    stream = torch.cuda.ExternalStream(ctx._L.dlv_stream(ctx._h), device=dev)
    with torch.cuda.stream(stream):
        vol = torch.zeros((nz, PY, PX), dtype=torch.uint16, device=dev)
        m = None
        if mask is not None:
            m = torch.as_tensor(np.ascontiguousarray(mask[z0:z1]) if isinstance(mask, np.ndarray) else mask[z0:z1])
            m = m.to(device=dev, dtype=torch.uint8).contiguous()
    stream.synchronize()
    r1 = min(z1, Z)
    if r1 > z0:
        ctx.load_tiff_planes(planes[z0:r1], Y, X, vol[: r1 - z0], threshold=-1 if threshold is None else int(threshold),
                             mask=m, nthreads=nthreads)
    return vol, (Z, Y, X)
