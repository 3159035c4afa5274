"""Shim for connected-components-3d 3.12.3 (requirements.txt:9; count_blobs.py:4,61,64,85).

STAND-IN: cc3d itself is absent.  ``connected_components`` is restated with
scipy.ndimage.label (3x3x3 structure == 26-connectivity); scipy numbers
components in order of their first voxel in C-order raster scan, which is
cc3d's published renumbering rule.  ``statistics`` restates cc3d's arrays:
index 0 is the background, centroids are float64 sum(coord)/count in array
axis order, bounding boxes are inclusive [a0min, a0max, a1min, a1max, a2min,
a2max].  Output dtypes are cc3d's choice and are not contractual here.
"""
import numpy as np
from scipy import ndimage


def connected_components(data, connectivity=26, return_N=False, out_file=None, **_):
    if connectivity != 26:
        raise NotImplementedError("shim restates the default connectivity only")
    lab, n = ndimage.label(np.asarray(data) != 0, structure=np.ones((3, 3, 3), dtype=bool))
    lab = lab.astype(np.uint32)
    if out_file is not None:
        mm = np.lib.format.open_memmap(out_file, mode="w+", dtype=lab.dtype, shape=lab.shape)
        mm[...] = lab
        lab = mm
    return (lab, n) if return_N else lab


def statistics(labels, no_slice_conversion=False):
    labels = np.asarray(labels)
    n = int(labels.max()) if labels.size else 0
    flat = labels.reshape(-1).astype(np.int64)
    counts = np.bincount(flat, minlength=n + 1).astype(np.uint64)
    cent = np.full((n + 1, 3), np.nan, dtype=np.float64)
    bbox = np.zeros((n + 1, 6), dtype=np.int64)
    coords = np.unravel_index(np.arange(flat.size, dtype=np.int64), labels.shape)
    for ax in range(3):
        c = coords[ax]
        s = np.bincount(flat, weights=c.astype(np.float64), minlength=n + 1)
        with np.errstate(invalid="ignore", divide="ignore"):
            cent[:, ax] = s / counts
        mn = np.full(n + 1, labels.shape[ax], dtype=np.int64)
        mx = np.full(n + 1, -1, dtype=np.int64)
        np.minimum.at(mn, flat, c)
        np.maximum.at(mx, flat, c)
        bbox[:, 2 * ax] = mn
        bbox[:, 2 * ax + 1] = mx
    if not no_slice_conversion:
        bbox = [tuple(slice(int(b[2 * a]), int(b[2 * a + 1]) + 1) for a in range(3)) for b in bbox]
    return {"voxel_counts": counts, "bounding_boxes": bbox, "centroids": cent}
