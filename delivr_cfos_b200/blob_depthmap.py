"""Drop-in for ``depth_map_blobs`` of the reference's ``blob_depthmap.py`` (:114-222) - SURVEY.md section 8, row f3.

Colours every blob with the distance of its centroid from the sample surface (Euclidean distance transform of the
down-sampled mask, :174-181, as ``dlv_edt``) and writes
``<out>/<brain>/<brain>_depthmap_tiffs/depthmap_NNNN.tif`` (uint16, LZW).  The painting loop (:198-207) runs as
``dlv_paint_boxes``; connected components as ``dlv_ccl``.

The reference function cannot run as shipped: it indexes the 3-D memmap with four indices (:137) and uses ``N``
without defining it when cached statistics exist (:198).  This mirror does what the loop evidently means -
3-D volume, ``N`` = number of components - and keeps the loop's other properties: ``range(N)`` starts at the
background row 0 (its box is the whole volume, so every blob first receives the background centroid's depth and
is then re-coloured by its own box) and never reaches the last component; pad_bb grows the stats rows in place.
There is no golden for it (the reference raises); tests compare with the restated loop (oracle/paint_ref.py).
"""
import datetime
import os
import pickle
import shutil

import numpy as np

from .blob_highlighter import _write_planes, padded_boxes
from .count_blobs import _context, load_cached_stats
from .slabs import ccl_any_size


def blob_depths(stats, distances, settings, n=None):
    """blob_depthmap.py:184-207: centroid -> down-sampled voxel -> distance value, for the rows the painting loop
    reads (0..N-1: the reference never touches the last component).  The background row's centroid may be NaN (no
    background voxel): such rows read voxel (0, 0, 0) instead of a garbage index; indices are clipped to the stack."""
    ds = settings["mask_detection"]["downsample_steps"]
    coordinates = np.array(stats["centroids"], dtype=np.float64)[:n].copy()
    coordinates[~np.isfinite(coordinates).all(axis=1)] = 0.0
    coordinates[:, 0] = coordinates[:, 0] / (ds["downsample_um_z"] / ds["original_um_z"])
    coordinates[:, 1] = coordinates[:, 1] / (ds["downsample_um_y"] / ds["original_um_y"])
    coordinates[:, 2] = coordinates[:, 2] / (ds["downsample_um_x"] / ds["original_um_x"])
    coordinates = np.clip(coordinates.astype(int), 0, np.array(distances.shape) - 1)
    return distances[coordinates[:, 0], coordinates[:, 1], coordinates[:, 2]]


def depth_map_blobs(settings, brain, stack_shape, device=0):
    import cv2
    vis = settings["visualization"]
    path_binary, path_out = vis["input_prediction_location"], vis["output_location"]
    path_out_depthmap = os.path.join(path_out, brain, brain + "_depthmap_tiffs")
    path_cache = os.path.join(vis["cache_location"], brain)
    os.makedirs(path_out_depthmap, exist_ok=True)
    os.makedirs(path_cache, exist_ok=True)
    path_brain_binary = path_binary + [x for x in os.listdir(path_binary) if brain in x][0] + "/binary_segmentations/binaries.npy"
    print(f"{datetime.datetime.now()} : Loading brain")
    shape = tuple(int(s) for s in stack_shape[2:])
    mask = np.ascontiguousarray(np.memmap(path_brain_binary, dtype=np.uint8, mode="r", shape=shape, offset=128))
    ctx = _context(device)
    print(f"{datetime.datetime.now()} : calculating connected-component analysis")
    cached = load_cached_stats(settings, brain)
    if not cached:
        table = ccl_any_size(ctx, mask, shape)
        stats = {"voxel_counts": table["voxel_counts"], "bounding_boxes": np.array(table["bounding_boxes"]),
                 "centroids": table["centroids"]}
    else:
        print(f"Found stats at {cached}")
        with open(cached, "rb") as file:
            stats = pickle.load(file)
    N = len(stats["voxel_counts"]) - 1

    print(f"{datetime.datetime.now()} : calculating euclidean distance transform")
    ds = settings["mask_detection"]["downsample_steps"]
    stack_path = os.path.join(settings["mask_detection"]["output_location"], brain, "downsampled_masked_stack.tif")
    ok, pages = cv2.imreadmulti(stack_path, flags=cv2.IMREAD_UNCHANGED)
    if not ok or not pages:
        raise IOError(f"cannot read {stack_path}")
    # the reference pads the stack with zeros, transforms, and crops again (:171-178): dlv_edt treats the outside as zero
    distances = ctx.edt(np.ascontiguousarray(np.stack(pages) != 0).view(np.uint8),
                        (ds["downsample_um_z"], ds["downsample_um_y"], ds["downsample_um_x"]))
    distances = distances.astype(np.uint16)

    print(f"{datetime.datetime.now()} : generating depth-coded blob map")
    depths = blob_depths(stats, distances, settings, n=N)
    ids = np.arange(N)
    boxes = padded_boxes(stats, ids, stack_shape)
    depthmap = np.empty(shape, dtype=np.uint16)
    ctx.paint_boxes(mask, shape, boxes, depths[:N].astype(np.int64), [depthmap])
    print(f"{datetime.datetime.now()} : exporting depth-coded tiffs")
    _write_planes(os.path.join(path_out_depthmap, "depthmap_{z}.tif"), depthmap)
    print(f"{datetime.datetime.now()} : Cleanup")
    shutil.rmtree(path_cache, ignore_errors=True)
