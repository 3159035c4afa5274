"""Drop-in for the reference's ``blob_highlighter.py`` (:17-170) - SURVEY.md section 8, row f3.

Same signature, config keys (``visualization.*``, ``postprocessing.output_location``, ``FLAGS.LOAD_ALL_RAM``) and
output files (``<out>/<brain>_rgb_tiffs/<brain>rgb_C0{0,1,2}_zNNNN.tif`` uint8 and
``<out>/<brain>/<brain>_region_id_tiffs/region_id_NNNN.tif`` uint16, LZW; written by ``dlv_tiff_write_planes`` on all
host threads - same pixels, not the same bytes as tifffile's).  The per-cell Python loops that paint
every blob through its bounding box (blob_highlighter.py:107-124, :143-151) run as ``dlv_paint_boxes`` on the GPU;
connected components, when no cached statistics exist (:81-90), run as ``dlv_ccl`` instead of a second cc3d pass.

Kept from the reference on purpose (tests/golden/p1_highlight.npz pins them):
* ``pad_bb`` grows the bounding-box rows of ``stats`` IN PLACE, so the region-id pass (and a repeated
  ``connected_component_id``) sees boxes that were already grown;
* a box re-colours every foreground voxel inside it - where boxes overlap, the last cell in CSV order wins.
A ``connected_component_id`` that occurs more than once makes the reference's ``bin * colours`` a shape mismatch:
the RGB pass counts an error and skips the cell (:121-124; a box exactly two voxels wide in x would broadcast
instead - not reproduced), the region-id pass raises ValueError (:151) - same here.
"""
import datetime
import os
import pickle
import shutil

import numpy as np

from .count_blobs import _context, load_cached_stats
from .slabs import ccl_any_size


def pad_bb(bb, stack_shape):
    """blob_highlighter.py:17-22 (in place)."""
    if bb[1] < stack_shape[2]:
        bb[1] += 1
    if bb[3] < stack_shape[3]:
        bb[3] += 1
    if bb[5] < stack_shape[4]:
        bb[5] += 1
    return bb


def padded_boxes(stats, cc_ids, stack_shape):
    """The boxes the reference's loop would slice with, in order, mutating ``stats['bounding_boxes']`` like pad_bb
    does.  Vectorised when every id occurs once; a repeated id needs the sequential rule."""
    bbs = stats["bounding_boxes"]
    ids = np.asarray(cc_ids, dtype=np.int64)
    if len(np.unique(ids)) == len(ids):
        rows = bbs[ids]
        for col, dim in ((1, 2), (3, 3), (5, 4)):
            rows[:, col] += rows[:, col] < stack_shape[dim]
        bbs[ids] = rows
        return rows.astype(np.int64)
    out = np.empty((len(ids), 6), dtype=np.int64)
    for k, i in enumerate(ids):
        out[k] = pad_bb(bbs[i], stack_shape)
    return out


def _write_planes(fmt, vol):
    """One LZW TIFF per z plane (blob_highlighter.py:127-133: tifffile.imwrite(..., compression='lzw') in a loop),
    compressed on all host threads by the library's writer."""
    from ._lib import tiff_write_planes
    tiff_write_planes([fmt.format(z=str(z).zfill(4)) for z in range(vol.shape[0])], vol, compression=5)


def blob_highlighter(settings, brain_item, stack_shape, device=0):
    """Colour blobs by their atlas region (blob_highlighter.py:38-170)."""
    import pandas as pd
    brain, highlight_area = brain_item[0], brain_item[1]
    if highlight_area != "":
        print(f"{datetime.datetime.now()} Highlighting {highlight_area} in {brain}")
    else:
        print(f"{datetime.datetime.now()} Highlighting everything in {brain}")
    vis = settings["visualization"]
    path_binary, path_cell_csv = vis["input_prediction_location"], vis["input_csv_location"]
    path_out, path_cache = vis["output_location"], vis["cache_location"]
    path_out_rgb = os.path.join(path_out, brain + "_rgb_tiffs")
    path_cache = os.path.join(path_cache, brain)
    os.makedirs(path_out_rgb, exist_ok=True)
    os.makedirs(path_cache, exist_ok=True)
    path_brain_binary = path_binary + [x for x in os.listdir(path_binary) if brain in x][0] + "/binary_segmentations/binaries.npy"

    if not vis["no_atlas_depthmap"]:
        path_brain_cell_csv = path_cell_csv + [x for x in os.listdir(path_cell_csv) if "cells_" + brain in x and ".csv" in x][0]
        print(path_brain_cell_csv)
        print(f"{datetime.datetime.now()} : Loading csv")
        cell_csv = pd.read_csv(path_brain_cell_csv, index_col=0)
        cell_csv = cell_csv.loc[cell_csv["acronym"] != "bgr"]

    print(f"{datetime.datetime.now()} : Loading brain")
    shape = tuple(int(s) for s in stack_shape[2:])
    bin_img = np.memmap(path_brain_binary, dtype=np.uint8, mode="r", shape=shape, offset=128)
    ctx = _context(device)
    cached = load_cached_stats(settings, brain)
    if not cached:
        table = ccl_any_size(ctx, np.ascontiguousarray(bin_img), shape)
        stats = {"voxel_counts": table["voxel_counts"], "bounding_boxes": np.array(table["bounding_boxes"]),
                 "centroids": table["centroids"]}
    else:
        print(f"Found stats at {cached}")
        with open(cached, "rb") as file:
            stats = pickle.load(file)
    mask = np.ascontiguousarray(bin_img)

    if vis["region_id_rgb"]:
        print(f"{datetime.datetime.now()} : coloring blobs")
        ids = cell_csv["connected_component_id"].to_numpy()
        boxes = padded_boxes(stats, ids, stack_shape)
        dup = cell_csv["connected_component_id"].duplicated(keep=False).to_numpy()
        for e, cc_id in enumerate(ids[dup], 1):
            print("error number ", e, " at cc_id ", cc_id)
        vals = cell_csv[["red", "green", "blue"]].to_numpy()[~dup]
        rgb = [np.empty(shape, dtype=np.uint8) for _ in range(3)]
        ctx.paint_boxes(mask, shape, boxes[~dup], vals, rgb)
        print(f"{datetime.datetime.now()} : Generating RGB tiffs")
        for c, vol in enumerate(rgb):
            _write_planes(os.path.join(path_out_rgb, brain + f"rgb_C0{c}_z" + "{z}.tif"), vol)

    print(f"{datetime.datetime.now()} : Generating region_id gray-value tiffs")
    if vis["region_id_grayvalues"]:
        path_out_region_id = os.path.join(path_out, brain, brain + "_region_id_tiffs")
        os.makedirs(path_out_region_id, exist_ok=True)
        ids = cell_csv["connected_component_id"].to_numpy()
        if cell_csv["connected_component_id"].duplicated().any():
            raise ValueError("operands could not be broadcast together: connected_component_id occurs more than once "
                             "(blob_highlighter.py:151)")
        boxes = padded_boxes(stats, ids, stack_shape)
        region = np.empty(shape, dtype=np.uint16)
        ctx.paint_boxes(mask, shape, boxes, cell_csv["graph_order"].to_numpy(), [region])
        _write_planes(os.path.join(path_out_region_id, "region_id_{z}.tif"), region)

    if vis["no_atlas_depthmap"]:
        from .blob_depthmap import depth_map_blobs
        depth_map_blobs(settings, brain, stack_shape, device=device)

    print(f"{datetime.datetime.now()} : Cleanup")
    shutil.rmtree(path_cache, ignore_errors=True)
