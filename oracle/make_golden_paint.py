#!/usr/bin/env python
"""Generate tests/golden/p1_highlight.npz by running the UNMODIFIED reference blob_highlighter (TEST INFRASTRUCTURE).

Runs only in the authoring container (needs /root/reference):

    python oracle/make_golden_paint.py

``oracle/shims`` supplies the third-party names blob_highlighter.py imports that are not installed here
(tifffile -> OpenCV libtiff, import-only stubs for skimage.morphology/draw/io, matplotlib, nibabel; cc3d ->
scipy-backed stand-in).  The reference's own ``count_blobs`` produces the statistics pickle the highlighter
loads (blob_highlighter.py:91-95), so the bounding boxes come from the reference's call sites too.

The volume holds small blobs plus one long diagonal blob whose bounding box covers several others: the
reference re-colours everything inside a box (blob_highlighter.py:112-113 says so), which is exactly the
"last box in CSV order wins" behaviour the GPU painter must reproduce.  Both passes are enabled
(``region_id_rgb`` and ``region_id_grayvalues``), so the second pass sees boxes that pad_bb already grew once.
"""
import os
import pickle
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DLV_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")


def paint_volume(shape, seed):
    rng = np.random.default_rng(seed)
    Z, Y, X = shape
    b = np.zeros(shape, dtype=np.uint8)
    for _ in range(60):
        c = [int(rng.integers(2, s - 2)) for s in shape]
        r = rng.integers(1, 3, size=3)
        b[max(c[0] - r[0], 0):c[0] + r[0], max(c[1] - r[1], 0):c[1] + r[1], max(c[2] - r[2], 0):c[2] + r[2]] = 1
    for i in range(4, min(shape) - 4):                         # long diagonal blob ("blood vessel")
        b[i, i + 3, i + 1] = 1
    b[Z - 1, Y - 1, X - 1] = 1                                   # a blob on the far corner: pad_bb's border rule
    b[0, 0, 0:3] = 1
    return b


def main():
    sys.path[:0] = [os.path.join(HERE, "shims"), REF, ROOT]
    import cv2
    import pandas as pd
    import blob_highlighter as BH          # the reference's file, unmodified
    import count_blobs as CB               # the reference's file, unmodified

    brain, shape, seed = "brainP", (40, 60, 50), 21
    stack_shape = (1, 1) + shape
    b = paint_volume(shape, seed)
    tmp = tempfile.mkdtemp(prefix="dlv_gold_paint_")
    d = {k: os.path.join(tmp, k) + "/" for k in ("pred", "post", "csv", "cache", "out")}
    for v in d.values():
        os.makedirs(v)
    os.makedirs(os.path.join(d["pred"], brain, "binary_segmentations"))
    mm = np.lib.format.open_memmap(os.path.join(d["pred"], brain, "binary_segmentations", "binaries.npy"), mode="w+",
                                   dtype=np.uint8, shape=shape)
    mm[...] = b
    mm.flush()
    del mm
    settings = {"postprocessing": {"output_location": d["post"]},
                "visualization": {"input_prediction_location": d["pred"], "input_csv_location": d["csv"],
                                  "output_location": d["out"], "cache_location": d["cache"], "region_id_rgb": True,
                                  "region_id_grayvalues": True, "no_atlas_depthmap": False},
                "FLAGS": {"LOAD_ALL_RAM": True}}
    CB.count_blobs(settings, d["pred"], 1, brain, stack_shape, -1, -1)
    with open(os.path.join(d["post"], f"{brain}-stats.pickle"), "rb") as f:
        stats = pickle.load(f)
    n = len(stats["voxel_counts"]) - 1

    rng = np.random.default_rng(seed + 1)
    ids = rng.permutation(np.arange(1, n + 1))[: n - 3]        # CSV order is NOT label order; three cells are missing
    df = pd.DataFrame({"connected_component_id": ids,
                       "acronym": np.where(rng.random(len(ids)) < 0.15, "bgr", "CTX"),
                       "red": rng.integers(1, 256, len(ids)), "green": rng.integers(0, 256, len(ids)),
                       "blue": rng.integers(0, 256, len(ids)), "graph_order": rng.integers(1, 1300, len(ids))})
    csv_path = os.path.join(d["csv"], f"cells_{brain}.csv")
    df.to_csv(csv_path)
    bbox0 = np.array(stats["bounding_boxes"]).copy()

    BH.blob_highlighter(settings, (brain, ""), stack_shape)

    def planes(fmt, dtype):
        out = np.zeros(shape, dtype=dtype)
        for z in range(shape[0]):
            p = cv2.imread(fmt.format(z=str(z).zfill(4)), cv2.IMREAD_UNCHANGED)
            assert p is not None and p.dtype == dtype, (fmt, z)
            out[z] = p
        return out

    rgb_dir = os.path.join(d["out"], brain + "_rgb_tiffs")
    rgb = [planes(os.path.join(rgb_dir, brain + f"rgb_C0{c}_z" + "{z}.tif"), np.uint8) for c in range(3)]
    reg = planes(os.path.join(d["out"], brain, brain + "_region_id_tiffs", "region_id_{z}.tif"), np.uint16)
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "p1_highlight.npz"), shape=np.array(shape), bits=np.packbits(b),
                        bounding_boxes=bbox0, voxel_counts=np.array(stats["voxel_counts"]),
                        centroids=np.array(stats["centroids"]), csv=np.frombuffer(open(csv_path, "rb").read(), dtype=np.uint8),
                        red=rgb[0], green=rgb[1], blue=rgb[2], region=reg,
                        files=np.array(sorted(os.path.relpath(os.path.join(r, f), d["out"]) for r, _, fs in os.walk(d["out"]) for f in fs)))
    print("painted voxels", int((rgb[0] > 0).sum()), "of", int(b.sum()), "foreground; components", n)


if __name__ == "__main__":
    main()
