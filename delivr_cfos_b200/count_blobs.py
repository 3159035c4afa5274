"""Drop-in for the reference's ``count_blobs.py`` (count_blobs.py:10-118).

Same signature, same cache probes, same three output files:
``{brain}-{N}-cc3d.npy`` (labels), ``{brain}-stats.pickle`` (dict with
``voxel_counts`` / ``bounding_boxes`` / ``centroids``, rows 0..N) and
``(Z, Y, X)_{brain}.csv`` with rows for labels 1..N-1 (the reference's
``range(1, N)`` loop, count_blobs.py:104, drops the last component - kept).
Labelling and statistics run in ``dlv_ccl`` (CUDA); the O(N^2) pandas
concat loop (count_blobs.py:101-110) is replaced by direct text emission of
the byte-identical CSV.
"""
import datetime
import os
import pickle

import numpy as np

from ._lib import Context

_CTX = {}


def _context(device=0):
    if device not in _CTX:
        _CTX[device] = Context(device)
    return _CTX[device]


def load_cached_brain(settings, brain):
    """count_blobs.py:10-21."""
    path_in = settings["postprocessing"]["output_location"]
    result = False
    for item in [x for x in os.listdir(path_in) if ".npy" in x]:
        if brain in item:
            result = os.path.join(path_in, item)
    return result


def load_cached_stats(settings, brain):
    """count_blobs.py:23-34."""
    path_in = settings["postprocessing"]["output_location"]
    result = False
    for item in [x for x in os.listdir(path_in) if ".pickle" in x]:
        if brain in item:
            result = os.path.join(path_in, item)
    return result


def csv_text(stats, n):
    """Exact text pandas writes for the reference's DataFrame (count_blobs.py:101-114).

    Header ``,Blob,Coords,Size``; one row per label 1..N-1: index column always 0,
    Coords = str(list of python floats) quoted because it contains commas.
    """
    cent = np.asarray(stats["centroids"])[1:n].tolist()          # python floats: str(list) prints their repr
    cnt = np.asarray(stats["voxel_counts"])[1:n].tolist()
    out = [",Blob,Coords,Size\n"]
    out.extend(f'0,{i},"{c}",{k}\n' for i, (c, k) in enumerate(zip(cent, cnt), 1))
    return "".join(out)


def statistics_from_labels(labels):
    """cc3d.statistics equivalent for a cached label volume (count_blobs.py:71-76,85): exact host reduction, one
    stable sort of the label volume and segment reductions (no per-label Python loop)."""
    lab = np.asarray(labels)
    n = int(lab.max()) if lab.size else 0
    flat = lab.reshape(-1).astype(np.int64)
    counts = np.bincount(flat, minlength=n + 1).astype(np.uint64)
    sums = np.zeros((n + 1, 3), dtype=np.uint64)
    bbox = np.zeros((n + 1, 6), dtype=np.int64)
    order = np.argsort(flat, kind="stable")
    bounds = np.concatenate([[0], np.cumsum(counts.astype(np.int64))])
    present = counts > 0
    starts = bounds[:-1][present]                       # reduceat needs non-empty, increasing segments
    coords = np.unravel_index(order, lab.shape)          # coordinates in label-sorted order
    for ax in range(3):
        c = coords[ax].astype(np.int64)
        bbox[:, 2 * ax], bbox[:, 2 * ax + 1] = lab.shape[ax], -1          # labels without voxels: neutral box
        if len(starts):
            sums[present, ax] = np.add.reduceat(c.astype(np.uint64), starts)
            bbox[present, 2 * ax] = np.minimum.reduceat(c, starts)
            bbox[present, 2 * ax + 1] = np.maximum.reduceat(c, starts)
    with np.errstate(invalid="ignore", divide="ignore"):
        cent = sums.astype(np.float64) / counts.astype(np.float64)[:, None]
    return {"voxel_counts": counts, "bounding_boxes": bbox, "centroids": cent}


def count_blobs(settings, path_in, brain_i, brain, stack_shape, min_size=-1, max_size=-1, device=0):
    """Same contract as the reference's count_blobs (count_blobs.py:36-118)."""
    path_out = settings["postprocessing"]["output_location"]
    if not os.path.exists(path_out):
        os.mkdir(path_out)

    len_b = len(os.listdir(path_in))
    start = datetime.datetime.now()
    print(f"{start} Now postprocessing inference for {brain} - {brain_i}/{len_b}")
    brain_path = os.path.join(path_in, brain, "binary_segmentations", "binaries.npy")
    bin_img = np.memmap(brain_path, dtype=np.uint8, mode="r", shape=tuple(stack_shape[2:]), offset=128)
    mid = datetime.datetime.now()
    print(f"{mid} Reading took {mid - start}")

    stats = None
    cached_brain = load_cached_brain(settings, brain)
    if not cached_brain:
        print("No cached brain found, performing connected components on the GPU...")
        labels = np.empty(bin_img.shape, dtype=np.uint32)
        from .slabs import ccl_any_size          # whole-brain volumes exceed one 32-bit label space: z sub-slabs, exact merge
        table = ccl_any_size(_context(device), np.ascontiguousarray(bin_img), bin_img.shape, labels_out=labels)
        N = table["n"]
        np.save(os.path.join(path_out, f"{brain}-{N}-cc3d.npy"), labels)
        stats = {k: table[k] for k in ("voxel_counts", "bounding_boxes", "centroids")}
    else:
        N = int(cached_brain.split("/")[-1].split("-")[1])
        print(f"Cached brain found at {cached_brain} with {N} components, loading...")
        labels = np.load(cached_brain)
    mid3 = datetime.datetime.now()
    print(f"{mid3} cc3d+writing/loading took {mid3 - mid} : {N}")

    cached_stats = load_cached_stats(settings, brain)
    if not cached_stats:
        if stats is None:
            stats = statistics_from_labels(labels)
        path_stats = os.path.join(path_out, f"{brain}-stats.pickle")
        with open(path_stats, "wb") as file:
            pickle.dump(stats, file, protocol=pickle.HIGHEST_PROTOCOL)
    else:
        print(f"Found stats at {cached_stats}")
        with open(cached_stats, "rb") as file:
            stats = pickle.load(file)
    mid4 = datetime.datetime.now()
    print(f"{mid4} stats took {mid4 - mid3}")

    output_name = f"{bin_img.shape}_{brain.replace('.nii.gz', '')}.csv"
    with open(path_out + output_name, "w") as f:       # no separator, like the reference (count_blobs.py:114)
        f.write(csv_text(stats, N))
    end = datetime.datetime.now()
    end_delta = end - start
    remaining_time = (len_b - brain_i) * end_delta
    print(f"{end} {brain} {brain_i} / {len_b} Done; Took {end_delta}, ETA {remaining_time}")
