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
    Coords = str(list of python floats) quoted because it contains commas.  Formatted by the library's host code
    (dlv_table_csv, all host threads): the 2.5 M rows of a whole brain took a Python loop 6.7 s - longer than the
    segmentation of the brain on 8 GPUs - and take 1.0 s this way (8 cores).
    """
    from ._lib import table_csv
    return table_csv(np.asarray(stats["centroids"]), np.asarray(stats["voxel_counts"]), n)


def _chunk_statistics(lab):
    """Exact statistics of one z-chunk of a label volume (local z coordinates), rows 0..max label: counts / coordinate
    sums / boxes from the foreground voxels only (one stable sort of the non-zero voxels), background row in closed form."""
    Z, Y, X = lab.shape
    flat = lab.reshape(-1)
    nz = np.flatnonzero(flat)
    l = flat[nz].astype(np.int64)
    n = int(l.max()) if len(l) else 0
    counts = np.bincount(l, minlength=n + 1).astype(np.uint64)
    sums = np.zeros((n + 1, 3), dtype=np.uint64)
    bbox = np.empty((n + 1, 6), dtype=np.int64)
    bbox[:, 0::2] = (Z, Y, X)                       # labels without voxels here: neutral box
    bbox[:, 1::2] = -1
    order = np.argsort(l, kind="stable")
    coords = np.unravel_index(nz[order], lab.shape)
    present = counts > 0
    starts = np.concatenate([[0], np.cumsum(counts.astype(np.int64))])[:-1][present]
    for ax in range(3):
        c = coords[ax].astype(np.int64)
        if len(starts):
            sums[present, ax] = np.add.reduceat(c.astype(np.uint64), starts)
            bbox[present, 2 * ax] = np.minimum.reduceat(c, starts)
            bbox[present, 2 * ax + 1] = np.maximum.reduceat(c, starts)
    # background (row 0, which cc3d.statistics reports too): everything minus the foreground
    bg = flat.size - len(nz)
    counts[0] = bg
    dims = (Z, Y, X)
    zero = lab == 0
    for ax in range(3):
        total = (dims[ax] * (dims[ax] - 1) // 2) * (flat.size // dims[ax])
        sums[0, ax] = total - int(sums[1:, ax].sum())
        other = tuple(a for a in range(3) if a != ax)
        idx = np.flatnonzero(zero.any(axis=other))
        bbox[0, 2 * ax], bbox[0, 2 * ax + 1] = (int(idx[0]), int(idx[-1])) if len(idx) else (dims[ax], -1)
    return {"n": n, "voxel_counts": counts, "sums": sums, "bounding_boxes": bbox}


def statistics_from_labels(labels, chunk_voxels=1 << 28):
    """cc3d.statistics equivalent for a cached label volume (count_blobs.py:71-76,85): one streaming pass over
    z-chunks (bounded host memory: the cached file may be a whole-brain memmap), per-chunk exact integer tables merged
    associatively by dlv_table_merge."""
    from ._lib import table_merge
    lab = np.asarray(labels) if not isinstance(labels, np.memmap) else labels
    Z, Y, X = lab.shape
    zs = max(1, int(chunk_voxels) // max(1, Y * X))
    tables, z0s = [], []
    for z0 in range(0, Z, zs):
        tables.append(_chunk_statistics(np.ascontiguousarray(lab[z0:z0 + zs])))
        z0s.append(z0)
    n = max(t["n"] for t in tables) if tables else 0
    luts = [np.arange(t["n"] + 1, dtype=np.uint32) for t in tables]
    out = table_merge(tables, luts, z0s, n, (Z, Y, X))
    return {k: out[k] for k in ("voxel_counts", "bounding_boxes", "centroids")}


def _dist():
    """(rank, world, torch.distributed) when this process is one rank of an initialised process group, else (0, 1, None)."""
    try:
        import torch.distributed as dist
    except ImportError:
        return 0, 1, None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size(), dist
    return 0, 1, None


def _local_device():
    return int(os.environ.get("LOCAL_RANK", 0))


def _cuda_label_slab(device, binaries_np):
    """Product factory of the labelling-stage object of one rank (tests substitute an oracle-backed one)."""
    import torch
    from .slabs import CudaLabelSlab
    ctx = _context(device)
    pinned = torch.from_numpy(np.ascontiguousarray(binaries_np))
    return CudaLabelSlab(ctx, pinned.to(torch.device("cuda", device)))


_LABEL_SLAB_FACTORY = _cuda_label_slab


def _labels_to_numpy(lab):
    return (lab.cpu().numpy() if hasattr(lab, "cpu") else np.asarray(lab)).view(np.uint32)


def _label_and_count(bin_img, labels_path, device):
    """Single process: table first (labels never leave the device), then - only if ``labels_path`` - a second
    labelling pass that streams the globally numbered labels into the .npy file sub-slab by sub-slab."""
    from .slabs import ccl_any_size          # whole-brain volumes exceed one 32-bit label space: z sub-slabs, exact merge
    ctx = _context(device)
    table = ccl_any_size(ctx, bin_img, bin_img.shape)
    if labels_path:
        out = np.lib.format.open_memmap(labels_path, mode="w+", dtype=np.uint32, shape=tuple(bin_img.shape))

        def sink(z0, z1, lab):
            out[z0:z1] = _labels_to_numpy(lab)

        ccl_any_size(ctx, bin_img, bin_img.shape, labels_sink=sink)
        out.flush()
        del out
    return table


def _label_and_count_distributed(bin_img, labels_path, rank, world, dist, device):
    """One rank per GPU (torchrun): every rank labels an equal share of the planes read from binaries.npy, one label
    plane per seam and two padded all-gathers make the numbering and the table global (slabs.distributed_label)."""
    import torch
    from .slabs import TorchComm, distributed_label
    Z = bin_img.shape[0]
    z0, z1 = Z * rank // world, Z * (rank + 1) // world
    slab = _LABEL_SLAB_FACTORY(device, bin_img[z0:z1])
    comm = TorchComm(torch.device("cuda", device) if dist.get_backend() == "nccl" else None)
    slab.ccl()
    table = distributed_label(slab, comm, tuple(bin_img.shape))
    if labels_path:
        if rank == 0:
            out = np.lib.format.open_memmap(labels_path, mode="w+", dtype=np.uint32, shape=tuple(bin_img.shape))
            del out
        comm.barrier()
        out = np.load(labels_path, mmap_mode="r+")
        if z1 > z0:
            out[z0:z1] = _labels_to_numpy(slab.labels)
        out.flush()
        del out
        comm.barrier()
    return table


def count_blobs(settings, path_in, brain_i, brain, stack_shape, min_size=-1, max_size=-1, device=None):
    """Same contract as the reference's count_blobs (count_blobs.py:36-118).

    Extensions (no effect on the reference's keys): under ``torchrun`` (an initialised process group) the ranks share
    the labelling and rank 0 writes the files; ``settings["FLAGS"]["SAVE_CC3D_LABELS"] = False`` skips the
    ``{brain}-{N}-cc3d.npy`` dump (4 B/voxel - the file is only this function's own cache, count_blobs.py:10-21,65)."""
    path_out = settings["postprocessing"]["output_location"]
    rank, world, dist = _dist()
    device = _local_device() if device is None else device
    if not os.path.exists(path_out):
        os.makedirs(path_out, exist_ok=True)

    len_b = len(os.listdir(path_in))
    start = datetime.datetime.now()
    print(f"{start} Now postprocessing inference for {brain} - {brain_i}/{len_b}")
    brain_path = os.path.join(path_in, brain, "binary_segmentations", "binaries.npy")
    bin_img = np.memmap(brain_path, dtype=np.uint8, mode="r", shape=tuple(stack_shape[2:]), offset=128)
    mid = datetime.datetime.now()
    print(f"{mid} Reading took {mid - start}")

    stats = None
    labels = None
    cached_brain = load_cached_brain(settings, brain)
    cached_stats = load_cached_stats(settings, brain)
    if dist is not None:
        dist.barrier()                       # every rank has probed the caches before any rank writes a file
    if not cached_brain:
        print("No cached brain found, performing connected components on the GPU...")
        save_labels = bool(settings.get("FLAGS", {}).get("SAVE_CC3D_LABELS", True))
        # the reference's temporary name for the label store (count_blobs.py:55); renamed once N is known
        tmp = os.path.join(path_out, brain + "temp_cc3d_store.npy") if save_labels else None
        if dist is None:
            table = _label_and_count(bin_img, tmp, device)
        else:
            table = _label_and_count_distributed(bin_img, tmp, rank, world, dist, device)
        N = table["n"]
        if tmp and rank == 0:
            os.replace(tmp, os.path.join(path_out, f"{brain}-{N}-cc3d.npy"))
        stats = {k: table[k] for k in ("voxel_counts", "bounding_boxes", "centroids")}
    else:
        N = int(cached_brain.split("/")[-1].split("-")[1])
        print(f"Cached brain found at {cached_brain} with {N} components, loading...")
        labels = np.load(cached_brain, mmap_mode="r")
    mid3 = datetime.datetime.now()
    print(f"{mid3} cc3d+writing/loading took {mid3 - mid} : {N}")

    if not cached_stats:
        if stats is None:
            stats = statistics_from_labels(labels)
        if rank == 0:
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
    if rank == 0:
        with open(path_out + output_name, "w") as f:       # no separator, like the reference (count_blobs.py:114)
            f.write(csv_text(stats, N))
    if dist is not None:
        dist.barrier()
    end = datetime.datetime.now()
    end_delta = end - start
    remaining_time = (len_b - brain_i) * end_delta
    print(f"{end} {brain} {brain_i} / {len_b} Done; Took {end_delta}, ETA {remaining_time}")
