"""z-slab sharding of the blob_detection hot path over the GPUs of one box (one process per GPU).

The volume's *windows* (z-major order) are partitioned contiguously over the ranks, cut at window granularity so that
every rank runs the same number of active windows.  Each rank gathers and runs its own windows and blends them into
a local fixed-point accumulator; three small exchanges make the result identical to a single-GPU run, bit for bit:

1. logits  - the planes a rank touched beyond the ones it owns are sent down the chain r -> r+1 and added (int32
             adds: exact, order independent); the owner of a plane ends up with the complete sum;
2. activity- the per-window "input max > 0" flags are all-gathered (the averaging needs the skipped windows);
3. labels  - after a per-slab labelling, rank r sends its last label plane to r+1, which lists the 26-adjacent
             label pairs; pairs + per-slab component counts are all-gathered and every rank resolves the same
             global numbering (components ordered by their first voxel in raster order), relabels locally and
             merges the integer statistics tables associatively.

The erosion halo (31 input planes below a slab) is read with the slab, so it needs no collective.

Two drivers share the planning and merge logic: ``run_distributed`` (one rank per process; ``TorchComm`` =
torch.distributed, NCCL on GPUs / gloo in the CPU tests) and ``run_virtual`` (N virtual slabs in one process,
which exercises the same exchange and merge logic on a single GPU).
Compute goes through a ``worker`` object; the product worker is :class:`CudaSlabWorker` (libdelivr_b200.so).
"""
import math

import numpy as np

EROSION_ITERS = 30


# ------------------------------------------------------------------------------------------- planning (pure host)
class SlabPlan:
    """Partition of the windows (z-major order of dense_patch_slices) and the plane ranges that follow from it.

    Rank r runs the contiguous window range ``wrange[r]`` - cut at WINDOW granularity so that every rank gets the
    same share of the (active) windows; a window z-layer may therefore be shared by two or more ranks.  Rank r owns
    the planes from the first plane of the layer its range starts in up to the next rank's first such plane; sums
    for planes beyond its ownership travel down the chain r -> r+1 (-> r+2 ... when a rank owns less than it is sent).
    """

    def __init__(self, shape_real, roi, overlap, world, starts=None, erosion_iters=EROSION_ITERS, layer_weights=None,
                 window_weights=None):
        self.shape_real = tuple(int(s) for s in shape_real)
        self.roi = tuple(int(r) for r in roi)
        self.overlap = float(overlap)
        self.world = int(world)
        self.iters = int(erosion_iters)
        self.shape_pad = tuple(int(math.ceil(d / r) * r) for d, r in zip(self.shape_real, self.roi))   # inference.py:229-231
        if starts is None:
            from ._lib import window_grid
            starts = window_grid(self.shape_pad, self.roi, self.overlap)
        self.sz, self.sy, self.sx = ([int(v) for v in s] for s in starts)
        nz = len(self.sz)
        self.per_layer = len(self.sy) * len(self.sx)
        nwin = nz * self.per_layer
        if window_weights is not None:
            w = np.maximum(np.asarray(window_weights, dtype=np.float64).reshape(-1), 1e-9)
            assert len(w) == nwin
        elif layer_weights is not None:
            w = np.repeat(np.maximum(np.asarray(layer_weights, dtype=np.float64), 1e-9) / self.per_layer, self.per_layer)
        else:
            w = np.ones(nwin)
        cum = np.concatenate([[0.0], np.cumsum(w)])
        k = min(self.world, nwin)               # ranks beyond the window count get empty ranges at the end
        cuts = [0]
        for r in range(1, k):
            c = int(np.searchsorted(cum, cum[-1] * r / k, side="left"))
            cuts.append(min(max(c, cuts[-1] + 1), nwin - (k - r)))
        cuts += [nwin] * (self.world - k + 1)
        self.wrange = [(cuts[r], cuts[r + 1]) for r in range(self.world)]
        # window layers touched by each rank (half-open; neighbouring ranks may share one) - reporting / emptiness
        self.layers = [((c0 // self.per_layer, (c1 - 1) // self.per_layer + 1) if c1 > c0 else (nz, nz)) for c0, c1 in self.wrange]
        self._info = {}

    def rank(self, r):
        """Plane ranges of rank r (global plane numbers, half-open)."""
        if r in self._info:
            return self._info[r]
        c0, c1 = self.wrange[r]
        PZ, Z = self.shape_pad[0], self.shape_real[0]
        rz = self.roi[0]
        if c0 == c1:     # more ranks than windows: nothing to do
            info = dict(layers=self.layers[r], win=(0, 0), own=(0, 0), slab=(0, 0), send=None, recv=None, own_real=(0, 0))
            self._info[r] = info
            return info
        la, lb = self.layers[r][0], self.layers[r][1] - 1
        prv, nxt = self._prev_nonempty(r), self._next_nonempty(r)
        win = (self.sz[la], self.sz[lb] + rz)
        own = (0 if prv is None else self.sz[la], PZ if nxt is None else self.sz[self.layers[nxt][0]])
        recv = None if prv is None else self.rank(prv)["send"]          # what the previous rank touched beyond its planes
        touched_end = max(win[1], recv[1] if recv else 0)
        send = (own[1], touched_end) if (nxt is not None and touched_end > own[1]) else None
        slab = (max(0, min(win[0], own[0] - (self.iters + 1))), max(touched_end, min(PZ, own[1] + self.iters + 1)))
        info = dict(layers=self.layers[r], win=win, own=own, slab=slab, send=send, recv=recv,
                    own_real=(min(own[0], Z), min(own[1], Z)))
        self._info[r] = info
        return info

    def _next_nonempty(self, r):
        for q in range(r + 1, self.world):
            if self.wrange[q][0] < self.wrange[q][1]:
                return q
        return None

    def _prev_nonempty(self, r):
        for q in range(r - 1, -1, -1):
            if self.wrange[q][0] < self.wrange[q][1]:
                return q
        return None

    def windows_of(self, r):
        """int32 [n,3] global origins of rank r's windows, z-major / x fastest (dense_patch_slices order)."""
        c0, c1 = self.wrange[r]
        if c1 <= c0:
            return np.zeros((0, 3), dtype=np.int32)
        idx = np.arange(c0, c1, dtype=np.int64)
        iz, rem = idx // self.per_layer, idx % self.per_layer
        iy, ix = rem // len(self.sx), rem % len(self.sx)
        out = np.stack([np.asarray(self.sz, dtype=np.int32)[iz], np.asarray(self.sy, dtype=np.int32)[iy],
                        np.asarray(self.sx, dtype=np.int32)[ix]], axis=1)
        return np.ascontiguousarray(out, dtype=np.int32).reshape(-1, 3)


def resolve_global_labels(counts, pairs):
    """Global component numbering from per-slab counts and boundary pairs.

    counts[r] = N_r; pairs[r] = uint32 [k,2] of (label in slab r-1, label in slab r) (pairs[0] is empty).
    Components are numbered by their first voxel in raster order: slabs in z order, a merged component takes the
    number of its member in the lowest slab.  -> (list of uint32 lookup tables [N_r+1], N_global)
    Runs in the library's host code (dlv_resolve_labels): a Python union-find over the ~1.4e5 seam pairs of a whole
    brain on 8 GPUs cost ~0.4 s per step on every rank.
    """
    from ._lib import resolve_labels
    return resolve_labels(counts, pairs)


def merge_tables(tables, luts, z_offsets, n_global, shape_real):
    """Exact merge of per-slab statistics (local z coordinates) into the global table rows 0..N: integer adds /
    min / max per global label, then one fp64 divide per centroid coordinate.  Runs in the library's host code
    (dlv_table_merge) - numpy scatter / sort formulations cost 0.3-0.6 s per step at 6e5 components."""
    from ._lib import table_merge
    return table_merge(tables, luts, z_offsets, n_global, shape_real)


# ------------------------------------------------------------------------------------------- volumes beyond one label space
MAX_CCL_VOXELS = (1 << 32) - 2       # dlv_ccl labels provisional components with 32-bit voxel indices


def ccl_any_size(ctx, mask, shape, labels_out=None, max_voxels=MAX_CCL_VOXELS, labels_sink=None, bytes_free=None):
    """``Context.ccl`` for volumes of any size (count_blobs.py:61,85 on a whole brain: 2.4e10 voxels).

    A volume with more voxels than one 32-bit label space - or, for host-resident masks, than the device can hold
    next to its labels - is cut along z into sub-slabs that are labelled one after the other on the same GPU and
    merged exactly like the slabs of different GPUs are: 26-adjacent label pairs across every cut
    (dlv_ccl_boundary_pairs) -> global numbering by first voxel in raster order (resolve_global_labels) ->
    dlv_relabel per sub-slab, integer tables merged associatively.  ``mask`` / ``labels_out``: numpy arrays (memmaps)
    or device tensors of shape ``shape`` (uint8 / uint32-or-int32).  ``labels_sink(z0, z1, labels_dev)``: called
    once per sub-slab, in z order, with the globally numbered int32 device labels of planes [z0, z1) (count_blobs
    streams them into the ``-cc3d.npy`` file instead of holding 4 B/voxel in host memory); sub-slabs whose labels
    do not fit on the device next to the others are labelled a second time for it (20 ms per 4e9 voxels).
    Same return value as ``Context.ccl``."""
    import torch
    Z, Y, X = (int(v) for v in shape)
    plane = Y * X
    host_mask = isinstance(mask, np.ndarray)
    dev_labels = labels_out is not None and not isinstance(labels_out, np.ndarray)
    if plane > max_voxels:
        raise ValueError(f"one plane ({Y}x{X}) exceeds the {max_voxels}-voxel label space of a sub-slab")
    zs = max(1, max_voxels // plane)
    if not dev_labels and (host_mask or labels_sink is not None):
        # per sub-slab: uploaded mask (1 B) + labels (4 B) + bit mask / scan scratch (< 1 B per voxel)
        free = int(bytes_free) if bytes_free is not None else int(torch.cuda.mem_get_info(ctx.device)[0])
        zs = max(1, min(zs, int(free * 0.7) // (6 * plane)))
    else:
        free = int(bytes_free) if bytes_free is not None else 0
    if zs >= Z and labels_sink is None:
        return ctx.ccl(mask, shape, labels_out=labels_out)
    dev = torch.device("cuda", ctx.device)
    stream = torch.cuda.ExternalStream(ctx._L.dlv_stream(ctx._h), device=dev)
    m3 = mask.reshape(Z, Y, X) if host_mask else mask.view(Z, Y, X)
    l3 = None
    if labels_out is not None:
        l3 = labels_out.reshape(Z, Y, X) if isinstance(labels_out, np.ndarray) else labels_out.view(Z, Y, X)
    host_labels = isinstance(l3, np.ndarray)
    cuts = list(range(0, Z, zs)) + [Z]
    spans = list(zip(cuts[:-1], cuts[1:]))
    tables, counts, pairs, prev_last = [], [], [None], None
    kept, kept_bytes = {}, 0                       # device labels held for the second pass (sink / host copy)
    second = labels_sink is not None or host_labels
    with torch.cuda.stream(stream):
        for i, (z0, z1) in enumerate(spans):
            sub = (z1 - z0, Y, X)
            # labels of this sub-slab: the caller's buffer when it is on the device, else a device scratch (the first
            # and last planes are needed on the device for the seam pairs anyway)
            lab = l3[z0:z1] if dev_labels else torch.empty(sub, dtype=torch.int32, device=dev)
            t = ctx.ccl(m3[z0:z1], sub, labels_out=lab)
            tables.append(t)
            counts.append(t["n"])
            if prev_last is not None:
                pairs.append(ctx.ccl_boundary_pairs(prev_last, lab[0]))
            prev_last = lab[-1].clone()
            if second and not dev_labels and kept_bytes + lab.numel() * 4 <= free * 0.7 - 6 * zs * plane:
                kept[i] = lab
                kept_bytes += lab.numel() * 4
            del lab
        luts, n_global = resolve_global_labels(counts, pairs)
        if l3 is not None or labels_sink is not None:
            for i, ((z0, z1), lut) in enumerate(zip(spans, luts)):
                lut_dev = torch.from_numpy(lut.view(np.int32)).to(dev)
                if dev_labels:
                    lab = l3[z0:z1]
                elif i in kept:
                    lab = kept.pop(i)
                else:                       # did not fit: label the sub-slab again (same local numbering)
                    lab = torch.empty((z1 - z0, Y, X), dtype=torch.int32, device=dev)
                    ctx.ccl(m3[z0:z1], (z1 - z0, Y, X), labels_out=lab)
                ctx.relabel(lab, lut_dev)
                if host_labels:
                    l3[z0:z1] = lab.cpu().numpy().view(l3.dtype)
                if labels_sink is not None:
                    ctx.synchronize()
                    labels_sink(z0, z1, lab)
                del lab
        ctx.synchronize()
    return merge_tables(tables, luts, cuts[:-1], n_global, (Z, Y, X))


# ------------------------------------------------------------------------------------------- communication
class TorchComm:
    """torch.distributed plumbing of the slab drivers (NCCL on GPUs, gloo in the CPU tests).

    Point-to-point transfers go through ``batch_isend_irecv`` (NCCL otherwise warns that un-batched P2P ops are
    serialised with every other op on the communicator); the small host-side arrays (activity flags, seam pairs,
    statistics tables) are exchanged as padded fixed-width integer tensors, not as pickled objects.
    ``device``: where collective buffers live - the rank's GPU under NCCL, None (CPU) under gloo."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        if device is None and dist.get_backend() == "nccl":       # NCCL moves device tensors only
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device

    def _p2p(self, op, tensor, peer):
        for w in self.dist.batch_isend_irecv([self.dist.P2POp(op, tensor, peer)]):
            w.wait()

    def send(self, tensor, src, dst, tag):
        self._p2p(self.dist.isend, tensor.contiguous(), dst)

    def recv(self, like, src, dst, tag):
        self._p2p(self.dist.irecv, like, src)
        return like

    def barrier(self):
        self.dist.barrier()

    def allgather_array(self, arr):
        """Every rank's array (same dtype and trailing dimensions, any length) -> list of numpy arrays, one per rank."""
        torch = self.torch
        a = np.ascontiguousarray(arr)
        kind = {1: np.uint8, 2: np.int16, 4: np.int32, 8: np.int64}[a.dtype.itemsize]
        flat = torch.from_numpy(a.reshape(-1).view(kind).copy())
        n = torch.tensor([flat.numel()], dtype=torch.int64)
        if self.device is not None:
            n = n.to(self.device)
        sizes = [torch.empty_like(n) for _ in range(self.world)]
        self.dist.all_gather(sizes, n)
        sizes = [int(x.item()) for x in sizes]
        cap = max(max(sizes), 1)
        buf = torch.zeros(cap, dtype=flat.dtype, device=self.device if self.device is not None else "cpu")
        buf[:flat.numel()] = flat.to(buf.device)
        out = torch.empty(self.world * cap, dtype=flat.dtype, device=buf.device)
        self.dist.all_gather_into_tensor(out, buf)
        out = out.cpu().numpy().reshape(self.world, cap)
        trail = a.shape[1:]
        return [out[r, :sizes[r]].view(a.dtype).reshape((-1,) + trail) for r in range(self.world)]

    def allgather_int(self, v):
        return [int(x[0]) for x in self.allgather_array(np.array([int(v)], dtype=np.int64))]

    def allgather(self, obj):
        """Pickled objects - kept for callers outside the per-step path."""
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out


def pack_table(table):
    """Statistics table -> ONE int64 vector of 10 (N + 1) entries, column blocks [counts | 3 coordinate sums | 6 box
    bounds] (uint64 bit patterns kept).  Column blocks, not rows: the receiving side then takes its three arrays as
    contiguous views of the gathered buffer (a row-interleaved [N + 1, 10] layout cost 0.2 s of strided copies per step
    for the 2.5 M components of a whole brain on 8 ranks)."""
    if table is None:
        return np.zeros(0, dtype=np.int64)
    n1 = int(table["n"]) + 1
    out = np.empty(n1 * 10, dtype=np.int64)
    out[:n1] = np.asarray(table["voxel_counts"]).reshape(-1).view(np.int64)
    out[n1:4 * n1] = np.asarray(table["sums"]).reshape(-1).view(np.int64)
    out[4 * n1:] = np.asarray(table["bounding_boxes"]).reshape(-1)
    return out


def unpack_table(vec):
    """Inverse of pack_table, without copies: views of ``vec``."""
    if len(vec) == 0:
        return None
    vec = np.ascontiguousarray(vec)
    n1 = len(vec) // 10
    return {"n": n1 - 1, "voxel_counts": vec[:n1].view(np.uint64), "sums": vec[n1:4 * n1].view(np.uint64).reshape(n1, 3),
            "bounding_boxes": vec[4 * n1:].reshape(n1, 6)}


# ------------------------------------------------------------------------------------------- overlapped slab upload
class StreamedSlab:
    """A device slab whose planes are still arriving.  ``feed(copy_chunk)`` runs in a background thread:
    ``copy_chunk(a, b)`` must enqueue the upload of planes [a, b) on ``self.side`` (a second CUDA stream); after each
    chunk an event is recorded.  The consumer iterates ``chunks()`` -> (planes landed so far, is_last) and calls
    ``wait(stream)`` to order its stream after the chunk just yielded."""

    def __init__(self, tensor, bounds, side_stream, torch):
        import queue
        import threading
        self.tensor, self.bounds, self.side, self.torch = tensor, list(bounds), side_stream, torch
        self._q = queue.Queue()
        self._threading = threading
        self._event = None
        self._thread = None
        self._error = None

    def feed(self, copy_chunk, after=None):
        def work():
            try:
                with self.torch.cuda.stream(self.side):
                    if after is not None:
                        self.side.wait_event(after)
                    a = 0
                    for b in self.bounds:
                        copy_chunk(a, b)
                        ev = self.torch.cuda.Event()
                        ev.record(self.side)
                        self._q.put((b, ev))
                        a = b
            except BaseException as e:      # surfaced in the consumer
                self._error = e
                self._q.put((None, None))
        self._thread = self._threading.Thread(target=work, daemon=True)
        self._thread.start()
        return self

    def chunks(self):
        for i in range(len(self.bounds)):
            b, ev = self._q.get()
            if b is None:
                raise self._error
            self._event = ev
            yield b, i == len(self.bounds) - 1
        self._thread.join()

    def wait(self, stream):
        stream.wait_event(self._event)


def layer_bounds(plan, rank):
    """Upload chunk ends (slab-local planes) for rank's slab: where each of its window z-layers ends, then the rest."""
    info = plan.rank(rank)
    z0, z1 = info["slab"]
    rz = plan.roi[0]
    ends = sorted({min(z1, plan.sz[l] + rz) - z0 for l in range(info["layers"][0], info["layers"][1])})
    ends = [e for e in ends if 0 < e < z1 - z0]
    return ends + [z1 - z0]


# ------------------------------------------------------------------------------------------- CUDA workers
class _CudaLabelOps:
    """Labelling stage of a slab of device-resident binaries (``self.binaries``): local labels + table, the seam
    planes and pairs, the final relabelling.  Needs self.ctx / torch / dev / stream / binaries."""

    def ccl(self):
        torch = self.torch
        with torch.cuda.stream(self.stream):
            self.labels = torch.empty(self.binaries.shape, dtype=torch.int32, device=self.dev)
        if self.binaries.shape[0] == 0:
            self.table = None
            return 0
        # a slab thicker than one 32-bit label space (cfg4 on 2 or 4 GPUs) is labelled in sub-slabs
        self.table = ccl_any_size(self.ctx, self.binaries, tuple(self.binaries.shape), labels_out=self.labels)
        return self.table["n"]

    def first_plane(self):
        return self.labels[0]

    def last_plane(self):
        return self.labels[-1]

    def empty_plane(self):
        with self.torch.cuda.stream(self.stream):
            return self.torch.empty(tuple(self.binaries.shape[1:]), dtype=self.torch.int32, device=self.dev)

    def boundary_pairs(self, lo_plane):
        return self.ctx.ccl_boundary_pairs(lo_plane, self.labels[0])

    def relabel(self, lut):
        if self.labels.numel():
            with self.torch.cuda.stream(self.stream):
                lut_dev = self.torch.from_numpy(lut.view(np.int32)).to(self.dev)
            self.ctx.relabel(self.labels, lut_dev)


class CudaLabelSlab(_CudaLabelOps):
    """A slab of binaries on the device, for the labelling stage alone (count_blobs over several GPUs)."""

    def __init__(self, ctx, binaries):
        import torch
        self.torch, self.ctx = torch, ctx
        self.dev = torch.device("cuda", ctx.device)
        self.stream = torch.cuda.ExternalStream(ctx._L.dlv_stream(ctx._h), device=self.dev)
        self.binaries = binaries
        self.labels = self.table = None


class CudaSlabWorker(_CudaLabelOps):
    """One rank's compute, every stage through libdelivr_b200.so."""

    def __init__(self, ctx, plan, rank, planes_fn, window_batch=0, threshold=0.5, tta=False, erosion_block_planes=0,
                 blend_mode=0, want_sigmoid=False, keep_avg=False):
        import torch
        self.torch = torch
        self.ctx, self.plan, self.r = ctx, plan, rank
        self.info = plan.rank(rank)
        self.dev = torch.device("cuda", ctx.device)
        # every torch-side operation of the worker (allocation fills, adds, NCCL sends) is enqueued on the library's
        # own (non-blocking) stream, so it is ordered with the kernels without any cross-stream event
        self.stream = torch.cuda.ExternalStream(ctx._L.dlv_stream(ctx._h), device=self.dev)
        self.window_batch, self.threshold, self.tta, self.ebp = window_batch, threshold, tta, erosion_block_planes
        self.blend_mode, self.want_sigmoid, self.keep_avg = blend_mode, want_sigmoid, keep_avg
        self.sigmoid = self.avg_own = None
        z0, z1 = self.info["slab"]
        self.loading = None
        with torch.cuda.stream(self.stream):
            got = planes_fn(z0, z1) if z1 > z0 else None              # uint16 (z1-z0, PY, PX) on the device
        if isinstance(got, StreamedSlab):                             # planes still arriving: see accumulate()
            self.loading, got = got, got.tensor
        self.slab = got
        self.acc = None

    def _schedule(self, sel):
        # inference.py:265-279: 13 passes = 5 x plain, 4 x flip z (dim 2), 4 x flip y (dim 3); identical passes are
        # evaluated once and blended `repeat` times (flip_dim | (repeat - 1) << 8, see dlv_seg_accumulate)
        flips = [0 | (4 << 8), 2 | (3 << 8), 3 | (3 << 8)] if self.tta else [0]
        if not len(sel):
            return np.zeros((0, 4), dtype=np.int32)
        return np.concatenate([np.concatenate([sel, np.full((len(sel), 1), f, dtype=np.int32)], axis=1) for f in flips])

    def _run(self, sched):
        if len(sched):
            self.ctx.seg_accumulate(self.slab, sched, self.plan.roi, self.acc, window_batch=self.window_batch, blend_mode=self.blend_mode,
                                    shape_pad=self.plan.shape_pad, overlap=self.plan.overlap, gz0=self.info["slab"][0])

    def accumulate(self):
        torch = self.torch
        if self.slab is None:
            return np.zeros(0, dtype=np.int32)
        z0 = self.info["slab"][0]
        wins = self.plan.windows_of(self.r)
        local = wins.copy()
        local[:, 0] -= z0
        with torch.cuda.stream(self.stream):
            self.acc = torch.zeros(self.slab.shape, dtype=torch.int32, device=self.dev)
        if self.loading is None:
            active = self.ctx.windows_active(self.slab, local, self.plan.roi)
            self._run(self._schedule(local[active != 0]))
            return active
        # the slab is still being uploaded (file read -> pinned staging -> device on a second stream): windows run as
        # soon as the chunk that holds their last plane has landed, whole batches at a time, the upload of the next
        # chunk overlapping them.  The blend is an integer sum: any order gives the same accumulator.
        active = np.zeros(len(local), dtype=np.int32)
        done, pending = 0, np.zeros((0, 3), dtype=np.int32)
        batch = max(1, self.window_batch or 128)
        rz = self.plan.roi[0]
        for z_end, last in self.loading.chunks():
            self.loading.wait(self.stream)                            # planes [0, z_end) of the slab are on the device
            n = done
            while n < len(local) and local[n, 0] + rz <= z_end:      # windows are z-major: eligible ones are a prefix
                n += 1
            if n > done:
                active[done:n] = self.ctx.windows_active(self.slab, local[done:n], self.plan.roi)
                pending = np.concatenate([pending, local[done:n][active[done:n] != 0]])
                done = n
            run = len(pending) if last else len(pending) // batch * batch
            if run:
                self._run(self._schedule(pending[:run]))
                pending = pending[run:]
        assert done == len(local) and len(pending) == 0
        self.loading = None
        return active

    def acc_planes(self, g0, g1):
        z0 = self.info["slab"][0]
        return self.acc[g0 - z0: g1 - z0]

    def add_planes(self, g0, g1, t):
        z0 = self.info["slab"][0]
        with self.torch.cuda.stream(self.stream):
            self.acc[g0 - z0: g1 - z0] += t

    def finalise(self, active_global):
        torch = self.torch
        o0, o1 = self.info["own_real"]
        Z, Y, X = self.plan.shape_real
        if self.slab is None or o1 <= o0:
            with torch.cuda.stream(self.stream):
                self.binaries = torch.zeros((0, Y, X), dtype=torch.uint8, device=self.dev)
            return self.binaries
        z0, z1 = self.info["slab"]
        self.ctx.seg_average(self.acc, z1 - z0, z0, self.plan.shape_pad, self.plan.roi, self.plan.overlap, active_global,
                             passes=13 if self.tta else 1, blend_mode=self.blend_mode)
        avg = self.acc.view(torch.float32)
        with torch.cuda.stream(self.stream):
            self.binaries = torch.empty((o1 - o0, Y, X), dtype=torch.uint8, device=self.dev)
            if self.want_sigmoid:        # network_output.npy (inference.py:43,72)
                self.sigmoid = torch.empty((o1 - o0, Y, X), dtype=torch.float32, device=self.dev)
        self.ctx.op_finalise_slab(avg, self.slab, z1 - z0, z0, self.plan.shape_real, o0, o1, self.binaries,
                                  threshold=self.threshold, erosion_iters=self.plan.iters, erosion_block_planes=self.ebp,
                                  sigmoid_out=self.sigmoid)
        if self.keep_avg:                # averaged logits of the planes this rank owns (inference_output.npy, :246)
            a0, a1 = self.info["own"]
            self.avg_own = avg[a0 - z0: a1 - z0]
        self.acc = None
        return self.binaries


# ------------------------------------------------------------------------------------------- drivers
def run_virtual(workers, plan):
    """All ranks in this process (LocalComm).  Returns the merged table; workers keep binaries / labels."""
    world = plan.world
    with _stream_ctx(workers[0]):
        return _run_virtual(workers, plan)


def _run_virtual(workers, plan):
    world = plan.world
    active = [w.accumulate() for w in workers]
    for r in range(world):                                    # exchange 1: logit halo, r -> next non-empty rank
        info = plan.rank(r)
        q = plan._next_nonempty(r)
        if info["send"] is not None and q is not None:
            g0, g1 = info["send"]
            workers[q].add_planes(g0, g1, workers[r].acc_planes(g0, g1))
    active_global = np.concatenate(active) if active else np.zeros(0, np.int32)   # exchange 2
    for w in workers:
        w.finalise(active_global)
    counts = [w.ccl() for w in workers]
    pairs = [None] * world                                    # exchange 3: boundary label planes
    for r in range(world):
        p = _prev_with_planes(workers, r)
        if p is not None and workers[r].labels.shape[0] > 0:
            pairs[r] = (p, workers[r].boundary_pairs(workers[p].last_plane()))
    return _merge(workers, plan, counts, pairs)


def _prev_with_planes(workers, r):
    for q in range(r - 1, -1, -1):
        if workers[q].labels.shape[0] > 0:
            return q
    return None


def _merge(workers, plan, counts, pairs_with_src):
    """Shared tail of both drivers when every rank's objects are at hand (virtual mode)."""
    world = plan.world
    # re-express pairs against the chain of non-empty slabs: resolve_global_labels expects (r-1, r) adjacency
    nonempty = [r for r in range(world) if counts[r] > 0 or workers[r].labels.shape[0] > 0]
    idx = {r: i for i, r in enumerate(nonempty)}
    c2 = [counts[r] for r in nonempty]
    p2 = [None] * len(nonempty)
    for r in nonempty:
        if pairs_with_src[r] is not None:
            src, p = pairs_with_src[r]
            assert idx[src] == idx[r] - 1
            p2[idx[r]] = p
    luts2, n_global = resolve_global_labels(c2, p2)
    luts = [np.zeros(1, np.uint32)] * world
    for r in nonempty:
        luts[r] = luts2[idx[r]]
        workers[r].relabel(luts[r])
    tables = [w.table for w in workers]
    zoff = [plan.rank(r)["own_real"][0] for r in range(world)]
    return merge_tables(tables, luts, zoff, n_global, plan.shape_real)


def _stream_ctx(worker):
    """CUDA workers: make the library stream torch's current stream (NCCL ops and tensor math are then ordered with
    the library's kernels).  CPU workers (tests): nothing to do."""
    import contextlib
    st = getattr(worker, "stream", None)
    if st is None:
        return contextlib.nullcontext()
    return worker.torch.cuda.stream(st)


def run_distributed(worker, plan, comm):
    """One rank per process: segmentation stage, then labelling stage.  Collectives: send/recv of the logit halo and
    of one label plane, padded integer all-gathers of the activity flags, seam pairs and statistics tables."""
    with _stream_ctx(worker):
        tr = _Trace(worker, comm.rank)
        distributed_segment(worker, plan, comm, tr)
        worker.ccl()
        tr.mark("ccl")
        table = distributed_label(worker, comm, plan.shape_real, tr)
        tr.dump()
        return table


def distributed_segment(worker, plan, comm, tr=None):
    """Exchanges 1 and 2 around the window passes: on return ``worker.binaries`` holds the rank's own planes
    (what run_inference writes to binaries.npy).  -> global per-window activity flags."""
    r = comm.rank
    tr = tr or _Trace(worker, r)
    info = plan.rank(r)
    active = worker.accumulate()
    tr.mark("accumulate")
    nxt, prv = plan._next_nonempty(r), plan._prev_nonempty(r)
    have = plan.wrange[r][1] > plan.wrange[r][0]
    # exchange 1: a chain - receive and add first, then send (what is sent may contain what was just received, when
    # this rank owns fewer planes than the previous one touched); stream-ordered on NCCL, blocking pairs on gloo
    if have and info["recv"] is not None and prv is not None:
        g0, g1 = info["recv"]
        buf = worker.acc_planes(g0, g1).clone()
        worker.add_planes(g0, g1, comm.recv(buf, prv, r, "acc"))
    if have and info["send"] is not None and nxt is not None:
        g0, g1 = info["send"]
        comm.send(worker.acc_planes(g0, g1), r, nxt, "acc")
    tr.mark("halo exchange")
    active_global = np.concatenate(comm.allgather_array(np.asarray(active, dtype=np.int32)))     # exchange 2
    tr.mark("flags all-gather")
    worker.finalise(active_global)
    tr.mark("finalise")
    return active_global


def distributed_label(worker, comm, shape_real, tr=None):
    """Exchange 3: ``worker`` has labelled its own planes (``worker.ccl()``: local labels 1..N_r, ``worker.table``);
    on return its labels carry the global numbering and the merged table is returned on every rank.  The slabs are
    the ranks' plane ranges in rank order (z offsets = running sum of the plane counts)."""
    r, world = comm.rank, comm.world
    tr = tr or _Trace(worker, r)
    n_local = 0 if worker.table is None else int(worker.table["n"])
    meta = comm.allgather_array(np.array([int(worker.labels.shape[0]), n_local], dtype=np.int64))
    nplanes = [int(m[0]) for m in meta]
    counts = [int(m[1]) for m in meta]
    src = next((q for q in range(r - 1, -1, -1) if nplanes[q] > 0), None)
    dst = next((q for q in range(r + 1, world) if nplanes[q] > 0), None)
    pairs = np.zeros((0, 2), dtype=np.uint32)

    def _send_l():
        if nplanes[r] > 0 and dst is not None:
            comm.send(worker.last_plane(), r, dst, "lab")

    def _recv_l():
        nonlocal pairs
        if nplanes[r] > 0 and src is not None:
            lo = comm.recv(worker.empty_plane(), src, r, "lab")
            pairs = np.asarray(worker.boundary_pairs(lo), dtype=np.uint32).reshape(-1, 2)

    order = [q for q in range(world) if nplanes[q] > 0]
    pos = order.index(r) if r in order else 0
    if pos % 2 == 0:
        _send_l(); _recv_l()
    else:
        _recv_l(); _send_l()
    tr.mark("label plane + pairs")
    all_pairs = comm.allgather_array(pairs)
    all_rows = comm.allgather_array(pack_table(worker.table))
    tr.mark("pairs / tables all-gather")
    luts2, n_global = resolve_global_labels([counts[q] for q in order], [all_pairs[q] for q in order])
    luts = [np.zeros(1, np.uint32)] * world
    for i, q in enumerate(order):
        luts[q] = luts2[i]
    worker.relabel(luts[r])
    tr.mark("resolve + relabel")
    zoff = np.concatenate([[0], np.cumsum(nplanes)])[:-1]
    table = merge_tables([unpack_table(rows) for rows in all_rows], luts, zoff, n_global, shape_real)
    tr.mark("table merge")
    return table


def run_streamed(make_worker, plan, sink=None, label=True):
    """The virtual slabs of ``plan`` one after the other in ONE process, each released before the next is loaded: the
    out-of-core mode of a single GPU (the reference streams the volume from memmaps, inference/inference.py:234,
    244-247).  ``make_worker(r)`` builds slab r's worker (loading its planes); what a slab touches beyond its own
    planes is kept (a few window depths of int32 sums) and added to the next slab.  ``sink(r, worker)`` is called once
    slab r's binaries are final (own planes only) - e.g. to copy them into binaries.npy.  With ``label`` the slabs
    are labelled as they pass and the exact global table is returned (labels themselves are not kept), else None.
    Bit-identical to the single-slab run (tested)."""
    world = plan.world
    pending = None                      # (g0, g1, tensor) sums handed down the chain
    active = []
    tables, counts, pairs, nplanes, prev_last = [], [], [], [], None
    for r in range(world):
        info = plan.rank(r)
        have = plan.wrange[r][1] > plan.wrange[r][0]
        w = make_worker(r)
        with _stream_ctx(w):
            act = w.accumulate()
            active.append(np.asarray(act, dtype=np.int32))
            if have and info["recv"] is not None and pending is not None:
                g0, g1, t = pending
                assert (g0, g1) == tuple(info["recv"])
                w.add_planes(g0, g1, t)
                pending = None
            if have and info["send"] is not None and plan._next_nonempty(r) is not None:
                g0, g1 = info["send"]
                pending = (g0, g1, w.acc_planes(g0, g1).clone())
            # flags of the windows of later slabs are not known yet; they never cover a plane this slab owns
            nrest = sum(c1 - c0 for c0, c1 in plan.wrange[r + 1:])
            w.finalise(np.concatenate(active + [np.ones(nrest, dtype=np.int32)]))
            if sink is not None:
                sink(r, w)
            if label:
                n = w.ccl()
                nplanes.append(int(w.labels.shape[0]))
                if nplanes[-1] > 0:
                    tables.append(w.table); counts.append(n)
                    pairs.append(w.boundary_pairs(prev_last) if prev_last is not None else None)
                    prev_last = w.last_plane().clone()
                else:
                    tables.append(None)
        del w
    if not label:
        return None
    keep = [r for r in range(world) if nplanes[r] > 0]
    luts2, n_global = resolve_global_labels(counts, pairs)
    luts = [np.zeros(1, np.uint32)] * world
    for i, r in enumerate(keep):
        luts[r] = luts2[i]
    zoff = np.concatenate([[0], np.cumsum(nplanes)])[:-1]
    return merge_tables(tables, luts, zoff, n_global, plan.shape_real)


# ------------------------------------------------------------------------------------------- bench entry (N > 1)
def balanced_plan(ctx, comm, shape, roi, overlap, planes_fn):
    """Load-balanced partition: every rank scans an equal share of the windows with the
    skip rule's max pre-pass (dlv_windows_active), the per-window flags are all-gathered and the window list is
    re-cut so that every rank runs the same number of ACTIVE windows.  -> (SlabPlan, active counts per window layer)"""
    plan0 = SlabPlan(shape, roi, overlap, comm.world)
    info = plan0.rank(comm.rank)
    act = np.zeros(0, dtype=np.int32)
    if plan0.wrange[comm.rank][1] > plan0.wrange[comm.rank][0]:
        z0, z1 = info["win"]
        slab0 = planes_fn(z0, z1)
        local = plan0.windows_of(comm.rank).copy()
        local[:, 0] -= z0
        act = ctx.windows_active(slab0, local, roi)
        del slab0
    all_act = np.concatenate(comm.allgather_array(np.asarray(act, dtype=np.int32)))
    per_layer = all_act.reshape(len(plan0.sz), -1).sum(axis=1)
    return SlabPlan(shape, roi, overlap, comm.world, window_weights=all_act), per_layer


def _roofline(B, workload, windows_active, conv_ms_max, world, roi=None):
    tf_peak, _, peak_kind = B.peaks()
    roi = tuple(roi or B.ROI)
    flop = 2.0 * B.MAC_PER_PATCH_VOXEL * windows_active * roi[0] * roi[1] * roi[2]
    achieved = flop / (conv_ms_max * 1e-3) / 1e12 / world if conv_ms_max > 0 else 0.0
    traffic, src = B.conv_traffic(workload, windows_active) if roi == tuple(B.ROI) else (None, None)
    return {"bound": "tensor", "kernel": "conv_is_kernel / conv_tc_kernel (all conv/deconv launches of one step)",
            "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s per GPU", "frac": achieved / tf_peak,
            "traffic": traffic, "traffic_source": src, "peak_kind": f"bf16_tflops_sustained, {peak_kind}",
            "conv_ms_per_step_max_rank": conv_ms_max}


class _TimedWorker(CudaSlabWorker):
    """Bench only: device-synchronised wall time of the finalise and labelling stages of one (untimed) step."""
    t_finalise = t_ccl = 0.0

    def _timed(self, fn, *a):
        import time
        self.stream.synchronize()
        t0 = time.perf_counter()
        out = fn(*a)
        self.stream.synchronize()
        return out, (time.perf_counter() - t0) * 1e3

    def finalise(self, active_global):
        out, self.t_finalise = self._timed(super().finalise, active_global)
        return out

    def ccl(self):
        out, self.t_ccl = self._timed(super().ccl)
        return out


CFG5_WINDOWS = (64, 96, 128, 160, 192)
CFG5_OVERLAPS = (0.25, 0.5, 0.75)


def bench_cfg5(B, args, ctx, comm, stream, dev, rank, world, wdesc):
    """BASELINE.json configs[4]: sliding-window patch-size / overlap sweep (64^3 - 192^3, overlap 0.25 - 0.75) on one
    volume sharded over the N GPUs; per point the throughput, the tensor roofline fraction of the convolutions and the
    HBM roofline fractions of the blend, finalise and labelling stages (algorithmic bytes of SURVEY.md section 8d)."""
    wl = B.WORKLOADS["cfg5"]
    _, hbm_peak, _ = B.peaks()
    pts = []
    only = getattr(args, "sweep_only", None)
    for win in CFG5_WINDOWS:
        for ov in CFG5_OVERLAPS:
            if only and f"{win}:{ov}" not in only:
                continue
            roi = (win, win, win)
            r = _bench_job(B, args, ctx, comm, stream, dev, rank, world, wl, True, False, args.steps, min(args.warmup, 1), False, True,
                           roi=roi, overlap=ov, want_stages=True)
            if rank != 0:
                continue
            nv, pv = r["nvox"], float(r["windows_active"]) * win ** 3
            sm = r["stage_ms_max_rank"]
            gbs = lambda nbytes, ms: (nbytes / world / (ms * 1e-3) / 1e9) if ms > 0 else None
            blend, fin, ccl = gbs(10.0 * pv, sm["blend"]), gbs(7.0 * nv, sm["finalise"]), gbs(9.0 * nv, sm["ccl"])
            pts.append({"window": win, "overlap": ov, "gvoxels_per_s": r["gvoxels_per_s"], "ms_per_step": r["ms_per_step"],
                        "windows_active": r["windows_active"], "conv_tflops_per_gpu": r["roofline"]["achieved"],
                        "conv_frac_of_bf16_peak": r["roofline"]["frac"], "non_conv_share": r["non_conv_share"],
                        "blend_gbs_per_gpu": blend, "blend_frac_of_hbm": blend / hbm_peak if blend else None,
                        "finalise_gbs_per_gpu": fin, "finalise_frac_of_hbm": fin / hbm_peak if fin else None,
                        "ccl_gbs_per_gpu": ccl, "ccl_frac_of_hbm": ccl / hbm_peak if ccl else None,
                        "stage_ms_max_rank": sm, "clocks": r["clocks"], "launches": r["launches"]})
    if rank != 0:
        return
    head = next((p for p in pts if p["window"] == 96 and p["overlap"] == 0.5), pts[0])
    shape = wl["shape"]
    B.emit({"metric": "Gvoxels/s seg+CC", "value": head["gvoxels_per_s"], "unit": "Gvoxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": min(args.warmup, 1), "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": f"synthetic; {wdesc}",
            "config": {"workload": f"{wl['name']}, z-slab sharded over {world} GPUs; headline = window {head['window']}^3, overlap {head['overlap']}",
                       "sweep": pts, "tta": False, "blend": "constant",
                       "fused_path": "windows whose x extent keeps the fused input-stationary conv's stages in shared memory (x <= ~150) run it; "
                                     "192^3 runs the per-tap tcgen05 kernel + separate norm passes",
                       "algorithmic_bytes": {"blend_per_active_patch_voxel": 10, "finalise_per_voxel": 7, "ccl_per_voxel": 9},
                       "timing": "CUDA events on the library stream between barriers, max over ranks; stage times from one extra, serialised step",
                       "l2": "inputs larger than L2"},
            "gpu_launches": sum(p["launches"] for p in pts), "clocks": head["clocks"],
            "roofline": {"bound": "tensor", "kernel": "conv_is_kernel / conv_tc_kernel", "achieved": head["conv_tflops_per_gpu"],
                         "peak": B.peaks()[0], "unit": "TFLOP/s per GPU", "frac": head["conv_frac_of_bf16_peak"], "traffic": None}})


class _Trace:
    """DLV_TRACE_SLABS=1: per-phase device-synchronised wall times of one step on stderr (rank 0 prints its own and,
    at the end of a step, nothing else - the phases are bracketed by stream synchronisations, so a traced step is not a
    timed one)."""

    def __init__(self, worker, rank):
        import os
        self.on = os.environ.get("DLV_TRACE_SLABS") == "1"
        self.w, self.rank = worker, rank
        self.rows = []
        if self.on:
            import time
            self.time = time
            self.sync()
            self.t = time.perf_counter()

    def sync(self):
        st = getattr(self.w, "stream", None)
        if st is not None:
            st.synchronize()

    def mark(self, what):
        if not self.on:
            return
        self.sync()
        now = self.time.perf_counter()
        self.rows.append((what, (now - self.t) * 1e3))
        self.t = now

    def dump(self):
        if self.on:
            import sys
            print(f"[slabs rank {self.rank}] " + " | ".join(f"{k} {v:.1f}" for k, v in self.rows) +
                  f" | total {sum(v for _, v in self.rows):.1f} ms", file=sys.stderr, flush=True)


def _bench_job(B, args, ctx, comm, stream, dev, rank, world, wl, whole, tta, steps, warmup, want_e2e, want_roofline,
               roi=None, overlap=None, want_stages=False):
    """One measured configuration of the N-GPU bench: the volume (N stacked copies of the workload, or ONE whole
    volume) is sharded by balanced_plan and every step is the product path - CudaSlabWorker + run_distributed, what
    run_inference / count_blobs drive under torchrun.  -> dict of results on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    from .synth import synth_volume_cuda
    from .inference.inference import erosion_block_planes
    z1, Y, X = wl["shape"]
    roi = tuple(roi or B.ROI)
    overlap = B.OVERLAP if overlap is None else float(overlap)
    shape = (z1, Y, X) if whole else (z1 * world, Y, X)
    evaluated = 3 if tta else 1                # 13 reference passes = 3 distinct ones blended 5 / 4 / 4 times
    PZ, PY, PX = SlabPlan(shape, roi, overlap, world).shape_pad

    def planes(z0, z1_):
        full = torch.zeros((z1_ - z0, PY, PX), dtype=torch.uint16, device=dev)
        r0, r1 = min(z0, shape[0]), min(z1_, shape[0])
        # plane z of the job = plane z mod z1 of the workload's own volume: every GPU added brings one more copy of
        # the single-GPU workload (same active-window fraction), plus the window layer that straddles the seam
        for k in range(r0 // z1, (r1 + z1 - 1) // z1 if r1 > r0 else 0):
            a, b = max(r0, k * z1), min(r1, (k + 1) * z1)
            full[a - z0: b - z0, :Y, :X] = synth_volume_cuda((z1, Y, X), wl["seed"], device=dev, z_range=(a - k * z1, b - k * z1))
        return full

    ebp = erosion_block_planes(shape)
    with torch.cuda.stream(stream):
        plan, per_layer = balanced_plan(ctx, comm, shape, roi, overlap, planes)
        info = plan.rank(rank)
        slab = planes(*info["slab"])
        o0, o1 = info["own_real"]
        hslab = hbin = None
        if want_e2e:
            # end-to-end leg: the rank's slab starts in pinned host memory, binaries end in pinned host memory
            hslab = torch.empty(slab.shape, dtype=torch.uint16).pin_memory()
            hslab.copy_(slab)
            hbin = torch.empty((max(o1 - o0, 0), Y, X), dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()

    def step(host):
        def load(a, b):
            if not host:
                return slab
            d = torch.empty(slab.shape, dtype=torch.uint16, device=dev)
            d.copy_(hslab, non_blocking=True)
            return d
        w = (_TimedWorker if timing["on"] else CudaSlabWorker)(ctx, plan, rank, load, erosion_block_planes=ebp, tta=tta)
        table = run_distributed(w, plan, comm)
        if timing["on"]:
            timing.update(finalise=w.t_finalise, ccl=w.t_ccl)
        if host:
            with torch.cuda.stream(stream):
                hbin.copy_(w.binaries, non_blocking=True)
            stream.synchronize()
        return table, w

    timing = {"on": False}

    def timed(host, n):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            table, w = step(host)
            del w
        e1.record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        dt = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)      # device time, max over ranks
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt.item()) * 1e3 / n, table

    for _ in range(warmup):
        step(False)
    sampler = B.ClockSampler(dev.index) if rank == 0 else None
    l0 = ctx.launches
    ms, table = timed(False, steps)
    launches = torch.tensor([ctx.launches - l0], device=dev, dtype=torch.int64)
    dist.all_reduce(launches)
    clocks = sampler.stop() if sampler else None
    ms_e2e, io = None, None
    if want_e2e:
        step(True)
        ms_e2e, _ = timed(True, steps)
        io = torch.tensor([hslab.numel() * 2, hbin.numel()], device=dev, dtype=torch.int64)
        dist.all_reduce(io)
    conv_ms = None
    if want_roofline:
        # roofline of the tcgen05 convolutions: one extra step with per-launch event timing on every rank; the job's
        # algorithmic FLOP / the slowest rank's conv time, per GPU
        ctx.set_conv_timing(True)
        timing["on"] = bool(want_stages)
        step(False)
        timing["on"] = False
        st = ctx.stage_times_ms()
        conv = torch.tensor([st["conv"], st["blend"], st["norm"], st["gather"], timing.get("finalise", 0.0), timing.get("ccl", 0.0)],
                            device=dev, dtype=torch.float64)
        ctx.set_conv_timing(False)
        dist.all_reduce(conv, op=dist.ReduceOp.MAX)
        conv_ms = float(conv[0].item())
        stage_ms = {k: float(conv[i].item()) for i, k in enumerate(("conv", "blend", "norm", "gather", "finalise", "ccl"))}
    del slab, hslab, hbin
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    nvox = int(np.prod(shape))
    nrow = table["n"] + 1
    res = {"shape": list(shape), "nvox": nvox, "ms_per_step": ms, "gvoxels_per_s": nvox / (ms * 1e-3) / 1e9, "steps": steps, "warmup": warmup,
           "tta": tta, "passes_evaluated": evaluated, "components": int(table["n"]), "windows_active": int(per_layer.sum()),
           "active_windows_per_layer": [int(c) for c in per_layer], "layers_per_rank": plan.layers,
           "windows_per_rank": [c1 - c0 for c0, c1 in plan.wrange], "launches": int(launches.item()), "clocks": clocks}
    if want_e2e:
        res["e2e"] = {"value": nvox / (ms_e2e * 1e-3) / 1e9, "unit": "Gvoxels/s", "h2d_bytes_per_step": int(io[0].item()),
                      "d2h_bytes_per_step": int(io[1].item()) + nrow * (8 + 24 + 48), "ms_per_step": ms_e2e,
                      "note": "per-rank pinned host slab in, pinned host binaries + merged table out"}
    if want_roofline:
        res["roofline"] = _roofline(B, args.workload, int(per_layer.sum()) * evaluated, conv_ms, world, roi)
        res["stage_ms_max_rank"] = stage_ms
        # strong-scaling view of the same step: what the job would take if only the convolutions ran, perfectly split
        res["non_conv_share"] = 1.0 - conv_ms / ms if ms > 0 else None
    return res


def bench_main(args, rank, local_rank, world):
    """bench.py --gpus N.  Headline: weak scaling - N copies of the workload's volume stacked along z (the job at N = 1
    is the single-GPU workload itself), window z-layers partitioned over the ranks by active-window count.  With the
    default workload the same line also carries BASELINE.json configs[3] in ``config.cfg4``: ONE whole-brain volume
    (1500 x 4000 x 4000) sharded over the N GPUs (strong scaling), without and with the reference's default test-time
    augmentation.  --workload cfg4 makes that volume the headline instead."""
    import torch
    import torch.distributed as dist
    from . import Context
    import bench as B

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if not dist.is_initialized():
        import datetime
        # a rank that fails must surface as an abort within minutes, not after NCCL's 10-minute default
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    wl = B.WORKLOADS[args.workload]
    tta = bool(getattr(args, "tta", False))
    whole = bool(wl.get("whole"))              # one fixed volume sharded over the ranks (strong scaling) instead of N copies
    sd, wdesc = B.state_dict()
    ctx = Context(local_rank)
    ctx.load_weights(sd)
    stream = torch.cuda.ExternalStream(ctx._L.dlv_stream(ctx._h), device=dev)
    comm = TorchComm(dev)
    if args.workload == "cfg5":
        bench_cfg5(B, args, ctx, comm, stream, dev, rank, world, wdesc)
        dist.destroy_process_group()
        return
    # the end-to-end leg keeps a second device copy of the slab while it uploads: a whole brain on fewer than 8 GPUs has no room for it
    head = _bench_job(B, args, ctx, comm, stream, dev, rank, world, wl, whole, tta, args.steps, args.warmup, (not whole) or world >= 8, True)
    cfg4 = None
    if not whole and args.workload == "cfg2" and not getattr(args, "no_cfg4", False):
        w4 = B.WORKLOADS["cfg4"]
        cfg4 = {}
        for name, t in (("tta_off", False), ("tta_on", True)):
            # bounded cost (the line must finish within minutes at every N): TTA off = 1 warm-up + 2 timed steps + the
            # roofline step (+ the end-to-end leg on 8 GPUs); TTA on (3 evaluated passes, ~3x longer) = ONE timed step,
            # warm from the TTA-off steps, no separate roofline step (same kernels, same mix)
            r = _bench_job(B, args, ctx, comm, stream, dev, rank, world, w4, True, t, 1 if t else 2, 0 if t else 1, world >= 8 and not t, not t)
            if rank == 0:
                cfg4[name] = {"seconds_per_volume": r["ms_per_step"] * 1e-3, "gvoxels_per_s": r["gvoxels_per_s"], "steps": r["steps"],
                              "warmup": r["warmup"] if not t else "warm from the tta_off steps", "passes_evaluated": r["passes_evaluated"],
                              "windows_active": r["windows_active"], "components": r["components"],
                              "windows_per_rank": r["windows_per_rank"], "ms_per_step": r["ms_per_step"]}
                if "roofline" in r:
                    cfg4[name].update(conv_tflops_per_gpu=r["roofline"]["achieved"], conv_frac=r["roofline"]["frac"], non_conv_share=r["non_conv_share"])
                if "e2e" in r:
                    cfg4[name]["seconds_per_volume_e2e"] = r["e2e"]["ms_per_step"] * 1e-3
    if rank == 0:
        shape = head["shape"]
        out = {
            "metric": "Gvoxels/s seg+CC", "value": head["gvoxels_per_s"], "unit": "Gvoxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong" if whole else "weak", "vs_baseline": None,
            "dtype": "bf16", "data": f"synthetic; {wdesc}",
            "config": {"workload": (f"{wl['name']}, z-slab sharded over {world} GPUs" if whole else
                                    f"{world} copies of {wl['name']} stacked along z ({shape[0]}x{shape[1]}x{shape[2]}), z-slab sharded"),
                       "seconds_per_volume": head["ms_per_step"] * 1e-3,
                       "seconds_per_volume_e2e": head["e2e"]["ms_per_step"] * 1e-3 if "e2e" in head else None,
                       "window": list(B.ROI), "overlap": B.OVERLAP, "tta": tta, "passes_evaluated": head["passes_evaluated"], "blend": "constant",
                       "components": head["components"],
                       "layers_per_rank": head["layers_per_rank"], "windows_per_rank": head["windows_per_rank"],
                       "active_windows_per_layer": head["active_windows_per_layer"],
                       "windows_active": head["windows_active"],
                       "timing": "CUDA events on the library stream between barriers, max over ranks",
                       "l2": "inputs larger than L2 (slab + accumulator >> 126 MB)"},
            "gpu_launches": head["launches"], "clocks": head["clocks"],
            "e2e": {k: v for k, v in head["e2e"].items() if k != "ms_per_step"} if "e2e" in head else None,
            "roofline": head["roofline"],
        }
        out["roofline"]["non_conv_share_of_step"] = head["non_conv_share"]
        if cfg4 is not None:
            out["config"]["cfg4"] = dict(cfg4, workload=f"{B.WORKLOADS['cfg4']['name']}, ONE volume z-slab sharded over {world} GPUs (strong scaling; "
                                                         "BASELINE.json configs[3], target <= 60 s per volume on 8 GPUs)")
        B.emit(out)
    dist.destroy_process_group()
