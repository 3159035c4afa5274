"""ctypes binding of libdelivr_b200.so (include/delivr_b200.h).

There is no CPU fallback: if the shared library is missing, or the device is
not sm_100, every entry point raises.  PyTorch is used by callers only for
device memory / streams / torch.distributed plumbing; the library itself has
no torch dependency.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DLV_LIB selects an experiment build of the same sources (csrc/Makefile VARIANT=...); default: the product library
LIB_PATH = os.environ.get("DLV_LIB") or os.path.join(_HERE, "libdelivr_b200.so")

c_i32, c_i64, c_f32, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p


class DlvError(RuntimeError):
    pass


class SegParams(ctypes.Structure):
    _fields_ = [
        ("shape_pad", c_i64 * 3), ("shape_real", c_i64 * 3), ("roi", c_i32 * 3), ("overlap", c_f32),
        ("tta", c_i32), ("threshold", c_f32), ("erosion_iters", c_i32), ("erosion_block_planes", c_i64),
        ("blend_mode", c_i32), ("window_batch", c_i32), ("skip_empty", c_i32), ("flip_dim", c_i32),
    ]


class BlendGeom(ctypes.Structure):          # include/delivr_b200.h: dlv_blend_geom
    _fields_ = [("shape_pad", ctypes.c_int64 * 3), ("overlap", ctypes.c_float), ("gz0", ctypes.c_int64)]


class SegStats(ctypes.Structure):
    _fields_ = [
        ("windows_total", c_i64), ("windows_active", c_i64), ("passes", c_i64), ("kernel_launches", c_i64),
        ("ms_unet", ctypes.c_double), ("ms_finalise", ctypes.c_double), ("ms_conv", ctypes.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Table(ctypes.Structure):
    _fields_ = [("n", c_i64), ("voxel_counts", ctypes.POINTER(ctypes.c_uint64)),
                ("sums", ctypes.POINTER(ctypes.c_uint64)), ("bbox", ctypes.POINTER(c_i64)),
                ("centroids", ctypes.POINTER(ctypes.c_double))]


# every symbol include/delivr_b200.h declares (tests check the .so exports exactly these)
EXPORTS = [
    "dlv_abi_version", "dlv_init", "dlv_destroy", "dlv_last_error", "dlv_launch_count", "dlv_stream",
    "dlv_synchronize", "dlv_set_conv_timing", "dlv_conv_time_ms", "dlv_stage_time_ms", "dlv_load_weights", "dlv_segment", "dlv_ccl", "dlv_table_free",
    "dlv_ccl_last_timing", "dlv_unet_forward", "dlv_op_conv3d", "dlv_op_deconv", "dlv_op_finalise",
    "dlv_window_grid", "dlv_windows_active", "dlv_seg_accumulate", "dlv_seg_average", "dlv_op_finalise_slab",
    "dlv_ccl_boundary_pairs", "dlv_relabel", "dlv_table_merge", "dlv_table_csv", "dlv_resolve_labels", "dlv_tiff_info", "dlv_tiff_read_u16", "dlv_tiff_last_error",
    "dlv_load_tiff_planes", "dlv_tiff_write_planes", "dlv_tiff_write_last_error", "dlv_paint_boxes", "dlv_edt",
]

_lib = None


def load_library():
    """dlopen the extension; raises DlvError (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DlvError(f"{LIB_PATH} is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       f"or `make -C delivr_cfos_b200/csrc`; there is no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    P = ctypes.POINTER
    L.dlv_abi_version.restype = ctypes.c_int
    L.dlv_init.restype = ctypes.c_int
    L.dlv_init.argtypes = [ctypes.c_int, P(c_vp)]
    L.dlv_destroy.restype = None
    L.dlv_destroy.argtypes = [c_vp]
    L.dlv_last_error.restype = ctypes.c_char_p
    L.dlv_last_error.argtypes = [c_vp]
    L.dlv_launch_count.restype = c_i64
    L.dlv_launch_count.argtypes = [c_vp]
    L.dlv_stream.restype = c_vp
    L.dlv_stream.argtypes = [c_vp]
    L.dlv_synchronize.restype = ctypes.c_int
    L.dlv_synchronize.argtypes = [c_vp]
    L.dlv_set_conv_timing.restype = ctypes.c_int
    L.dlv_set_conv_timing.argtypes = [c_vp, ctypes.c_int]
    L.dlv_conv_time_ms.restype = ctypes.c_int
    L.dlv_conv_time_ms.argtypes = [c_vp, P(ctypes.c_double)]
    L.dlv_stage_time_ms.restype = ctypes.c_int
    L.dlv_stage_time_ms.argtypes = [c_vp, ctypes.c_int, P(ctypes.c_double)]
    L.dlv_load_weights.restype = ctypes.c_int
    L.dlv_load_weights.argtypes = [c_vp, ctypes.c_int, P(ctypes.c_char_p), P(c_vp), P(c_i64)]
    L.dlv_segment.restype = ctypes.c_int
    L.dlv_segment.argtypes = [c_vp, c_vp, P(SegParams), c_vp, c_vp, c_vp, P(SegStats)]
    L.dlv_ccl.restype = ctypes.c_int
    L.dlv_ccl.argtypes = [c_vp, c_vp, P(c_i64), ctypes.c_int, c_vp, P(P(Table))]
    L.dlv_table_free.restype = None
    L.dlv_table_free.argtypes = [P(Table)]
    L.dlv_ccl_last_timing.restype = ctypes.c_int
    L.dlv_ccl_last_timing.argtypes = [c_vp, P(ctypes.c_double), P(c_i64)]
    L.dlv_unet_forward.restype = ctypes.c_int
    L.dlv_unet_forward.argtypes = [c_vp, c_vp, ctypes.c_int, P(c_i32), c_vp]
    L.dlv_op_conv3d.restype = ctypes.c_int
    L.dlv_op_conv3d.argtypes = [c_vp, ctypes.c_char_p, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp]
    L.dlv_op_deconv.restype = ctypes.c_int
    L.dlv_op_deconv.argtypes = [c_vp, ctypes.c_char_p, c_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp]
    L.dlv_op_finalise.restype = ctypes.c_int
    L.dlv_op_finalise.argtypes = [c_vp, c_vp, c_vp, P(c_i64), P(c_i64), c_f32, ctypes.c_int, c_i64, c_vp, c_vp]
    L.dlv_window_grid.restype = ctypes.c_int
    L.dlv_window_grid.argtypes = [P(c_i64), P(c_i32), c_f32, P(c_i32), c_vp]
    L.dlv_windows_active.restype = ctypes.c_int
    L.dlv_windows_active.argtypes = [c_vp, c_vp, c_i64, c_i64, c_vp, ctypes.c_int, P(c_i32), c_vp]
    L.dlv_seg_accumulate.restype = ctypes.c_int
    L.dlv_seg_accumulate.argtypes = [c_vp, c_vp, c_i64, c_i64, c_vp, ctypes.c_int, P(c_i32), ctypes.c_int, ctypes.c_int, c_vp, c_vp]
    L.dlv_seg_average.restype = ctypes.c_int
    L.dlv_seg_average.argtypes = [c_vp, c_vp, c_i64, c_i64, P(c_i64), P(c_i32), c_f32, c_vp, ctypes.c_int, ctypes.c_int]
    L.dlv_op_finalise_slab.restype = ctypes.c_int
    L.dlv_op_finalise_slab.argtypes = [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, P(c_i64), c_f32, ctypes.c_int, c_i64,
                                       c_i64, c_i64, c_vp, c_vp]
    L.dlv_ccl_boundary_pairs.restype = ctypes.c_int
    L.dlv_ccl_boundary_pairs.argtypes = [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, P(c_i64)]
    L.dlv_relabel.restype = ctypes.c_int
    L.dlv_relabel.argtypes = [c_vp, c_vp, c_i64, c_vp, c_i64]
    L.dlv_tiff_info.restype = ctypes.c_int
    L.dlv_tiff_info.argtypes = [ctypes.c_char_p, P(c_i64), P(c_i64), P(c_i32), P(c_i32)]
    L.dlv_tiff_read_u16.restype = ctypes.c_int
    L.dlv_tiff_read_u16.argtypes = [ctypes.c_char_p, c_vp, c_i64, c_i64]
    L.dlv_tiff_last_error.restype = ctypes.c_char_p
    L.dlv_tiff_last_error.argtypes = []
    L.dlv_load_tiff_planes.restype = ctypes.c_int
    L.dlv_load_tiff_planes.argtypes = [c_vp, P(ctypes.c_char_p), ctypes.c_int, c_i64, c_i64, c_i32, c_vp, c_vp, c_i64, c_i64,
                                       ctypes.c_int]
    L.dlv_paint_boxes.restype = ctypes.c_int
    L.dlv_paint_boxes.argtypes = [c_vp, c_vp, P(c_i64), c_vp, c_vp, c_i64, ctypes.c_int, ctypes.c_int, P(c_vp), c_i64]
    L.dlv_tiff_write_planes.restype = ctypes.c_int
    L.dlv_tiff_write_planes.argtypes = [P(ctypes.c_char_p), ctypes.c_int, c_vp, c_i64, c_i64, c_i32, c_i32, ctypes.c_int]
    L.dlv_tiff_write_last_error.restype = ctypes.c_char_p
    L.dlv_tiff_write_last_error.argtypes = []
    L.dlv_resolve_labels.restype = ctypes.c_int
    L.dlv_resolve_labels.argtypes = [ctypes.c_int, P(c_i64), P(c_vp), P(c_i64), P(c_vp), P(c_i64)]
    L.dlv_table_csv.restype = c_i64
    L.dlv_table_csv.argtypes = [c_vp, c_vp, c_i64, c_vp, c_i64]
    L.dlv_table_merge.restype = ctypes.c_int
    L.dlv_table_merge.argtypes = [c_i64, ctypes.c_int, P(c_i64), P(c_vp), P(c_vp), P(c_vp), P(c_vp), P(c_i64), P(c_i64),
                                  c_vp, c_vp, c_vp, c_vp]
    L.dlv_edt.restype = ctypes.c_int
    L.dlv_edt.argtypes = [c_vp, c_vp, P(c_i64), P(ctypes.c_double), c_vp]
    if L.dlv_abi_version() != 4:
        raise DlvError("libdelivr_b200.so ABI version mismatch")
    _lib = L
    return L


def _ptr(x):
    """Device/host address of a torch tensor, numpy array or None."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return x.data_ptr()
    return int(x)


class _TableOwner:
    """Keeps a dlv_table alive while numpy views of its arrays exist."""

    def __init__(self, lib, tp):
        self._lib, self._tp = lib, tp

    def __del__(self):
        if self._tp is not None:
            self._lib.dlv_table_free(self._tp)
            self._tp = None


class Context:
    """One GPU context (``dlv_ctx``).  Raises DlvError on any failure."""

    def __init__(self, device=0):
        self._L = load_library()
        h = c_vp()
        rc = self._L.dlv_init(int(device), ctypes.byref(h))
        self._h = h
        if rc != 0:
            msg = self._L.dlv_last_error(h).decode() if h else "no CUDA device / dlv_init failed"
            if h:
                self._L.dlv_destroy(h)
            self._h = None
            raise DlvError(f"dlv_init({device}) failed ({rc}): {msg}")
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._L.dlv_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc, what):
        if rc != 0:
            raise DlvError(f"{what} failed ({rc}): {self._L.dlv_last_error(self._h).decode()}")

    def _after_torch(self, *xs):
        """Order the library's stream after torch's current stream when device tensors are passed in: a tensor that
        torch is still writing (e.g. ``(labels > 0).to(uint8)`` just enqueued) must be complete before the library
        reads it.  Results flow the other way through the entry points' own end-of-call synchronisation."""
        if not any(getattr(x, "is_cuda", False) for x in xs):
            return
        import torch
        cur = torch.cuda.current_stream(self.device)
        lib = self._L.dlv_stream(self._h)
        if cur.cuda_stream != lib:
            torch.cuda.ExternalStream(lib, device=self.device).wait_stream(cur)

    @property
    def launches(self):
        return int(self._L.dlv_launch_count(self._h))

    def synchronize(self):
        self._check(self._L.dlv_synchronize(self._h), "dlv_synchronize")

    def set_conv_timing(self, enable):
        self._check(self._L.dlv_set_conv_timing(self._h, int(bool(enable))), "dlv_set_conv_timing")

    def conv_time_ms(self):
        ms = ctypes.c_double()
        self._check(self._L.dlv_conv_time_ms(self._h, ctypes.byref(ms)), "dlv_conv_time_ms")
        return ms.value

    def stage_times_ms(self):
        """Device time per stage of the window loop since timing was enabled (set_conv_timing): conv, blend, norm, gather."""
        out = {}
        for i, k in enumerate(("conv", "blend", "norm", "gather")):
            ms = ctypes.c_double()
            self._check(self._L.dlv_stage_time_ms(self._h, i, ctypes.byref(ms)), "dlv_stage_time_ms")
            out[k] = ms.value
        return out

    # ---- weights (inference.py:190-200,217-222)
    def load_weights(self, state_dict):
        """state_dict: mapping name -> array-like fp32 (torch tensors or numpy), checkpoint["state_dict"]."""
        names, arrs = [], []
        for k, v in state_dict.items():
            a = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
            names.append(k.encode())
            arrs.append(np.ascontiguousarray(a, dtype=np.float32))
        n = len(names)
        c_names = (ctypes.c_char_p * n)(*names)
        c_ptrs = (c_vp * n)(*[a.ctypes.data for a in arrs])
        c_numel = (c_i64 * n)(*[a.size for a in arrs])
        self._check(self._L.dlv_load_weights(self._h, n, c_names, c_ptrs, c_numel), "dlv_load_weights")

    def load_checkpoint(self, path):
        import torch
        ck = torch.load(path, map_location="cpu", weights_only=True)
        self.load_weights(ck["state_dict"])

    # ---- segmentation
    def segment(self, volume, shape_pad, shape_real, roi, binaries_out, overlap=0.5, tta=False, threshold=0.5,
                erosion_iters=30, erosion_block_planes=0, blend_mode=0, window_batch=0, skip_empty=True,
                flip_dim=None, avg_logits_out=None, sigmoid_out=None):
        p = SegParams()
        self._after_torch(volume, binaries_out, avg_logits_out, sigmoid_out)
        p.shape_pad[:] = [int(s) for s in shape_pad]
        p.shape_real[:] = [int(s) for s in shape_real]
        p.roi[:] = [int(s) for s in roi]
        p.overlap, p.tta, p.threshold = float(overlap), int(bool(tta)), float(threshold)
        p.erosion_iters, p.erosion_block_planes = int(erosion_iters), int(erosion_block_planes)
        p.blend_mode, p.window_batch, p.skip_empty = int(blend_mode), int(window_batch), int(bool(skip_empty))
        p.flip_dim = int(flip_dim or 0)
        st = SegStats()
        rc = self._L.dlv_segment(self._h, _ptr(volume), ctypes.byref(p), _ptr(binaries_out), _ptr(avg_logits_out),
                                 _ptr(sigmoid_out), ctypes.byref(st))
        self._check(rc, "dlv_segment")
        return st.as_dict()

    # ---- connected components (count_blobs.py:61,85)
    def ccl(self, mask, shape, labels_out=None):
        """-> dict(n, voxel_counts uint64[N+1], sums uint64[N+1,3], bounding_boxes int64[N+1,6], centroids f64[N+1,3])."""
        shp = (c_i64 * 3)(*[int(s) for s in shape])
        tp = ctypes.POINTER(Table)()
        self._after_torch(mask, labels_out)
        self._check(self._L.dlv_ccl(self._h, _ptr(mask), shp, 26, _ptr(labels_out), ctypes.byref(tp)), "dlv_ccl")
        # zero-copy views of the library's pinned table block; it goes back to the library's pool when the last
        # array that refers to it is garbage-collected (dlv_table_free)
        owner = _TableOwner(self._L, tp)
        t = tp.contents
        n = int(t.n)

        def view(ptr, ctype, dtype, shape):
            buf = (ctype * int(np.prod(shape))).from_address(ctypes.addressof(ptr.contents))
            buf._owner = owner
            return np.frombuffer(buf, dtype=dtype).reshape(shape)

        return {"n": n,
                "voxel_counts": view(t.voxel_counts, ctypes.c_uint64, np.uint64, (n + 1,)),
                "sums": view(t.sums, ctypes.c_uint64, np.uint64, (n + 1, 3)),
                "bounding_boxes": view(t.bbox, c_i64, np.int64, (n + 1, 6)),
                # cc3d's centroid definition: sum(coord) / count, one fp64 divide (count_blobs.py:85)
                "centroids": view(t.centroids, ctypes.c_double, np.float64, (n + 1, 3))}

    def ccl_last_timing(self):
        ms, ln = ctypes.c_double(), c_i64()
        self._check(self._L.dlv_ccl_last_timing(self._h, ctypes.byref(ms), ctypes.byref(ln)), "dlv_ccl_last_timing")
        return ms.value, ln.value

    # ---- operator-level entry points (device pointers)
    def unet_forward(self, windows_u16, roi, logits_out):
        nwin = int(windows_u16.shape[0])
        self._after_torch(windows_u16, logits_out)
        r = (c_i32 * 3)(*[int(s) for s in roi])
        self._check(self._L.dlv_unet_forward(self._h, _ptr(windows_u16), nwin, r, _ptr(logits_out)), "dlv_unet_forward")

    def op_conv3d(self, layer, x, y_out, stats_out):
        n, _, D, H, W = [int(s) for s in x.shape]
        self._after_torch(x, y_out, stats_out)
        self._check(self._L.dlv_op_conv3d(self._h, layer.encode(), _ptr(x), n, D, H, W, _ptr(y_out), _ptr(stats_out)),
                    "dlv_op_conv3d")

    def op_deconv(self, upcat, x, y_out):
        n, _, D, H, W = [int(s) for s in x.shape]
        self._after_torch(x, y_out)
        self._check(self._L.dlv_op_deconv(self._h, upcat.encode(), _ptr(x), n, D, H, W, _ptr(y_out)), "dlv_op_deconv")

    def op_finalise(self, avg_logits, volume, shape_pad, shape_real, binaries_out, threshold=0.5, erosion_iters=30,
                    erosion_block_planes=0, sigmoid_out=None):
        sp = (c_i64 * 3)(*[int(s) for s in shape_pad])
        sr = (c_i64 * 3)(*[int(s) for s in shape_real])
        self._after_torch(avg_logits, volume, binaries_out, sigmoid_out)
        self._check(self._L.dlv_op_finalise(self._h, _ptr(avg_logits), _ptr(volume), sp, sr, float(threshold),
                                            int(erosion_iters), int(erosion_block_planes), _ptr(binaries_out),
                                            _ptr(sigmoid_out)), "dlv_op_finalise")

    # ---- slab-level entry points (z-sharded runs, delivr_cfos_b200/slabs.py)
    def windows_active(self, slab, origins, roi):
        """origins int32 [n,3] local to the slab (device uint16 (SZ,SY,SX)) -> int32 [n] activity flags."""
        o = np.ascontiguousarray(origins, dtype=np.int32)
        out = np.zeros(len(o), dtype=np.int32)
        self._after_torch(slab)
        r = (c_i32 * 3)(*[int(v) for v in roi])
        self._check(self._L.dlv_windows_active(self._h, _ptr(slab), int(slab.shape[1]), int(slab.shape[2]), o.ctypes.data,
                                               len(o), r, out.ctypes.data), "dlv_windows_active")
        return out

    def seg_accumulate(self, slab, windows, roi, acc, window_batch=0, blend_mode=0, shape_pad=None, overlap=0.5, gz0=0):
        """windows int32 [n,4] = (oz,oy,ox,flip_dim) local to the slab; acc int32 device tensor shaped like slab.
        The gaussian blend (blend_mode 1) also needs the window grid: padded shape, overlap, global plane of slab[0]."""
        w = np.ascontiguousarray(windows, dtype=np.int32).reshape(-1, 4)
        self._after_torch(slab, acc)
        r = (c_i32 * 3)(*[int(v) for v in roi])
        geom = None
        if blend_mode:
            geom = BlendGeom()
            geom.shape_pad[:] = [int(v) for v in shape_pad]
            geom.overlap, geom.gz0 = float(overlap), int(gz0)
        self._check(self._L.dlv_seg_accumulate(self._h, _ptr(slab), int(slab.shape[1]), int(slab.shape[2]), w.ctypes.data,
                                               len(w), r, int(window_batch), int(blend_mode),
                                               ctypes.byref(geom) if geom is not None else None, _ptr(acc)), "dlv_seg_accumulate")

    def seg_average(self, acc, nplanes, gz0, shape_pad, roi, overlap, active, passes=1, blend_mode=0):
        sp = (c_i64 * 3)(*[int(v) for v in shape_pad])
        r = (c_i32 * 3)(*[int(v) for v in roi])
        a = np.ascontiguousarray(active, dtype=np.int32)
        self._after_torch(acc)
        self._check(self._L.dlv_seg_average(self._h, _ptr(acc), int(nplanes), int(gz0), sp, r, float(overlap), a.ctypes.data,
                                            int(passes), int(blend_mode)), "dlv_seg_average")

    def op_finalise_slab(self, avg, volume, nplanes, gz0, shape_real, oz0, oz1, binaries_out, threshold=0.5,
                         erosion_iters=30, erosion_block_planes=0, sigmoid_out=None):
        sr = (c_i64 * 3)(*[int(v) for v in shape_real])
        self._after_torch(avg, volume, binaries_out, sigmoid_out)
        self._check(self._L.dlv_op_finalise_slab(self._h, _ptr(avg), _ptr(volume), int(volume.shape[1]), int(volume.shape[2]),
                                                 int(nplanes), int(gz0), sr, float(threshold), int(erosion_iters),
                                                 int(erosion_block_planes), int(oz0), int(oz1), _ptr(binaries_out),
                                                 _ptr(sigmoid_out)), "dlv_op_finalise_slab")

    def ccl_boundary_pairs(self, labels_lo_plane, labels_hi_plane):
        """-> uint32 [k,2] unique (lo label, hi label) pairs of 26-adjacent voxels across the boundary."""
        import torch
        Y, X = int(labels_hi_plane.shape[-2]), int(labels_hi_plane.shape[-1])
        cap = max(1024, Y * X // 8)
        self._after_torch(labels_lo_plane, labels_hi_plane)
        while True:
            buf = torch.empty((cap, 2), dtype=torch.int32, device=labels_hi_plane.device)
            cnt = c_i64()
            self._check(self._L.dlv_ccl_boundary_pairs(self._h, _ptr(labels_lo_plane), _ptr(labels_hi_plane), Y, X, _ptr(buf),
                                                       cap, ctypes.byref(cnt)), "dlv_ccl_boundary_pairs")
            if cnt.value <= cap:
                p = buf[:cnt.value].cpu().numpy().view(np.uint32)
                return np.unique(p, axis=0) if len(p) else p.reshape(0, 2)
            cap = int(cnt.value)

    def load_tiff_planes(self, paths, Y, X, slab, threshold=-1, mask=None, nthreads=0):
        """slab: device uint16 (n, SY, SX) <- the n TIFF planes (downsample_and_mask.py:398-414 without the .npy)."""
        n = len(paths)
        arr = (ctypes.c_char_p * n)(*[os.fsencode(p) for p in paths])
        self._after_torch(mask, slab)
        self._check(self._L.dlv_load_tiff_planes(self._h, arr, n, int(Y), int(X), int(threshold), _ptr(mask), _ptr(slab),
                                                 int(slab.shape[1]), int(slab.shape[2]), int(nthreads)), "dlv_load_tiff_planes")

    def paint_boxes(self, mask, shape, boxes, values, outs, chunk_voxels=0):
        """outs[c][box_k] = mask[box_k] * values[k][c] for k in order (blob_highlighter.py:107-124,143-151;
        blob_depthmap.py:198-207).  boxes int64 [n,6] numpy slice bounds, values int [n,nch], outs: list of nch
        (Z,Y,X) uint8 or uint16 arrays / device tensors (all the same dtype), fully overwritten."""
        b = np.ascontiguousarray(boxes, dtype=np.int64).reshape(-1, 6)
        nch = len(outs)
        v = np.ascontiguousarray(values, dtype=np.int64).reshape(len(b), nch)
        eb = {1, 2} & {int(o.element_size()) if hasattr(o, "element_size") else int(o.dtype.itemsize) for o in outs}
        if len(eb) != 1:
            raise ValueError("outputs must all be uint8 or all be uint16")
        shp = (c_i64 * 3)(*[int(s) for s in shape])
        arr = (c_vp * nch)(*[_ptr(o) for o in outs])
        self._after_torch(mask, *outs)
        self._check(self._L.dlv_paint_boxes(self._h, _ptr(mask), shp, b.ctypes.data, v.ctypes.data, len(b), nch, eb.pop(), arr,
                                            int(chunk_voxels)), "dlv_paint_boxes")

    def edt(self, nonzero, sampling, out=None):
        """Exact EDT of a zero-surrounded volume (blob_depthmap.py:174-181): uint8 (Z,Y,X) -> float64 distances."""
        shape = tuple(int(v) for v in nonzero.shape)
        if out is None:
            out = np.empty(shape, dtype=np.float64)
        shp = (c_i64 * 3)(*shape)
        smp = (ctypes.c_double * 3)(*[float(v) for v in sampling])
        self._after_torch(nonzero, out)
        self._check(self._L.dlv_edt(self._h, _ptr(nonzero), shp, smp, _ptr(out)), "dlv_edt")
        return out

    def relabel(self, labels, lut):
        self._after_torch(labels, lut)
        self._check(self._L.dlv_relabel(self._h, _ptr(labels), int(labels.numel()), _ptr(lut), int(lut.numel())), "dlv_relabel")


def tiff_info(path):
    """-> (height, width, bits, compression) of IFD 0 (host only; no GPU needed)."""
    L = load_library()
    h, w, b, c = c_i64(), c_i64(), c_i32(), c_i32()
    if L.dlv_tiff_info(os.fsencode(path), ctypes.byref(h), ctypes.byref(w), ctypes.byref(b), ctypes.byref(c)) != 0:
        raise DlvError(f"dlv_tiff_info: {L.dlv_tiff_last_error().decode()}")
    return h.value, w.value, b.value, c.value


def tiff_read_u16(path):
    """One TIFF plane as a uint16 numpy array (what cv2.imread(path, -1).astype(np.uint16) returns; host only)."""
    L = load_library()
    h, w, _, _ = tiff_info(path)
    out = np.empty((h, w), dtype=np.uint16)
    if L.dlv_tiff_read_u16(os.fsencode(path), out.ctypes.data, h, w) != 0:
        raise DlvError(f"dlv_tiff_read_u16: {L.dlv_tiff_last_error().decode()}")
    return out


def tiff_write_planes(paths, volume, compression=5, nthreads=0):
    """Plane z of ``volume`` ((Z, Y, X) uint8 or uint16, host) -> paths[z]; LZW by default like the reference's
    tifffile.imwrite(..., compression='lzw') (blob_highlighter.py:131-133); planes are compressed on all host threads."""
    L = load_library()
    v = np.ascontiguousarray(volume)
    if v.ndim != 3 or v.dtype not in (np.uint8, np.uint16) or len(paths) != v.shape[0]:
        raise ValueError("volume must be (Z, Y, X) uint8 / uint16 with one path per plane")
    arr = (ctypes.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
    if L.dlv_tiff_write_planes(arr, len(paths), v.ctypes.data, v.shape[1], v.shape[2], v.dtype.itemsize * 8, int(compression),
                               int(nthreads)) != 0:
        raise DlvError(f"dlv_tiff_write_planes: {L.dlv_tiff_write_last_error().decode()}")


def resolve_labels(counts, pairs):
    """dlv_resolve_labels (host only): per-slab component counts + seam pairs -> (uint32 lookup tables, N_global)."""
    L = load_library()
    ns = len(counts)
    cnt = (c_i64 * ns)(*[int(c) for c in counts])
    keep, pp, npp = [], (c_vp * ns)(), (c_i64 * ns)()
    for r in range(ns):
        p = pairs[r] if r < len(pairs) else None
        if r == 0 or p is None or len(p) == 0:
            pp[r], npp[r] = None, 0
            continue
        a = np.ascontiguousarray(p, dtype=np.uint32).reshape(-1, 2)
        keep.append(a)
        pp[r], npp[r] = a.ctypes.data, len(a)
    luts = [np.empty(int(c) + 1, dtype=np.uint32) for c in counts]
    lp = (c_vp * ns)(*[l.ctypes.data for l in luts])
    n = c_i64()
    rc = L.dlv_resolve_labels(ns, cnt, pp, npp, lp, ctypes.byref(n))
    if rc != 0:
        raise DlvError(f"dlv_resolve_labels failed ({rc}): a seam pair names a label outside its slab's 1..N")
    return luts, int(n.value)


def table_merge(tables, luts, z_offsets, n_global, shape_real):
    """dlv_table_merge (host only): per-slab tables + label maps -> global table dict (see slabs.merge_tables)."""
    L = load_library()
    nt = len(tables)
    keep = []                                           # keep the contiguous copies alive during the call
    rows = (c_i64 * nt)()
    lp, cp, sp, bp = (c_vp * nt)(), (c_vp * nt)(), (c_vp * nt)(), (c_vp * nt)()
    for i, (t, lut) in enumerate(zip(tables, luts)):
        if t is None:
            rows[i] = 0
            lp[i] = cp[i] = sp[i] = bp[i] = None
            continue
        a = [np.ascontiguousarray(lut, dtype=np.uint32), np.ascontiguousarray(t["voxel_counts"], dtype=np.uint64),
             np.ascontiguousarray(t["sums"], dtype=np.uint64), np.ascontiguousarray(t["bounding_boxes"], dtype=np.int64)]
        if not (len(a[0]) == len(a[1]) == len(a[2]) == len(a[3])):
            raise ValueError("table / label-map row counts differ")
        keep.append(a)
        rows[i] = len(a[0])
        lp[i], cp[i], sp[i], bp[i] = (x.ctypes.data for x in a)
    n = int(n_global)
    counts = np.empty(n + 1, dtype=np.uint64)
    sums = np.empty((n + 1, 3), dtype=np.uint64)
    bbox = np.empty((n + 1, 6), dtype=np.int64)
    cent = np.empty((n + 1, 3), dtype=np.float64)
    zo = (c_i64 * nt)(*[int(z) for z in z_offsets])
    shp = (c_i64 * 3)(*[int(v) for v in shape_real])
    rc = L.dlv_table_merge(n, nt, rows, lp, cp, sp, bp, zo, shp, counts.ctypes.data, sums.ctypes.data, bbox.ctypes.data,
                           cent.ctypes.data)
    if rc != 0:
        raise DlvError(f"dlv_table_merge failed ({rc}): a label map points outside rows 0..{n}")
    return {"n": n, "voxel_counts": counts, "sums": sums, "bounding_boxes": bbox, "centroids": cent}


def table_csv(centroids, voxel_counts, n):
    """dlv_table_csv (host only): the per-cell CSV text of count_blobs.py:101-114 for table rows 1..n-1 -> str."""
    L = load_library()
    n = int(n)
    cent = np.ascontiguousarray(centroids, dtype=np.float64)
    cnt = np.ascontiguousarray(voxel_counts, dtype=np.uint64)
    if cent.shape != (len(cnt), 3) or len(cnt) < max(n, 1):
        raise ValueError(f"table_csv: centroids {cent.shape} / voxel_counts {cnt.shape} do not hold rows 0..{n - 1}")
    cap = 32 + max(n - 1, 0) * 96                 # typical rows are ~60 bytes; the call reports the size it needs
    while True:
        buf = ctypes.create_string_buffer(cap)
        need = int(L.dlv_table_csv(cent.ctypes.data, cnt.ctypes.data, n, buf, cap))
        if need < 0:
            raise DlvError(f"dlv_table_csv failed ({need})")
        if need <= cap:
            return buf.raw[:need].decode("ascii")
        cap = need


def window_grid(shape_pad, roi, overlap):
    """Per-dimension window start lists (host logic in the library, no GPU needed)."""
    L = load_library()
    sp = (c_i64 * 3)(*[int(v) for v in shape_pad])
    r = (c_i32 * 3)(*[int(v) for v in roi])
    counts = (c_i32 * 3)()
    if L.dlv_window_grid(sp, r, float(overlap), counts, None) != 0:
        raise DlvError("dlv_window_grid: bad arguments")
    buf = (c_i32 * (counts[0] + counts[1] + counts[2]))()
    L.dlv_window_grid(sp, r, float(overlap), counts, buf)
    flat = list(buf)
    return [flat[:counts[0]], flat[counts[0]:counts[0] + counts[1]], flat[counts[0] + counts[1]:]]
