"""ctypes wrapper for the plain-C oracle ``ccl_ref.c`` (TEST INFRASTRUCTURE)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.run(["make", "-C", _HERE, "--no-print-directory"], check=True, stdout=subprocess.DEVNULL)


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libdlvref.so")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_HERE, "ccl_ref.c")):
            build()
        L = ctypes.CDLL(so)
        i64, vp = ctypes.c_int64, ctypes.c_void_p
        L.dlvref_ccl26.restype = i64
        L.dlvref_ccl26.argtypes = [vp, i64, i64, i64, vp]
        L.dlvref_stats.restype = None
        L.dlvref_stats.argtypes = [vp, i64, i64, i64, i64, vp, vp, vp]
        L.dlvref_erode6.restype = ctypes.c_int
        L.dlvref_erode6.argtypes = [vp, i64, i64, i64, ctypes.c_int, vp]
        _LIB = L
    return _LIB


def connected_components26(mask):
    """-> (labels uint32 (Z,Y,X), N).  cc3d.connected_components semantics (count_blobs.py:61)."""
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    Z, Y, X = m.shape
    lab = np.empty((Z, Y, X), dtype=np.uint32)
    n = _lib().dlvref_ccl26(m.ctypes.data, Z, Y, X, lab.ctypes.data)
    if n < 0:
        raise MemoryError("dlvref_ccl26 failed")
    return lab, int(n)


def statistics(labels, N):
    """-> dict like cc3d.statistics(no_slice_conversion=True) plus exact integer ``sums`` (count_blobs.py:85)."""
    lab = np.ascontiguousarray(labels, dtype=np.uint32)
    Z, Y, X = lab.shape
    counts = np.empty(N + 1, dtype=np.uint64)
    sums = np.empty((N + 1, 3), dtype=np.uint64)
    bbox = np.empty((N + 1, 6), dtype=np.int64)
    _lib().dlvref_stats(lab.ctypes.data, Z, Y, X, N, counts.ctypes.data, sums.ctypes.data, bbox.ctypes.data)
    with np.errstate(invalid="ignore", divide="ignore"):
        cent = sums.astype(np.float64) / counts.astype(np.float64)[:, None]
    return {"voxel_counts": counts, "bounding_boxes": bbox, "centroids": cent, "sums": sums}


def erode6(mask, iterations=30):
    """scipy.ndimage.binary_erosion(mask, iterations=iterations, border_value=1) (inference.py:82)."""
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    Z, Y, X = m.shape
    out = np.empty_like(m)
    if _lib().dlvref_erode6(m.ctypes.data, Z, Y, X, int(iterations), out.ctypes.data) != 0:
        raise MemoryError("dlvref_erode6 failed")
    return out
