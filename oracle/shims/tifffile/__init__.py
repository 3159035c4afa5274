"""tifffile stand-in (TEST INFRASTRUCTURE): only the two calls the reference's visualisation files make
(blob_highlighter.py:131-133,160; blob_depthmap.py:29,167), backed by OpenCV's libtiff (LZW by default)."""
import cv2
import numpy as np


def imwrite(path, data, compression=None, **_):
    if not cv2.imwrite(str(path), np.ascontiguousarray(data)):
        raise IOError(f"cv2.imwrite failed for {path}")


def imread(path, **_):
    if isinstance(path, (list, tuple)):
        return np.stack([imread(p) for p in path])
    ok, pages = cv2.imreadmulti(str(path), flags=cv2.IMREAD_UNCHANGED)
    if not ok or not pages:
        raise IOError(f"cannot read {path}")
    return pages[0] if len(pages) == 1 else np.stack(pages)
