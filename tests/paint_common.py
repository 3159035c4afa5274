"""Helpers shared by the painter tests."""
import io
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_paint_golden():
    """tests/golden/p1_highlight.npz: inputs + outputs of the UNMODIFIED reference blob_highlighter
    (oracle/make_golden_paint.py)."""
    import pandas as pd
    g = dict(np.load(os.path.join(GOLD, "p1_highlight.npz")))
    shape = tuple(int(s) for s in g["shape"])
    g["shape"] = shape
    g["mask"] = np.unpackbits(g["bits"])[: int(np.prod(shape))].reshape(shape)
    g["csv_bytes"] = bytes(g["csv"])
    cells = pd.read_csv(io.BytesIO(g["csv_bytes"]), index_col=0)
    g["cells"] = cells.loc[cells["acronym"] != "bgr"]            # blob_highlighter.py:75
    g["stack_shape"] = (1, 1) + shape
    return g


def random_paint_case(shape, n, seed, big=False, vmax=70000):
    """Random mask (values 0..3), n boxes in random order (overlapping, some empty, some beyond the array end)."""
    rng = np.random.default_rng(seed)
    Z, Y, X = shape
    mask = (rng.random(shape) < 0.3).astype(np.uint8) * rng.integers(1, 4, size=shape, dtype=np.uint8)
    lo = np.stack([rng.integers(0, s, n) for s in shape], 1)
    ext = np.stack([rng.integers(0, max(2, s // 3), n) for s in shape], 1)
    boxes = np.empty((n, 6), dtype=np.int64)
    boxes[:, 0::2] = lo
    boxes[:, 1::2] = lo + ext                                       # may exceed the array end (numpy clips); ext 0 = empty
    if big and n:
        boxes[n // 2] = [0, Z, 0, Y, 0, X]                          # whole volume: the "whole grid" path
    values = rng.integers(0, vmax, size=(n, 3))
    return mask, boxes, values
