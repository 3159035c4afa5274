"""dlv_table_csv (host code of the library) against the text pandas / Python produce for the reference's per-cell table
(count_blobs.py:101-114): byte equality, including Python's float repr rules."""
import numpy as np

from delivr_cfos_b200.count_blobs import csv_text


def python_text(stats, n):
    cent = np.asarray(stats["centroids"])[1:n].tolist()
    cnt = np.asarray(stats["voxel_counts"])[1:n].tolist()
    return "".join([",Blob,Coords,Size\n"] + [f'0,{i},"{c}",{k}\n' for i, (c, k) in enumerate(zip(cent, cnt), 1)])


def test_rows_like_a_component_table_threaded():
    rng = np.random.default_rng(0)
    n = 120000                                     # above the library's single-thread limit
    cnt = rng.integers(1, 5000, n + 1).astype(np.uint64)
    sums = rng.integers(0, 4000, (n + 1, 3)).astype(np.uint64) * cnt[:, None] + rng.integers(0, 5000, (n + 1, 3)).astype(np.uint64)
    st = {"centroids": sums.astype(np.float64) / cnt.astype(np.float64)[:, None], "voxel_counts": cnt}
    text = csv_text(st, n)
    assert text == python_text(st, n)
    assert text.count("\n") == n                   # header + rows 1 .. n-1: the last component is not listed


def test_python_float_repr_rules():
    vals = np.array([0.0, 1.0, 1e-4, 9.999e-5, 1e-5, 1.5e-7, 123456789012345.6, 1e15, 1e16, 1.2345e16, 9999999999999998.0, 1e22,
                     1e-300, 5e-324, 0.1, 1 / 3, 2 / 3, 1e3, float("nan"), float("inf"), -1.5, -0.0, -1e-7,
                     1.7976931348623157e308, 4503599627370496.5, 0.30000000000000004, 100.0, 100.00000000000001], dtype=np.float64)
    m = len(vals)
    cent = np.concatenate([np.zeros((1, 3)), np.stack([vals, np.roll(vals, 1), np.roll(vals, 2)], 1), np.zeros((1, 3))])
    st = {"centroids": cent, "voxel_counts": np.arange(m + 2, dtype=np.uint64) * np.uint64(12345678901234)}
    assert csv_text(st, m + 1) == python_text(st, m + 1)


def test_random_doubles_of_every_magnitude():
    rng = np.random.default_rng(5)
    x = rng.integers(0, 2 ** 63 - 1, (60000, 3), dtype=np.int64).view(np.float64)
    x = np.where(np.isfinite(x), x, 1.0)
    st = {"centroids": np.concatenate([np.zeros((1, 3)), x]), "voxel_counts": np.ones(len(x) + 1, np.uint64)}
    assert csv_text(st, len(x) + 1) == python_text(st, len(x) + 1)


def test_empty_tables():
    z = {"centroids": np.full((2, 3), np.nan), "voxel_counts": np.zeros(2, np.uint64)}
    assert csv_text(z, 0) == ",Blob,Coords,Size\n"
    assert csv_text(z, 1) == ",Blob,Coords,Size\n"      # N = 1: range(1, 1) is empty
