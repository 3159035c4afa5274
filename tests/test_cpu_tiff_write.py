"""dlv_tiff_write_planes (host only): what it writes must read back identically through an independent decoder
(OpenCV's libtiff - the reader the reference's own pipeline uses for planes) and through the library's own reader."""
import os

import cv2
import numpy as np
import pytest

from delivr_cfos_b200._lib import DlvError, tiff_info, tiff_read_u16, tiff_write_planes


def _roundtrip(tmp_path, vol, comp, tag):
    paths = [str(tmp_path / f"{tag}_{i:04d}.tif") for i in range(vol.shape[0])]
    tiff_write_planes(paths, vol, compression=comp)
    for i, p in enumerate(paths):
        a = cv2.imread(p, cv2.IMREAD_UNCHANGED)
        assert a is not None and a.dtype == vol.dtype and np.array_equal(a, vol[i]), (tag, i)
        assert np.array_equal(tiff_read_u16(p), vol[i].astype(np.uint16))
        h, w, bits, c = tiff_info(p)
        assert (h, w, bits, c) == (vol.shape[1], vol.shape[2], vol.dtype.itemsize * 8, comp)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
@pytest.mark.parametrize("comp", [1, 5, 8])
@pytest.mark.parametrize("shape", [(2, 1, 1), (3, 7, 5), (2, 300, 517), (1, 700, 900), (2, 33, 4099)])
def test_written_planes_read_back(tmp_path, dtype, comp, shape):
    rng = np.random.default_rng(shape[2])
    hi = np.iinfo(dtype).max
    for kind, v in (("zero", np.zeros(shape, dtype)),
                    ("sparse", ((rng.random(shape) < 0.02) * rng.integers(1, 250, shape)).astype(dtype)),
                    ("noise", rng.integers(0, hi, shape).astype(dtype)),
                    ("ramp", (np.arange(int(np.prod(shape))).reshape(shape) % hi).astype(dtype))):
        _roundtrip(tmp_path, v, comp, kind)


def test_lzw_code_width_and_table_reset_boundaries(tmp_path):
    """Stream lengths around every 9->10->11->12 bit switch and the 4094-entry table reset, low and high entropy."""
    rng = np.random.default_rng(0)
    for n in list(range(1, 70)) + list(range(500, 530)) + list(range(3990, 4110, 3)) + [8191, 8192, 70000]:
        _roundtrip(tmp_path, rng.integers(0, 7, (1, 1, n)).astype(np.uint8), 5, f"lo{n}")
        _roundtrip(tmp_path, rng.integers(0, 256, (1, 1, n)).astype(np.uint8), 5, f"hi{n}")


def test_writer_rejects_bad_input(tmp_path):
    with pytest.raises(ValueError):
        tiff_write_planes([str(tmp_path / "a.tif")], np.zeros((1, 4, 4), dtype=np.float32))
    with pytest.raises(ValueError):
        tiff_write_planes([str(tmp_path / "a.tif")], np.zeros((2, 4, 4), dtype=np.uint8))
    with pytest.raises(DlvError):
        tiff_write_planes([str(tmp_path / "no_such_dir" / "a.tif")], np.zeros((1, 4, 4), dtype=np.uint8))
    with pytest.raises(DlvError):
        tiff_write_planes([str(tmp_path / "a.tif")], np.zeros((1, 4, 4), dtype=np.uint8), compression=7)
