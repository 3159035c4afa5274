"""Host-side TIFF plane reader (dlv_tiff_read_u16, no GPU) against cv2.imread(path, -1) - the reader the reference
uses for raw planes (downsample_and_mask.py:27,400) - over the encodings cv2 / Fiji-style writers produce."""
import os
import struct
import zlib

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from delivr_cfos_b200 import DlvError  # noqa: E402
from delivr_cfos_b200._lib import tiff_info, tiff_read_u16  # noqa: E402


def _images():
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:301, 0:517]
    smooth = (2000 + 1500 * np.sin(yy / 17.0) * np.cos(xx / 23.0)).astype(np.uint16)      # long LZW strings, width changes
    noisy = rng.integers(0, 65536, (301, 517)).astype(np.uint16)                           # table fills up -> clear codes
    sparse = np.zeros((64, 1000), np.uint16); sparse[10:20, 100:900] = 40000               # runs: KwKwK case
    tiny = np.array([[7]], np.uint16)
    return {"smooth": smooth, "noisy": noisy, "sparse": sparse, "tiny": tiny}


@pytest.mark.parametrize("name", ["smooth", "noisy", "sparse", "tiny"])
@pytest.mark.parametrize("comp", [1, 5, 8, 32773, 32946])
@pytest.mark.parametrize("rows_per_strip", [0, 7])
def test_reader_matches_cv2(tmp_path, name, comp, rows_per_strip):
    img = _images()[name]
    p = str(tmp_path / f"{name}_{comp}_{rows_per_strip}.tif")
    params = [cv2.IMWRITE_TIFF_COMPRESSION, comp]
    if rows_per_strip:
        params += [cv2.IMWRITE_TIFF_ROWSPERSTRIP, rows_per_strip]
    assert cv2.imwrite(p, img, params)
    ref = cv2.imread(p, -1)
    assert ref.dtype == np.uint16 and np.array_equal(ref, img)
    h, w, bits, c = tiff_info(p)
    assert (h, w, bits) == (img.shape[0], img.shape[1], 16) and c == comp
    assert np.array_equal(tiff_read_u16(p), ref)


def test_reader_8bit_matches_cv2_astype(tmp_path):
    img = (np.random.default_rng(1).integers(0, 256, (77, 130))).astype(np.uint8)
    for comp in (1, 5):
        p = str(tmp_path / f"u8_{comp}.tif")
        cv2.imwrite(p, img, [cv2.IMWRITE_TIFF_COMPRESSION, comp])
        assert np.array_equal(tiff_read_u16(p), cv2.imread(p, -1).astype(np.uint16))     # downsample_and_mask.py:411


def _write_raw_tiff(path, img, big_endian, compress=None):
    """Minimal classic-TIFF writer (one strip) for byte orders / encodings cv2 cannot be asked to write."""
    bo = ">" if big_endian else "<"
    data = img.astype(bo + "u2").tobytes()
    comp = 1
    if compress == "zlib":
        data, comp = zlib.compress(data), 8
    h, w = img.shape
    tags = [(256, 4, 1, w), (257, 4, 1, h), (258, 3, 1, 16), (259, 3, 1, comp), (262, 3, 1, 1), (273, 4, 1, 8),
            (277, 3, 1, 1), (278, 4, 1, h), (279, 4, 1, len(data))]
    ifd_off = 8 + len(data) + (len(data) & 1)
    out = (b"MM" if big_endian else b"II") + struct.pack(bo + "HI", 42, ifd_off) + data + (b"\0" if len(data) & 1 else b"")
    out += struct.pack(bo + "H", len(tags))
    for tag, typ, cnt, val in tags:
        out += struct.pack(bo + "HHI", tag, typ, cnt) + (struct.pack(bo + "HH", val, 0) if typ == 3 else struct.pack(bo + "I", val))
    out += struct.pack(bo + "I", 0)
    open(path, "wb").write(out)


@pytest.mark.parametrize("big_endian", [False, True])
@pytest.mark.parametrize("compress", [None, "zlib"])
def test_reader_byte_orders(tmp_path, big_endian, compress):
    img = np.random.default_rng(2).integers(0, 65536, (40, 33)).astype(np.uint16)
    p = str(tmp_path / "raw.tif")
    _write_raw_tiff(p, img, big_endian, compress)
    assert np.array_equal(cv2.imread(p, -1), img)          # cv2 agrees the file is valid
    assert np.array_equal(tiff_read_u16(p), img)


def test_reader_rejects_what_it_cannot_decode(tmp_path):
    rgb = np.zeros((8, 8, 3), np.uint8)
    p = str(tmp_path / "rgb.tif")
    cv2.imwrite(p, rgb)
    with pytest.raises(DlvError, match="single-channel"):
        tiff_read_u16(p)
    f32 = np.zeros((8, 8), np.float32)
    p = str(tmp_path / "f32.tif")
    cv2.imwrite(p, f32)
    with pytest.raises(DlvError):
        tiff_read_u16(p)
    p = str(tmp_path / "garbage.tif")
    open(p, "wb").write(b"not a tiff at all")
    with pytest.raises(DlvError, match="TIFF"):
        tiff_read_u16(p)
    img = np.arange(40 * 33, dtype=np.uint16).reshape(40, 33)
    p = str(tmp_path / "trunc.tif")
    cv2.imwrite(p, img, [cv2.IMWRITE_TIFF_COMPRESSION, 5])
    raw = bytearray(open(p, "rb").read())
    raw[20:60] = b"\xff" * 40                               # corrupt the LZW stream
    open(p, "wb").write(bytes(raw))
    with pytest.raises(DlvError):
        tiff_read_u16(p)
    with pytest.raises(DlvError, match="cannot open"):
        tiff_read_u16(str(tmp_path / "missing.tif"))


def test_get_real_size_and_plane_order(tmp_path):
    from delivr_cfos_b200 import tiff_planes
    for i in (2, 0, 1):
        cv2.imwrite(str(tmp_path / f"plane_Z{i:04d}.tif"), np.full((12, 20), i, np.uint16))
    open(tmp_path / "notes.txt", "w").write("x")
    assert tiff_planes.get_real_size(str(tmp_path)) == (3, 12, 20)
    assert [os.path.basename(p) for p in tiff_planes.list_planes(str(tmp_path))] == [f"plane_Z{i:04d}.tif" for i in range(3)]
    assert tiff_planes.padded_shape((3, 12, 20), (4, 8, 16)) == (4, 16, 32)
