"""TIFF planes -> device volume (dlv_load_tiff_planes) against a numpy restatement of the reference's masking loop
(downsample_and_mask.py:391-414, with cv2.imread as the reader), and run_inference fed from it."""
import os

import numpy as np
import pytest
import torch

from gpu_common import ctx_with

cv2 = pytest.importorskip("cv2")
pytestmark = pytest.mark.gpu


def _reference_masked_nifti(raw_location, crop_size, threshold=None, mask=None):
    """downsample_and_mask.py:391-414 restated (array kept in memory instead of an .npy memmap)."""
    items = sorted([x for x in os.listdir(raw_location) if ".tif" in x])
    first = cv2.imread(os.path.join(raw_location, items[0]), -1)
    raw_shape = (len(items),) + first.shape
    pad = [int(np.ceil(d / c) * c) for d, c in zip(raw_shape, crop_size)]
    out = np.zeros((1, 1, *pad), dtype=np.uint16)
    for i, item in enumerate(items):
        img = cv2.imread(os.path.join(raw_location, item), -1)
        if mask is not None:
            img *= mask[i, :, :]
        elif threshold is not None:
            img[img < int(threshold)] = 0
        out[0, 0, i, 0:raw_shape[1], 0:raw_shape[2]] = img.astype(np.uint16)
    return out, raw_shape


def _write_planes(folder, vol, comp):
    os.makedirs(folder, exist_ok=True)
    for z in range(vol.shape[0]):
        cv2.imwrite(os.path.join(folder, f"raw_Z{z:04d}.tif"), vol[z], [cv2.IMWRITE_TIFF_COMPRESSION, comp])


@pytest.mark.parametrize("shape,crop,comp,mode", [((21, 50, 70), (16, 32, 32), 5, "threshold"), ((9, 33, 41), (8, 16, 16), 1, "mask"),
                                                  ((40, 64, 64), (32, 32, 32), 8, "none"), ((5, 100, 136), (4, 32, 8), 5, "threshold")])
def test_load_planes_equals_reference_loop(tmp_path, shape, crop, comp, mode):
    from delivr_cfos_b200 import Context, tiff_planes
    rng = np.random.default_rng(11)
    vol = rng.integers(0, 3000, shape).astype(np.uint16)
    raw = str(tmp_path / "raw")
    _write_planes(raw, vol, comp)
    mask = (rng.random(shape) < 0.7).astype(np.uint8) if mode == "mask" else None
    thr = 900 if mode == "threshold" else None
    ref, raw_shape = _reference_masked_nifti(raw, crop, thr, mask)
    ctx = Context(0)
    for nthreads in (1, 3):
        got, real = tiff_planes.load_masked_volume(ctx, raw, crop, threshold=thr, mask=mask, nthreads=nthreads)
        assert real == raw_shape == shape
        assert np.array_equal(got.cpu().numpy(), ref[0, 0])
    # z-sharded load: a slab of planes equals the same planes of the full volume
    z0, z1 = 2, min(shape[0], 7)
    part, _ = tiff_planes.load_masked_volume(ctx, raw, crop, threshold=thr, mask=mask, z_range=(z0, z1))
    assert np.array_equal(part.cpu().numpy(), ref[0, 0, z0:z1])


def test_run_inference_from_tiff_planes(tmp_path):
    """binaries.npy from the TIFF-fed path == binaries.npy from the masked_nifti.npy path."""
    from delivr_cfos_b200 import tiff_planes
    from delivr_cfos_b200.inference import inference as inf
    from delivr_cfos_b200.inference.sliding_window_inferer import DelivrNet
    from oracle import pipeline_ref as P
    _, sd, _ = ctx_with("random")
    net = DelivrNet(state_dict=sd)
    shape, roi = (70, 90, 80), (32, 48, 32)
    vol = P.synth_volume(shape, 21)
    vol = np.where(vol == 0, 1, vol).astype(np.uint16)
    vol[:2] = 0
    raw = str(tmp_path / "raw")
    _write_planes(raw, vol, 5)
    settings = {"blob_detection": {"window_dimensions": {"window_dim_0": roi[0], "window_dim_1": roi[1], "window_dim_2": roi[2]}},
                "FLAGS": {"SAVE_ACTIVATED_OUTPUT": False, "LOAD_ALL_RAM": True}}
    ref, raw_shape = _reference_masked_nifti(raw, roi, threshold=0)
    nif = str(tmp_path / "masked_nifti.npy")
    mm = np.lib.format.open_memmap(nif, mode="w+", dtype=np.uint16, shape=ref.shape)
    mm[...] = ref
    mm.flush()
    s1 = inf.run_inference([nif], str(tmp_path / "a"), (1, 1) + shape, comment="b", load_all_ram=True, settings=settings, _net=net)
    dvol, real = tiff_planes.load_masked_volume(net.ctx, raw, roi, threshold=0)
    s2 = inf.run_inference(None, str(tmp_path / "b"), (1, 1) + real, comment="b", load_all_ram=True, settings=settings, _net=net,
                           volume=dvol)
    b1 = np.load(os.path.join(s1, "binary_segmentations", "binaries.npy"))
    b2 = np.load(os.path.join(s2, "binary_segmentations", "binaries.npy"))
    assert b1.shape == shape and b1.sum() > 0 and np.array_equal(b1, b2)
