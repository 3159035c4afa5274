"""Drives the reference's UNMODIFIED ``__main__.py`` (run with runpy from /root/reference, or from the copy that
``__graft_entry__.build()`` stages under the git-ignored baseline/_ref/reference/ for the GPU box) with the two import
lines of INTEGRATION.md section 1 redirected to the drop-ins: ``from inference import inference`` and
``from count_blobs import count_blobs`` resolve to delivr_cfos_b200's mirrors.  Everything else __main__ imports
(mask/downsample, atlas alignment, region assignment, visualisation) is outside the hot path and stubbed."""
import json
import os
import runpy
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_main_path():
    for p in ("/root/reference/__main__.py", os.path.join(ROOT, "baseline", "_ref", "reference", "__main__.py")):
        if os.path.exists(p):
            return p
    return None


def make_tree(tmp, volume, roi, brain="brainA", tta=False, save_act=False, load_all_ram=True):
    """Input tree as the mask stage leaves it: raw TIFF planes (only their count / size is read, get_real_size),
    <mask out>/<brain>/masked_niftis/masked_nifti.npy (uint16 (1,1,Zp,Yp,Xp), 128-byte header) and a config."""
    from delivr_cfos_b200._lib import tiff_write_planes
    Z, Y, X = volume.shape
    raw = os.path.join(tmp, "raw", brain)
    os.makedirs(raw)
    tiff_write_planes([os.path.join(raw, f"plane_{z:04d}.tif") for z in range(Z)], np.ascontiguousarray(volume), compression=1)
    pad = tuple(int(np.ceil(d / r) * r) for d, r in zip(volume.shape, roi))
    mdir = os.path.join(tmp, "out", "01_mask", brain, "masked_niftis")
    os.makedirs(mdir)
    m = np.lib.format.open_memmap(os.path.join(mdir, "masked_nifti.npy"), mode="w+", dtype=np.uint16, shape=(1, 1) + pad)
    m[0, 0, :Z, :Y, :X] = volume
    m.flush()
    del m
    out = os.path.join(tmp, "out")
    wp = lambda n: os.path.join(out, n, "output") + "/"
    settings = {
        "raw_location": os.path.join(tmp, "raw") + "/", "output_location": out + "/",
        "mask_detection": {"output_location": os.path.join(out, "01_mask") + "/"},
        "blob_detection": {"input_location": "", "model_location": os.path.join(ROOT, "baseline", "_ref", "inference_weights.tar"),
                           "output_location": wp("02_blob_detection"),
                           "window_dimensions": {"window_dim_0": roi[0], "window_dim_1": roi[1], "window_dim_2": roi[2]}},
        "postprocessing": {"input_location": wp("02_blob_detection"), "output_location": wp("03_postprocessing"), "min_size": -1, "max_size": -1},
        "atlas_alignment": {"output_location": wp("04_atlas_alignment"), "collection_folder": os.path.join(out, "04_atlas_alignment", "collection") + "/"},
        "region_assignment": {"output_location": wp("05_region_assignment")},
        "visualization": {"output_location": wp("06_visualization")},
        "FLAGS": {"ABSPATHS": True, "LOAD_ALL_RAM": load_all_ram, "TEST_TIME_AUGMENTATION": tta, "MASK_DOWNSAMPLE": False,
                  "BLOB_DETECTION": True, "POSTPROCESSING": True, "ATLAS_ALIGNMENT": False, "REGION_ASSIGNMENT": False,
                  "VISUALIZATION": False, "SAVE_MASK_OUTPUT": True, "SAVE_NETWORK_OUTPUT": True, "SAVE_ACTIVATED_OUTPUT": save_act,
                  "SAVE_POSTPROCESSING_OUTPUT": True, "SAVE_ATLAS_OUTPUT": True},
    }
    cfg = os.path.join(tmp, "config.json")
    with open(cfg, "w") as f:
        json.dump(settings, f)
    return cfg, settings


def run_reference_main(config_path):
    """runpy the reference's __main__.py with the import swap in place."""
    main_py = reference_main_path()
    assert main_py is not None, "reference __main__.py neither at /root/reference nor staged under baseline/_ref/reference"
    import delivr_cfos_b200.count_blobs as cb
    import delivr_cfos_b200.inference.inference as inf
    from delivr_cfos_b200 import tiff_planes
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))        # `path` (Path: .dirs() / .files() / .name / .parent)
    saved = {k: sys.modules.get(k) for k in ("inference", "count_blobs", "downsample", "downsample.downsample_and_mask",
                                             "automate_mBrainaligner", "cells_to_atlas", "blob_highlighter")}
    pkg = types.ModuleType("inference")
    pkg.inference = inf                                              # __main__.py:8   from inference import inference
    sys.modules["inference"] = pkg
    sys.modules["count_blobs"] = cb                                  # __main__.py:9   from count_blobs import count_blobs
    ds = types.ModuleType("downsample")
    dm = types.ModuleType("downsample.downsample_and_mask")
    dm.get_real_size = tiff_planes.get_real_size                     # mirror of downsample_and_mask.py:25-30 (row a11)
    dm.downsample_mask = lambda *a, **k: None
    ds.downsample_and_mask = dm
    sys.modules["downsample"], sys.modules["downsample.downsample_and_mask"] = ds, dm
    for name, attr in (("automate_mBrainaligner", "run_mbrainaligner_and_swc_reg"), ("cells_to_atlas", "map_cells_to_atlas"),
                       ("blob_highlighter", "blob_highlighter")):
        mod = types.ModuleType(name)
        setattr(mod, attr, lambda *a, **k: None)
        sys.modules[name] = mod
    argv = sys.argv
    sys.argv = [main_py, config_path]
    try:
        runpy.run_path(main_py, run_name="__main__")
    finally:
        sys.argv = argv
        sys.path.remove(os.path.join(ROOT, "oracle", "shims"))
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
