#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference files (TEST INFRASTRUCTURE).

Runs only in the authoring container, where /root/reference exists:

    python oracle/make_golden.py            # writes tests/golden/*.npz, stages the checkpoint

What it does
  1. puts ``oracle/shims`` + ``/root/reference`` on sys.path and imports the
     reference's own ``inference/inference.py``,
     ``inference/sliding_window_inferer.py`` and ``count_blobs.py`` unmodified;
  2. applies the CPU-only monkey-patches (``Tensor.cuda`` -> identity, fixed
     ``mem_get_info`` so inference.py:172-186 yields a chosen sw_batch_size);
  3. runs ``run_inference`` + ``count_blobs`` on small seeded volumes with the
     shipped ``models/inference_weights.tar``;
  4. stores inputs (compressed), the reference's ``binaries.npy`` (bit-packed),
     sub-sampled fp16 averaged logits, N, the stats table and the CSV text;
  5. copies the checkpoint to the git-ignored ``baseline/_ref/`` so the GPU box
     (no /root/reference there) can run the real weights.

The cc3d stand-in is scipy.ndimage.label (see shims/cc3d) - stated wherever
the goldens are used.
"""
import io
import json
import os
import pickle
import shutil
import sys
import tempfile
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DLV_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
LOGIT_STRIDE = (2, 3, 3)

CASES = [
    # name, real shape, roi, seed, tta, load_all_ram, sw_batch
    dict(name="g1_notta", shape=(70, 130, 90), roi=(64, 64, 32), seed=11, tta=False, ram=True, sw=7),
    dict(name="g2_memmap", shape=(40, 100, 70), roi=(32, 48, 32), seed=12, tta=False, ram=False, sw=4),
    dict(name="g3_tta", shape=(40, 100, 70), roi=(32, 48, 32), seed=13, tta=True, ram=True, sw=64),
]


def golden_volume(shape, roi, seed):
    """Small seeded uint16 volume: bright field with blobs, zero corner wedge + zero ball (erosion work)."""
    from oracle import pipeline_ref as P
    rng = np.random.default_rng(seed)
    Z, Y, X = shape
    vol = P.synth_volume(shape, seed)            # ellipsoid brain, 0 outside
    full = np.exp(rng.normal(7.4, 0.35, size=shape))
    vol = np.where(vol > 0, vol, np.clip(full, 1, 65535).astype(np.uint16))   # fill the outside again
    zz, yy, xx = np.ogrid[:Z, :Y, :X]
    vol[(zz + yy + xx) < 0.18 * (Z + Y + X)] = 0                                # zero wedge
    vol[((zz - Z * 0.6) ** 2 + (yy - Y * 0.7) ** 2 + (xx - X * 0.55) ** 2) < 36] = 0   # zero ball
    vol[:, :, X - 4:] = 0                                                       # zero x-slab
    ps = P.padded_shape(shape, roi)
    out = np.zeros(ps, dtype=np.uint16)
    out[:Z, :Y, :X] = vol
    return out


def write_npy_v1_128(path, arr):
    """NPY v1 file whose data start at byte 128 (what np.memmap(..., offset=128) expects)."""
    mm = np.lib.format.open_memmap(path, mode="w+", dtype=arr.dtype, shape=arr.shape)
    assert mm.offset == 128, mm.offset
    mm[...] = arr
    mm.flush()
    del mm


def main():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(HERE, "shims"))
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.set_num_threads(os.cpu_count())

    import inference.inference as ref_inf          # unmodified reference
    import count_blobs as ref_cb                   # unmodified reference
    assert ref_inf.__file__.startswith(REF) and ref_cb.__file__.startswith(REF)
    from path import Path

    weights = os.path.join(REF, "models", "inference_weights.tar")
    stage = os.path.join(ROOT, "baseline", "_ref")
    os.makedirs(stage, exist_ok=True)
    shutil.copyfile(weights, os.path.join(stage, "inference_weights.tar"))

    torch.Tensor.cuda = lambda self, *a, **k: self             # sliding_window_inferer.py:208
    torch.cuda.empty_cache = lambda: None
    torch.cuda.device_count = lambda: 1
    os.makedirs(GOLD, exist_ok=True)

    for case in CASES:
        roi, shape = case["roi"], case["shape"]
        vol = golden_volume(shape, roi, case["seed"])
        per_win_mb = roi[0] * roi[1] * roi[2] * 32 * 45 / (1024 ** 2)
        free_bytes = int(case["sw"] * per_win_mb * (1024 ** 2) / 0.95) + 1024
        torch.cuda.mem_get_info = lambda i=0, fb=free_bytes: (fb, fb)
        with tempfile.TemporaryDirectory() as tmp:
            brain = "brainA"
            nif_dir = os.path.join(tmp, "in", brain, "masked_niftis")
            os.makedirs(nif_dir)
            write_npy_v1_128(os.path.join(nif_dir, "masked_nifti.npy"), vol[None, None])
            out_dir = os.path.join(tmp, "out02")
            post_dir = os.path.join(tmp, "out03") + "/"
            os.makedirs(out_dir)
            settings = {
                "blob_detection": {"window_dimensions": {"window_dim_0": roi[0], "window_dim_1": roi[1],
                                                         "window_dim_2": roi[2]}},
                "postprocessing": {"output_location": post_dir},
                "FLAGS": {"SAVE_ACTIVATED_OUTPUT": True, "LOAD_ALL_RAM": case["ram"]},
            }
            cwd = os.getcwd()
            session = ref_inf.run_inference(
                niftis=[Path(os.path.join(nif_dir, "masked_nifti.npy"))], output_folder=out_dir,
                stack_shape=(1, 1, *shape), model_weights=weights, tta=case["tta"], comment=brain,
                load_all_ram=case["ram"], settings=settings)
            os.chdir(cwd)
            binaries = np.load(os.path.join(session, "binary_segmentations", "binaries.npy"))
            sig = np.load(os.path.join(session, "binary_segmentations", "network_output.npy"))
            files = sorted(os.listdir(session)) + sorted(os.listdir(os.path.join(session, "binary_segmentations")))
            avg = None
            if not case["ram"]:
                avg = np.load(os.path.join(session, "inference_output.npy"))[0, 0]
            ref_cb.count_blobs(settings, out_dir, 0, brain, (1, 1, *shape))
            post_files = sorted(os.listdir(post_dir))
            csv_file = [f for f in post_files if f.endswith(".csv")][0]
            csv = open(os.path.join(post_dir, csv_file)).read()
            lab_file = [f for f in post_files if f.endswith("-cc3d.npy")][0]
            n_comp = int(lab_file.split("-")[1])
            labels = np.load(os.path.join(post_dir, lab_file))
            with open(os.path.join(post_dir, brain + "-stats.pickle"), "rb") as f:
                stats = pickle.load(f)
        s = LOGIT_STRIDE
        gold = dict(
            meta=json.dumps(dict(case, files=files, post_files=post_files, csv_file=csv_file,
                                 logit_stride=s, cc3d_stand_in="scipy.ndimage.label 3x3x3")),
            volume=vol,
            volume_crc=np.uint32(zlib.crc32(vol.tobytes())),
            binaries_packed=np.packbits(binaries.reshape(-1)),
            binaries_sum=np.int64(binaries.sum()),
            sigmoid_sub=sig[::s[0], ::s[1], ::s[2]].astype(np.float32),
            labels_crc=np.uint32(zlib.crc32(np.ascontiguousarray(labels, dtype=np.uint32).tobytes())),
            n_components=np.int64(n_comp),
            voxel_counts=np.asarray(stats["voxel_counts"], dtype=np.uint64),
            bounding_boxes=np.asarray(stats["bounding_boxes"], dtype=np.int64),
            centroids=np.asarray(stats["centroids"], dtype=np.float64),
            csv=np.frombuffer(csv.encode(), dtype=np.uint8),
        )
        if avg is not None:
            gold["avg_logits_sub"] = avg[::s[0], ::s[1], ::s[2]]
        np.savez_compressed(os.path.join(GOLD, case["name"] + ".npz"), **gold)
        print(case["name"], "N =", n_comp, "fg =", int(binaries.sum()), "csv rows =", csv.count("\n") - 1,
              "files:", files, post_files)


if __name__ == "__main__":
    main()
