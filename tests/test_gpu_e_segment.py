"""End-to-end parity through the reference-facing API: run_inference + count_blobs vs the oracle and the goldens.

Logit tolerance (bf16 operands, fp32 accumulation; compared where the eroded mask is 1):
max |logit - ref| <= 1.0 + 0.02 |ref|; >= 99.9 % of binaries.npy voxels agree; the component
table / CSV are bit-exact given the same binaries.
"""
import json
import os
import pickle

import numpy as np
import pytest
import torch

from conftest import weights_path
from helpers import load_golden
from oracle import pipeline_ref as P, unet_ref

pytestmark = pytest.mark.gpu


def _write_input(tmp, brain, vol):
    d = os.path.join(tmp, "in", brain, "masked_niftis")
    os.makedirs(d)
    path = os.path.join(d, "masked_nifti.npy")
    mm = np.lib.format.open_memmap(path, mode="w+", dtype=np.uint16, shape=(1, 1) + vol.shape)
    assert mm.offset == 128
    mm[0, 0] = vol
    mm.flush()
    return path


def _settings(tmp, roi, ram=True, save=True):
    return {"blob_detection": {"window_dimensions": {"window_dim_0": roi[0], "window_dim_1": roi[1], "window_dim_2": roi[2]}},
            "postprocessing": {"output_location": os.path.join(tmp, "out03") + "/"},
            "FLAGS": {"SAVE_ACTIVATED_OUTPUT": save, "LOAD_ALL_RAM": ram}}


def _run(tmp, vol, shape, roi, net, tta=False, ram=True):
    from delivr_cfos_b200.inference import inference as inf
    from delivr_cfos_b200.count_blobs import count_blobs
    brain = "brainA"
    nif = _write_input(tmp, brain, vol)
    out02 = os.path.join(tmp, "out02")
    os.makedirs(out02)
    settings = _settings(tmp, roi, ram=ram)
    session = inf.run_inference([nif], out02, (1, 1) + tuple(shape), model_weights="unused", tta=tta, comment=brain,
                                load_all_ram=ram, settings=settings, _net=net)
    count_blobs(settings, out02, 0, brain, (1, 1) + tuple(shape))
    return session, settings


@pytest.mark.parametrize("name", ["g2_memmap", "g1_notta", "g3_tta"])
def test_golden_from_unmodified_reference(name, tmp_path):
    wp = weights_path()
    if wp is None:
        pytest.skip("shipped checkpoint not staged")
    from delivr_cfos_b200.inference.sliding_window_inferer import DelivrNet
    g = load_golden(name)
    m = g["meta"]
    net = DelivrNet(checkpoint_path=wp)
    session, settings = _run(str(tmp_path), g["volume"], m["shape"], m["roi"], net, tta=m["tta"], ram=m["ram"])
    files = sorted(os.listdir(session)) + sorted(os.listdir(os.path.join(session, "binary_segmentations")))
    assert files == m["files"], (files, m["files"])
    b = np.load(os.path.join(session, "binary_segmentations", "binaries.npy"))
    assert b.dtype == np.uint8 and b.shape == tuple(m["shape"])
    agree = (b == g["binaries"]).mean()
    print(name, "binaries agreement", agree, "fg", int(b.sum()), "ref fg", int(g["binaries_sum"]))
    assert agree >= 0.999
    sig = np.load(os.path.join(session, "binary_segmentations", "network_output.npy"))
    s = m["logit_stride"]
    mask = P.ccl_ref.erode6((g["volume"][:m["shape"][0], :m["shape"][1], :m["shape"][2]] > 0).astype(np.uint8), 30)
    sub = (slice(None, None, s[0]), slice(None, None, s[1]), slice(None, None, s[2]))
    d = np.abs(sig[sub] - g["sigmoid_sub"])[mask[sub] > 0]
    print(name, "sigmoid max diff on eroded mask", d.max() if d.size else 0.0)
    assert d.size == 0 or d.max() < 0.2
    post = settings["postprocessing"]["output_location"]
    pf = sorted(os.listdir(post))
    assert [f.split("-")[0] if "cc3d" in f else f for f in pf] == [f.split("-")[0] if "cc3d" in f else f for f in m["post_files"]]
    # the table must be bit-exact for OUR binaries: check against the oracle run on the same mask
    lab, n, st = P.blob_table(b)
    csv = open(os.path.join(post, m["csv_file"])).read()
    assert csv == P.csv_text(st, n)
    labf = [f for f in pf if f.endswith("-cc3d.npy")][0]
    assert int(labf.split("-")[1]) == n
    assert np.array_equal(np.load(os.path.join(post, labf)), lab)
    with open(os.path.join(post, "brainA-stats.pickle"), "rb") as f:
        stp = pickle.load(f)
    assert np.array_equal(stp["voxel_counts"], st["voxel_counts"])
    assert np.array_equal(stp["bounding_boxes"], st["bounding_boxes"])
    assert np.array_equal(stp["centroids"], st["centroids"], equal_nan=True)
    if name == "g3_tta":
        # this golden is reproduced voxel for voxel (963 / 963 foreground, every run since round 1), so the whole byte path
        # - binaries.npy, then the CSV the unmodified reference wrote - is pinned unconditionally here
        assert np.array_equal(b, g["binaries"])
    if np.array_equal(b, g["binaries"]):
        assert csv == g["csv"]


@pytest.mark.parametrize("shape,roi,tta", [((40, 100, 70), (32, 48, 32), False), ((50, 90, 60), (32, 32, 32), True)])
def test_against_oracle_random_weights(shape, roi, tta, tmp_path):
    """No checkpoint needed: same seeded random weights in the CUDA library and in the torch-fp32 oracle."""
    from delivr_cfos_b200.inference.sliding_window_inferer import DelivrNet
    sd = unet_ref.random_state_dict(3)
    net = DelivrNet(state_dict=sd)
    onet = unet_ref.BasicUNet(dropout=0.1)
    onet.load_state_dict(unet_ref.strip_module_prefix(sd), strict=True)
    onet = onet.eval().cuda()
    vol = P.synth_volume(shape, 31, roi=roi)
    vol[:shape[0], :shape[1], :shape[2]][vol[:shape[0], :shape[1], :shape[2]] == 0] = 500   # keep the eroded mask non-empty
    vol[:8] = 0
    pred = lambda t: onet(t.cuda()).cpu()
    avg = P.infer_average(vol, roi, 0.5, pred, 6, tta=tta)
    ref_b, ref_sig = P.create_binaries(avg, vol, shape, 0.5, return_sigmoid=True)
    session, settings = _run(str(tmp_path), vol, shape, roi, net, tta=tta, ram=False)
    b = np.load(os.path.join(session, "binary_segmentations", "binaries.npy"))
    agree = (b == ref_b).mean()
    mine = np.load(os.path.join(session, "inference_output.npy"))[0, 0].astype(np.float32)
    mask = P.ccl_ref.erode6((vol[:shape[0], :shape[1], :shape[2]] > 0).astype(np.uint8), 30) > 0
    ref = avg[:shape[0], :shape[1], :shape[2]].astype(np.float32)
    d = np.abs(mine[:shape[0], :shape[1], :shape[2]] - ref)[mask]
    print("agreement", agree, "max logit diff on mask", d.max(), "median", np.median(d), "fg", b.sum(), ref_b.sum())
    assert agree >= 0.999
    assert (d <= 1.0 + 0.02 * np.abs(ref[mask])).all()
    assert np.array_equal(b[~mask], np.zeros_like(b[~mask]))


@pytest.mark.parametrize("shape,roi,overlap", [((60, 90, 70), (32, 32, 32), 0.25), ((40, 70, 66), (32, 32, 32), 0.75),
                                               ((70, 130, 64), (64, 64, 64), 0.25), ((50, 100, 48), (48, 32, 16), 0.75),
                                               ((130, 140, 64), (128, 128, 64), 0.5),
                                               # cfg5's cubic windows: 96^3 (two-tile columns), 128^3 / 160^3 (one-tile columns of the
                                               # fused conv), 192^3 (per-tap kernel: the fused conv's stages no longer fit)
                                               ((90, 100, 100), (96, 96, 96), 0.75), ((120, 130, 140), (128, 128, 128), 0.5),
                                               ((100, 170, 150), (160, 160, 160), 0.5), ((200, 100, 180), (192, 192, 192), 0.25)])
def test_window_and_overlap_sweep_against_oracle(shape, roi, overlap):
    """BASELINE.json configs[4] (patch-size / overlap sweep) at oracle-sized volumes: window grid of
    sliding_window_inferer.py:140-143 for overlaps 0.25 / 0.5 / 0.75 and other window shapes, averaged logits and
    binaries against the torch-fp32 restatement."""
    from gpu_common import ctx_with
    ctx, sd, onet = ctx_with("random")
    vol = P.synth_volume(shape, 41, roi=roi)
    vol[:shape[0], :shape[1], :shape[2]][vol[:shape[0], :shape[1], :shape[2]] == 0] = 500
    vol[:6] = 0
    shape_pad = vol.shape
    pred = lambda t: onet(t.cuda()).cpu()
    avg = P.infer_average(vol, roi, overlap, pred, 4)
    ref_b = P.create_binaries(avg, vol, shape, 0.5)
    b = np.empty(shape, dtype=np.uint8)
    mine = np.empty(shape_pad, dtype=np.float32)
    st = ctx.segment(vol, shape_pad, shape, roi, b, overlap=overlap, avg_logits_out=mine)
    grid = [len(s) for s in __import__("delivr_cfos_b200")._lib.window_grid(shape_pad, roi, overlap)]
    assert st["windows_total"] == grid[0] * grid[1] * grid[2]
    mask = P.ccl_ref.erode6((vol[:shape[0], :shape[1], :shape[2]] > 0).astype(np.uint8), 30) > 0
    ref = avg[:shape[0], :shape[1], :shape[2]].astype(np.float32)
    d = np.abs(mine[:shape[0], :shape[1], :shape[2]] - ref)[mask]
    print("windows", st["windows_total"], "agreement", (b == ref_b).mean(), "max diff", d.max() if d.size else None)
    assert mask.any()
    assert (d <= 1.0 + 0.02 * np.abs(ref[mask])).all()
    # random weights put many logits next to the threshold and a low overlap averages fewer windows, so the 99.9 %
    # bar of the shipped-weights tests is replaced by its cause: a voxel may only differ where the reference logit is
    # within the bf16 error of zero
    mism = b != ref_b
    assert mism.mean() <= 0.01
    assert (np.abs(ref[mism]) <= 0.5).all(), np.abs(ref[mism]).max()
