"""The reference's public names on the path that round 1 left untested: SlidingWindowInferer.__call__ /
sliding_window_inference (in-place sums and counts, flip_dim), create_nifti_seg, the optional Gaussian blend against
the restated MONAI importance map, count_blobs on the reference's own golden masks (CSV bytes unconditional), the
out-of-core / streamed paths of the drop-ins, and the unmodified reference __main__.py through the import swap."""
import os
import pickle

import numpy as np
import pytest
import torch

import main_flow
from conftest import weights_path
from helpers import load_golden
from oracle import pipeline_ref as P, unet_ref

pytestmark = pytest.mark.gpu


def _nets(seed=3):
    from delivr_cfos_b200.inference.sliding_window_inferer import DelivrNet
    sd = unet_ref.random_state_dict(seed)
    onet = unet_ref.BasicUNet(dropout=0.1)
    onet.load_state_dict(unet_ref.strip_module_prefix(sd), strict=True)
    return DelivrNet(state_dict=sd), onet.eval().cuda()


def _volume(shape, roi, seed):
    vol = P.synth_volume(shape, seed, roi=roi)
    sub = vol[:shape[0], :shape[1], :shape[2]]
    sub[sub == 0] = 500
    vol[:6] = 0                                   # first window layer partly empty; with roi_z <= 6 whole windows are skipped
    return vol


@pytest.mark.parametrize("flip_dim", [None, 2, 3])
def test_sliding_window_inferer_call_accumulates_like_the_reference(flip_dim):
    """sliding_window_inferer.py:86-253: `output_image +=` (fp16 sums), `count_map += 1`, windows flipped before the
    net and back after it.  The oracle runs the reference's loop with batch size 1 (its skip rule is per batch)."""
    from delivr_cfos_b200.inference.sliding_window_inferer import SlidingWindowInferer
    roi, shape = (32, 48, 32), (64, 96, 64)
    net, onet = _nets()
    vol = _volume(shape, roi, 51)
    vol[:40, :50] = 0                             # whole windows skipped (-1000)
    ref_sum = np.zeros(vol.shape, dtype=np.float16)
    ref_cnt = np.zeros(vol.shape, dtype=np.uint8)
    P.sliding_window_pass(vol, roi, 0.5, lambda t: onet(t.cuda()).cpu(), 1, ref_sum, ref_cnt, flip_dim=flip_dim)
    out = torch.zeros((1, 1) + vol.shape, dtype=torch.float16)
    out[0, 0, 5, 5, 5] = 3.0                      # in place: what is already there stays
    cnt = torch.zeros((1, 1) + vol.shape, dtype=torch.uint8)
    inferer = SlidingWindowInferer(roi_size=roi, sw_batch_size=4, overlap=0.5, mode="gaussian")   # mode ignored like the reference (:148)
    inferer(torch.from_numpy(vol.astype(np.int32))[None, None], net, output_image=out, count_map=cnt, flip_dim=flip_dim)
    assert np.array_equal(cnt[0, 0].numpy(), ref_cnt)
    mine = out[0, 0].numpy().astype(np.float32)
    mine[5, 5, 5] -= 3.0
    ref = ref_sum.astype(np.float32)
    c = ref_cnt.astype(np.float32)
    tol = c * 1.0 + 0.02 * np.abs(ref) + np.abs(ref) * 2.0 ** -9 + 0.51      # bf16 bar per window + two fp16 roundings
    assert (np.abs(mine - ref) <= tol).all(), float((np.abs(mine - ref) - tol).max())
    allskipped = (ref_cnt > 0) & (ref == -1000.0 * c)          # every covering window skipped: exactly -1000 each
    assert allskipped.any() and np.array_equal(mine[allskipped], ref[allskipped])


def test_create_nifti_seg_mirror(tmp_path):
    """inference.py:31-95 through the mirror's own signature: binaries bit-exact, sigmoid to fp32 rounding."""
    from delivr_cfos_b200.inference.inference import create_nifti_seg
    shape, pad = (70, 75, 66), (96, 96, 96)
    rng = np.random.default_rng(8)
    vol = np.zeros(pad, dtype=np.uint16)
    vol[:shape[0], :shape[1], :shape[2]] = rng.integers(0, 3, size=shape) + (rng.random(shape) < 0.9)
    vol[2:68, 3:72, 2:64] |= 1
    logits = (rng.standard_normal(pad) * 3).astype(np.float16)
    logits[rng.random(pad) < 0.01] = 0.0
    ref_b, ref_sig = P.create_binaries(logits, vol, shape, 0.5, return_sigmoid=True)
    out = os.path.join(str(tmp_path), "binaries.npy")
    act = os.path.join(str(tmp_path), "network_output.npy")
    create_nifti_seg(0.5, torch.from_numpy(logits)[None, None], out, act, vol[None, None], (1, 1) + shape)
    b = np.load(out)
    assert b.dtype == np.uint8 and np.array_equal(b, ref_b) and ref_b.sum() > 0
    assert np.abs(np.load(act) - ref_sig).max() < 1e-6
    assert np.lib.format.open_memmap(out, mode="r").offset == 128


def test_gaussian_blend_against_restated_importance_map():
    """blend_mode = 1 (dlv_segment.cu gaussian_1d): weighted average of the window logits with the separable
    importance map of SURVEY.md section 8(c).  No reference oracle exists for this mode (parity unpinned)."""
    from gpu_common import ctx_with
    ctx, sd, onet = ctx_with("random")
    roi, shape = (32, 48, 32), (60, 100, 70)
    vol = _volume(shape, roi, 61)
    w = P.gaussian_importance_map(roi)
    ref = P.infer_average_weighted(vol, roi, 0.5, lambda t: onet(t.cuda()).cpu(), w)
    b = np.empty(shape, dtype=np.uint8)
    mine = np.empty(vol.shape, dtype=np.float32)
    ctx.segment(vol, vol.shape, shape, roi, b, overlap=0.5, blend_mode=1, avg_logits_out=mine)
    mask = P.ccl_ref.erode6((vol[:shape[0], :shape[1], :shape[2]] > 0).astype(np.uint8), 30) > 0
    r = ref[:shape[0], :shape[1], :shape[2]]
    d = np.abs(mine[:shape[0], :shape[1], :shape[2]] - r)
    assert mask.any() and (d[mask] <= 1.0 + 0.02 * np.abs(r[mask])).all(), float(d[mask].max())
    # and it is a different blend from the constant one
    const = np.empty(vol.shape, dtype=np.float32)
    ctx.segment(vol, vol.shape, shape, roi, b, overlap=0.5, blend_mode=0, avg_logits_out=const)
    assert np.abs(const - mine)[:shape[0], :shape[1], :shape[2]][mask].max() > 1e-3


@pytest.mark.parametrize("name", ["g1_notta", "g2_memmap", "g3_tta"])
def test_count_blobs_on_reference_masks_writes_reference_csv(name, tmp_path):
    """The mask the UNMODIFIED reference produced (golden binaries) through the count_blobs drop-in: CSV bytes equal
    the CSV the reference's own count_blobs wrote for it - unconditionally (no dependence on the network's numerics)."""
    from delivr_cfos_b200.count_blobs import count_blobs
    g = load_golden(name)
    m = g["meta"]
    tmp = str(tmp_path)
    d = os.path.join(tmp, "in", "brainA", "binary_segmentations")
    os.makedirs(d)
    mm = np.lib.format.open_memmap(os.path.join(d, "binaries.npy"), mode="w+", dtype=np.uint8, shape=tuple(m["shape"]))
    mm[...] = g["binaries"]
    mm.flush()
    post = os.path.join(tmp, "post") + "/"
    count_blobs({"postprocessing": {"output_location": post}, "FLAGS": {}}, os.path.join(tmp, "in"), 0, "brainA", (1, 1) + tuple(m["shape"]))
    assert open(os.path.join(post, m["csv_file"])).read() == g["csv"]
    lab, n, st = P.blob_table(g["binaries"])
    assert np.array_equal(np.load(os.path.join(post, f"brainA-{n}-cc3d.npy")), lab)


def _run_dropin(tmp, tag, vol, shape, roi, net, monkeypatch, chunks=None, tta=False):
    from delivr_cfos_b200.inference import inference as inf
    if chunks:
        monkeypatch.setattr(inf, "chunks_needed", lambda *a, **k: chunks)
    else:
        monkeypatch.undo()
    d = os.path.join(tmp, tag)
    os.makedirs(d)
    src = os.path.join(d, "masked_nifti.npy")
    mm = np.lib.format.open_memmap(src, mode="w+", dtype=np.uint16, shape=(1, 1) + vol.shape)
    mm[0, 0] = vol
    mm.flush()
    settings = {"blob_detection": {"window_dimensions": {"window_dim_0": roi[0], "window_dim_1": roi[1], "window_dim_2": roi[2]}},
                "FLAGS": {"SAVE_ACTIVATED_OUTPUT": True}}
    s = inf.run_inference([src], os.path.join(d, "out"), (1, 1) + tuple(shape), comment="b", model_weights="unused", tta=tta,
                          load_all_ram=False, settings=settings, _net=net)
    return {n: np.load(os.path.join(s, n)) for n in ("binary_segmentations/binaries.npy", "binary_segmentations/network_output.npy",
                                                     "inference_output.npy")}


@pytest.mark.parametrize("tta", [False, True])
def test_out_of_core_chunks_equal_in_core(tmp_path, monkeypatch, tta):
    """run_inference's z-chunk mode (a volume beyond device memory) writes the files of the in-core mode, bit for bit."""
    roi, shape = (32, 32, 32), (100, 70, 66)
    net, _ = _nets(5)
    vol = _volume(shape, roi, 71)
    ref = _run_dropin(str(tmp_path), "incore", vol, shape, roi, net, monkeypatch, tta=tta)
    assert ref["binary_segmentations/binaries.npy"].sum() > 0
    for k in (2, 4):
        got = _run_dropin(str(tmp_path), f"chunks{k}", vol, shape, roi, net, monkeypatch, chunks=k, tta=tta)
        for name in ref:
            assert np.array_equal(ref[name], got[name], equal_nan=True), (k, name)


def test_ccl_streamed_from_host_equals_single_call():
    """count_blobs on a mask larger than the device budget: sub-slabs uploaded one after the other, labels streamed to
    the sink - same labels and table as one dlv_ccl call."""
    from gpu_common import ctx_with
    from delivr_cfos_b200.slabs import ccl_any_size
    ctx = ctx_with("random")[0]
    m = P.synth_mask((90, 64, 80), 13)
    m[10:80, 30, 30] = 1
    ref_lab = np.empty(m.shape, dtype=np.uint32)
    ref = ctx.ccl(m, m.shape, labels_out=ref_lab)
    got_lab = np.zeros(m.shape, dtype=np.uint32)

    def sink(z0, z1, lab):
        got_lab[z0:z1] = lab.cpu().numpy().view(np.uint32)

    plane = m.shape[1] * m.shape[2]
    for budget in (6 * plane * 17 / 0.7, 6 * plane * 40 / 0.7):        # 17- and 40-plane sub-slabs; the small one re-labels for the sink
        got_lab[...] = 0
        t = ccl_any_size(ctx, m, m.shape, labels_sink=sink, bytes_free=int(budget))
        assert t["n"] == ref["n"] and np.array_equal(got_lab, ref_lab)
        for k in ("voxel_counts", "sums", "bounding_boxes"):
            assert np.array_equal(t[k], ref[k]), k
        t2 = ccl_any_size(ctx, m, m.shape, bytes_free=int(budget))
        assert t2["n"] == ref["n"] and np.array_equal(t2["sums"], ref["sums"])


@pytest.mark.skipif(main_flow.reference_main_path() is None, reason="reference __main__.py not staged (run __graft_entry__.build() where /root/reference exists)")
def test_unmodified_reference_main_on_the_gpu(tmp_path):
    """__main__.py:106-166 of the reference, unmodified, with the import swap of INTEGRATION.md section 1: the files of
    the blob_detection and postprocessing stages appear with the reference's names and the table is exact for the
    binaries written."""
    wp = weights_path()
    if wp is None:
        pytest.skip("shipped checkpoint not staged")
    roi, shape = (32, 48, 32), (70, 110, 80)
    vol = _volume(shape, roi, 81)[:shape[0], :shape[1], :shape[2]]
    cfg, settings = main_flow.make_tree(str(tmp_path), np.ascontiguousarray(vol), roi, tta=True, save_act=True, load_all_ram=False)
    main_flow.run_reference_main(cfg)
    session = os.path.join(settings["blob_detection"]["output_location"], "brainA")
    assert sorted(os.listdir(session)) == ["binary_segmentations", "inference_output.npy", "network_outputs"]
    assert sorted(os.listdir(os.path.join(session, "binary_segmentations"))) == ["binaries.npy", "network_output.npy"]
    b = np.load(os.path.join(session, "binary_segmentations", "binaries.npy"))
    assert b.shape == tuple(shape)
    lab, n, st = P.blob_table(b)
    post = settings["postprocessing"]["output_location"]
    assert sorted(os.listdir(post)) == sorted([f"brainA-{n}-cc3d.npy", "brainA-stats.pickle", f"{tuple(shape)}_brainA.csv"])
    assert open(os.path.join(post, f"{tuple(shape)}_brainA.csv")).read() == P.csv_text(st, n)
    assert np.array_equal(np.load(os.path.join(post, f"brainA-{n}-cc3d.npy")), lab)
    with open(os.path.join(post, "brainA-stats.pickle"), "rb") as f:
        assert np.array_equal(pickle.load(f)["voxel_counts"], st["voxel_counts"])
