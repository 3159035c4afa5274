"""Host logic of the drop-ins without a GPU: run_inference's three execution modes (in-core, out-of-core z-chunks,
rank processes over gloo) write identical files; count_blobs over ranks equals the single-process result; the
reference's unmodified __main__.py drives both through the import swap of INTEGRATION.md.  The device side is the
oracle-backed engine of tests/cpu_engine.py (the product has no CPU engine)."""
import os
import pickle
import socket
import subprocess
import sys

import numpy as np
import pytest

import cpu_engine
import main_flow
from delivr_cfos_b200 import count_blobs as cb
from delivr_cfos_b200.inference import inference as inf
from oracle import ccl_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROI = (32, 32, 32)


def _volume(shape=(70, 80, 72), seed=3):
    rng = np.random.default_rng(seed)
    v = rng.integers(1, 60000, size=shape).astype(np.uint16)
    v[:3] = 0
    v[:, :2] = 0
    return v


def _write_input(tmp, vol):
    pad = tuple(int(np.ceil(d / r) * r) for d, r in zip(vol.shape, ROI))
    p = os.path.join(tmp, "masked_nifti.npy")
    m = np.lib.format.open_memmap(p, mode="w+", dtype=np.uint16, shape=(1, 1) + pad)
    m[0, 0, :vol.shape[0], :vol.shape[1], :vol.shape[2]] = vol
    m.flush()
    return p


def _settings(save_act=True):
    return {"blob_detection": {"window_dimensions": {"window_dim_0": ROI[0], "window_dim_1": ROI[1], "window_dim_2": ROI[2]}},
            "FLAGS": {"SAVE_ACTIVATED_OUTPUT": save_act}}


def _read_outputs(session):
    out = {}
    for name in ("binary_segmentations/binaries.npy", "binary_segmentations/network_output.npy", "inference_output.npy"):
        p = os.path.join(session, name)
        out[name] = np.load(p) if os.path.exists(p) else None
    return out


def _run(tmp, tag, vol, monkeypatch, env=None, chunks=None, tta=False):
    monkeypatch.setattr(inf, "_ENGINE_FACTORY", cpu_engine.OracleEngine)
    for k in ("DLV_GPUS", "DLV_ENGINE"):
        monkeypatch.delenv(k, raising=False)
    for k, v in (env or {}).items():
        monkeypatch.setenv(k, v)
    if chunks:
        monkeypatch.setattr(inf, "chunks_needed", lambda *a, **k: chunks)
    src = _write_input(tmp, vol)
    out = os.path.join(tmp, tag)
    session = inf.run_inference([src], out, (1, 1) + vol.shape, comment="brainA", model_weights="unused", tta=tta,
                                load_all_ram=False, settings=_settings())
    assert session == os.path.abspath(out + "/brainA")
    assert os.path.isdir(os.path.join(session, "network_outputs"))
    return _read_outputs(session)


def test_run_inference_modes_write_identical_files(tmp_path, monkeypatch):
    vol = _volume()
    tmp = str(tmp_path)
    ref = _run(tmp, "incore", vol, monkeypatch)
    b = ref["binary_segmentations/binaries.npy"]
    assert b.shape == vol.shape and b.dtype == np.uint8 and 0 < b.sum() < b.size
    assert ref["inference_output.npy"].dtype == np.float16 and ref["inference_output.npy"].shape == (1, 1, 96, 96, 96)
    for tag, kw in (("chunks3", dict(chunks=3)), ("chunks5", dict(chunks=5)),
                    ("ranks2", dict(env={"DLV_GPUS": "2", "DLV_ENGINE": "cpu_engine:OracleEngine"})),
                    ("ranks3_unbalanced", dict(env={"DLV_GPUS": "3", "DLV_ENGINE": "cpu_engine:OracleEngine", "DLV_BALANCE": "0"}))):
        if "env" in kw:
            kw["env"]["PYTHONPATH"] = os.path.join(ROOT, "tests") + os.pathsep + os.environ.get("PYTHONPATH", "")
        got = _run(tmp, tag, vol, monkeypatch, **kw)
        for name, a in ref.items():
            assert a is not None and got[name] is not None, (tag, name)
            assert a.dtype == got[name].dtype and np.array_equal(a, got[name], equal_nan=True), (tag, name)


def test_run_inference_tta_chunks_equal_incore(tmp_path, monkeypatch):
    vol = _volume((66, 70, 68), seed=9)
    ref = _run(str(tmp_path), "a", vol, monkeypatch, tta=True)
    got = _run(str(tmp_path), "b", vol, monkeypatch, tta=True, chunks=2)
    assert np.array_equal(ref["binary_segmentations/binaries.npy"], got["binary_segmentations/binaries.npy"])


def test_chunks_needed_sizes():
    cfg4_pad, cfg4 = (1536, 4032, 4032), (1500, 4000, 4000)
    roi = (96, 96, 64)
    assert inf.chunks_needed((288, 2112, 2048), (256, 2048, 2048), roi, 178 << 30, False, False) == 1      # cfg2 fits
    k = inf.chunks_needed(cfg4_pad, cfg4, roi, 178 << 30, False, False)
    assert k == 2                                                                                             # a whole brain: two chunks
    assert inf.chunks_needed(cfg4_pad, cfg4, roi, 100 << 30, False, False) > k
    assert inf.chunks_needed(cfg4_pad, cfg4, roi, 178 << 30, True, True) >= k
    with pytest.raises(MemoryError):
        inf.chunks_needed(cfg4_pad, cfg4, roi, 50 << 30, False, False)


# ------------------------------------------------------------------------------------------- count_blobs
def _oracle_label_and_count(bin_img, labels_path, device):
    lab, n = ccl_ref.connected_components26(np.ascontiguousarray(bin_img))
    st = ccl_ref.statistics(lab, n)
    if labels_path:
        np.save(labels_path, lab.astype(np.uint32))
    return {"n": n, **st}


def _blob_inputs(tmp, shape=(40, 30, 34), seed=4):
    rng = np.random.default_rng(seed)
    b = (rng.random(shape) < 0.1).astype(np.uint8)
    b[5:35, 10, 10] = 1
    d = os.path.join(tmp, "in", "brainA", "binary_segmentations")
    os.makedirs(d, exist_ok=True)
    m = np.lib.format.open_memmap(os.path.join(d, "binaries.npy"), mode="w+", dtype=np.uint8, shape=shape)
    m[...] = b
    m.flush()
    return b


def _count_files(out):
    files = sorted(os.listdir(out))
    res = {}
    for f in files:
        p = os.path.join(out, f)
        if f.endswith(".npy"):
            res[f] = np.load(p)
        elif f.endswith(".pickle"):
            res[f] = pickle.load(open(p, "rb"))
        else:
            res[f] = open(p).read()
    return res


RANK_SCRIPT = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import cpu_engine
from delivr_cfos_b200 import count_blobs as cb
cb._LABEL_SLAB_FACTORY = lambda device, b: cpu_engine.OracleLabelSlab(b)
dist.init_process_group("gloo")
cb.count_blobs({settings!r}, {path_in!r}, 0, "brainA", {stack!r})
dist.destroy_process_group()
"""


@pytest.mark.parametrize("world", [2, 3])
def test_count_blobs_ranks_equal_single(tmp_path, monkeypatch, world):
    tmp = str(tmp_path)
    b = _blob_inputs(tmp)
    stack = (1, 1) + b.shape
    monkeypatch.setattr(cb, "_label_and_count", _oracle_label_and_count)
    s1 = {"postprocessing": {"output_location": os.path.join(tmp, "out1") + "/"}, "FLAGS": {}}
    cb.count_blobs(s1, os.path.join(tmp, "in"), 0, "brainA", stack)
    ref = _count_files(s1["postprocessing"]["output_location"])
    assert len(ref) == 3
    sN = {"postprocessing": {"output_location": os.path.join(tmp, "outN") + "/"}, "FLAGS": {}}
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    script = RANK_SCRIPT.format(root=ROOT, settings=sN, path_in=os.path.join(tmp, "in"), stack=stack)
    procs = [subprocess.Popen([sys.executable, "-c", script],
                              env=dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port)))
             for r in range(world)]
    assert [p.wait(timeout=300) for p in procs] == [0] * world
    got = _count_files(sN["postprocessing"]["output_location"])
    assert sorted(got) == sorted(ref)
    for k, v in ref.items():
        if isinstance(v, dict):
            for kk in v:
                assert np.array_equal(np.asarray(v[kk]), np.asarray(got[k][kk]), equal_nan=True), (k, kk)
        elif isinstance(v, np.ndarray):
            assert np.array_equal(v, got[k]), k
        else:
            assert v == got[k], k


def test_count_blobs_cached_branches(tmp_path, monkeypatch):
    """count_blobs.py:67-76 (cached label volume) and :90-94 (cached statistics): same CSV as the uncached run;
    FLAGS.SAVE_CC3D_LABELS = False skips the 4 B/voxel dump."""
    tmp = str(tmp_path)
    b = _blob_inputs(tmp, seed=6)
    stack = (1, 1) + b.shape
    monkeypatch.setattr(cb, "_label_and_count", _oracle_label_and_count)
    out = os.path.join(tmp, "out") + "/"
    s = {"postprocessing": {"output_location": out}, "FLAGS": {}}
    cb.count_blobs(s, os.path.join(tmp, "in"), 0, "brainA", stack)
    first = _count_files(out)
    csv_name = next(f for f in first if f.endswith(".csv"))
    # labels cached, statistics not: statistics_from_labels on the memmapped file
    os.remove(os.path.join(out, "brainA-stats.pickle"))
    os.remove(os.path.join(out, csv_name))
    monkeypatch.setattr(cb, "_label_and_count", lambda *a: (_ for _ in ()).throw(AssertionError("labelled again despite the cache")))
    cb.count_blobs(s, os.path.join(tmp, "in"), 0, "brainA", stack)
    second = _count_files(out)
    assert second[csv_name] == first[csv_name]
    for k in ("voxel_counts", "bounding_boxes", "centroids"):
        assert np.array_equal(np.asarray(second["brainA-stats.pickle"][k]), np.asarray(first["brainA-stats.pickle"][k]), equal_nan=True)
    # both cached
    os.remove(os.path.join(out, csv_name))
    cb.count_blobs(s, os.path.join(tmp, "in"), 0, "brainA", stack)
    assert _count_files(out)[csv_name] == first[csv_name]
    # no label dump
    monkeypatch.setattr(cb, "_label_and_count", _oracle_label_and_count)
    out2 = os.path.join(tmp, "out2") + "/"
    cb.count_blobs({"postprocessing": {"output_location": out2}, "FLAGS": {"SAVE_CC3D_LABELS": False}}, os.path.join(tmp, "in"), 0, "brainA", stack)
    assert sorted(os.listdir(out2)) == sorted(f for f in first if not f.endswith(".npy"))
    assert open(os.path.join(out2, csv_name)).read() == first[csv_name]


# ------------------------------------------------------------------------------------------- __main__.py
@pytest.mark.skipif(main_flow.reference_main_path() is None, reason="reference __main__.py not available")
def test_unmodified_reference_main_drives_the_dropins(tmp_path, monkeypatch):
    vol = _volume((66, 72, 70), seed=12)
    cfg, settings = main_flow.make_tree(str(tmp_path), vol, ROI, tta=True)
    monkeypatch.setattr(inf, "_ENGINE_FACTORY", cpu_engine.OracleEngine)
    monkeypatch.setattr(cb, "_label_and_count", _oracle_label_and_count)
    main_flow.run_reference_main(cfg)
    session = os.path.join(settings["blob_detection"]["output_location"], "brainA")
    b = np.load(os.path.join(session, "binary_segmentations", "binaries.npy"))
    assert b.shape == vol.shape and b.sum() > 0
    post = settings["postprocessing"]["output_location"]
    files = sorted(os.listdir(post))
    lab, n = ccl_ref.connected_components26(b)
    assert files == sorted([f"brainA-{n}-cc3d.npy", "brainA-stats.pickle", f"{vol.shape}_brainA.csv"])
    rows = open(os.path.join(post, f"{vol.shape}_brainA.csv")).read().splitlines()
    assert rows[0] == ",Blob,Coords,Size" and len(rows) == n          # header + labels 1..N-1 (count_blobs.py:104)
