"""CPU suite (-m "not gpu"): the oracle against the golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py), against scipy.ndimage, and its own invariants."""
import numpy as np
import pytest
from scipy import ndimage

from conftest import weights_path
from helpers import load_golden
from oracle import ccl_ref, pipeline_ref as P, unet_ref


def test_strict_checkpoint_load(real_weights):
    net = unet_ref.load_reference_net(real_weights)
    assert sum(p.numel() for p in net.parameters()) == 5749377


def test_random_state_dict_matches_checkpoint_layout(real_weights):
    import torch
    ck = torch.load(real_weights, map_location="cpu", weights_only=True)["state_dict"]
    rnd = unet_ref.random_state_dict(0)
    assert list(ck.keys()) == list(rnd.keys())
    assert all(ck[k].shape == rnd[k].shape for k in ck)


@pytest.mark.parametrize("name", ["g2_memmap"])
def test_pipeline_restatement_reproduces_reference_golden(name, real_weights):
    """Unmodified reference (through shims) == oracle restatement, bit for bit, on the same input."""
    g = load_golden(name)
    m = g["meta"]
    net = unet_ref.load_reference_net(real_weights)
    avg = P.infer_average(g["volume"], tuple(m["roi"]), 0.5, net, m["sw"], tta=m["tta"])
    b, sig = P.create_binaries(avg, g["volume"], tuple(m["shape"]), 0.5, return_sigmoid=True)
    s = m["logit_stride"]
    assert np.abs(sig[::s[0], ::s[1], ::s[2]] - g["sigmoid_sub"]).max() < 1e-4
    assert (b == g["binaries"]).mean() >= 0.9999
    a, r = avg[::s[0], ::s[1], ::s[2]], g["avg_logits_sub"]
    ok = np.isfinite(a) & np.isfinite(r)
    assert np.abs(a[ok].astype(np.float32) - r[ok].astype(np.float32)).max() < 2e-2


@pytest.mark.parametrize("name", ["g1_notta", "g2_memmap", "g3_tta"])
def test_table_and_csv_reproduce_reference_golden(name):
    """count_blobs.py semantics: N, stats table, CSV text (N-1 rows) and file name - exact, from the golden binaries."""
    g = load_golden(name)
    lab, n, st = P.blob_table(g["binaries"])
    assert n == int(g["n_components"])
    assert np.array_equal(st["voxel_counts"], g["voxel_counts"])
    assert np.array_equal(st["bounding_boxes"], g["bounding_boxes"])
    assert np.array_equal(st["centroids"], g["centroids"], equal_nan=True)
    assert P.csv_text(st, n) == g["csv"]
    assert g["csv"].count("\n") - 1 == n - 1                                # finding 9: last component dropped
    assert P.csv_name(g["meta"]["shape"], "brainA") == g["meta"]["csv_file"]
    import zlib
    assert np.uint32(zlib.crc32(lab.tobytes())) == g["labels_crc"]


def test_host_csv_writer_matches_golden():
    """The product's CSV emitter (no pandas) writes the reference's bytes."""
    from delivr_cfos_b200.count_blobs import csv_text, statistics_from_labels
    g = load_golden("g1_notta")
    lab, n, st = P.blob_table(g["binaries"])
    assert csv_text(st, n) == g["csv"]
    st2 = statistics_from_labels(lab)
    assert np.array_equal(st2["voxel_counts"], g["voxel_counts"])
    assert np.array_equal(st2["bounding_boxes"], g["bounding_boxes"])
    assert np.array_equal(st2["centroids"], g["centroids"], equal_nan=True)


@pytest.mark.parametrize("shape,p", [((20, 33, 17), 0.08), ((16, 16, 16), 0.5), ((5, 40, 40), 0.2), ((1, 1, 1), 1.0), ((3, 1, 7), 0.6)])
def test_ccl_oracle_equals_scipy_label(shape, p):
    rng = np.random.default_rng(0)
    m = (rng.random(shape) < p).astype(np.uint8)
    lab, n = ccl_ref.connected_components26(m)
    sl, sn = ndimage.label(m, structure=np.ones((3, 3, 3)))
    assert n == sn and np.array_equal(lab, sl)


@pytest.mark.parametrize("it", [1, 3, 30])
def test_erosion_oracle_equals_scipy(it):
    rng = np.random.default_rng(1)
    m = (rng.random((40, 50, 60)) < 0.995).astype(np.uint8)
    assert np.array_equal(ccl_ref.erode6(m, it), ndimage.binary_erosion(m, iterations=it, border_value=1).astype(np.uint8))


def test_erosion_is_l1_distance_identity():
    """SURVEY finding 8: erode_k(mask, border 1) == (L1 distance to the nearest zero voxel in the block > k)."""
    rng = np.random.default_rng(2)
    m = (rng.random((24, 30, 36)) < 0.999).astype(np.uint8)
    zs = np.argwhere(m == 0)
    zz, yy, xx = np.indices(m.shape)
    d = np.full(m.shape, 10 ** 6)
    for z, y, x in zs:
        d = np.minimum(d, np.abs(zz - z) + np.abs(yy - y) + np.abs(xx - x))
    for k in (1, 4, 9):
        assert np.array_equal(ccl_ref.erode6(m, k), (d > k).astype(np.uint8))


def test_window_grid_counts():
    """2k-1 windows per dim at overlap 0.5, starts clamp to dim - roi, z-major order."""
    assert [len(s) for s in P.window_starts((96, 576, 512), (96, 96, 64), P.scan_interval((96, 576, 512), (96, 96, 64), 0.5))] == [1, 11, 15]
    assert len(P.window_list((288, 2112, 2048), (96, 96, 64), 0.5)) == 13545
    st = P.window_starts((100,), (64,), (32,))
    assert st == [[0, 32, 36]]
    from delivr_cfos_b200.inference.sliding_window_inferer import dense_window_starts
    for img, roi, ov in [((96, 576, 512), (96, 96, 64), 0.5), ((100, 130, 70), (64, 64, 32), 0.25), ((64, 64, 64), (64, 32, 16), 0.75)]:
        assert dense_window_starts(img, roi, ov) == P.window_starts(img, roi, P.scan_interval(img, roi, ov))


def test_arrayterator_blocks_match_numpy():
    for shape, buf in [((70, 33, 21), 5000), ((10, 8, 8), 10 ** 9), ((1500, 40, 40), 1600 * 62)]:
        arr = np.zeros(shape, dtype=np.uint8)
        blocks = []
        z = 0
        for sub in np.lib.Arrayterator(arr, buf):
            assert sub.shape[1:] == shape[1:]
            blocks.append((z, z + sub.shape[0]))
            z += sub.shape[0]
        assert P.arrayterator_blocks(shape, buf) == blocks
    from delivr_cfos_b200.inference.inference import erosion_block_planes
    assert erosion_block_planes((1500, 4000, 4000)) == 62                    # SURVEY finding 8
    assert erosion_block_planes((64, 512, 512)) == 0


@pytest.mark.parametrize("shape,p", [((20, 33, 17), 0.08), ((16, 16, 16), 0.5), ((1, 1, 1), 1.0), ((40, 64, 64), 0.1)])
def test_statistics_from_cached_labels(shape, p):
    """count_blobs.py:71-76: a cached label volume without a cached statistics pickle -> statistics from the labels.
    The host reduction (product code, no GPU involved) equals the oracle, also when a label value is missing."""
    from delivr_cfos_b200.count_blobs import statistics_from_labels
    m = P.synth_mask(shape, 3, kind="bernoulli", p=p)
    lab, n = ccl_ref.connected_components26(m)
    r = ccl_ref.statistics(lab, n)
    s = statistics_from_labels(lab)
    assert np.array_equal(s["voxel_counts"], r["voxel_counts"])
    assert np.array_equal(s["bounding_boxes"], r["bounding_boxes"])
    assert np.array_equal(s["centroids"], r["centroids"], equal_nan=True)
    if n >= 3:
        lab2 = lab.copy()
        lab2[lab2 == 2] = 0
        s2 = statistics_from_labels(lab2)
        assert s2["voxel_counts"][2] == 0 and tuple(s2["bounding_boxes"][2]) == (shape[0], -1, shape[1], -1, shape[2], -1)
        assert np.array_equal(s2["voxel_counts"][3:], r["voxel_counts"][3:])
