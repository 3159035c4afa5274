"""dlv_paint_boxes vs the restated reference loop and the reference's own golden."""
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import paint_ref
from paint_common import load_paint_golden, random_paint_case

pytestmark = pytest.mark.gpu


def _ctx():
    from delivr_cfos_b200 import Context
    _ctx.c = getattr(_ctx, "c", None) or Context(0)
    return _ctx.c


@pytest.mark.parametrize("shape,n,dtype,big,chunk,dev", [
    ((9, 14, 11), 30, np.uint8, False, 0, False), ((9, 14, 11), 30, np.uint16, False, 0, True),
    ((24, 50, 64), 200, np.uint8, True, 0, True), ((24, 50, 64), 200, np.uint16, True, 50 * 64 * 5, False),
    ((7, 33, 129), 64, np.uint8, True, 33 * 129 * 2, True), ((40, 64, 96), 500, np.uint16, True, 0, True),
    ((5, 6, 7), 0, np.uint8, False, 0, False), ((1, 1, 1), 3, np.uint16, False, 0, False),
])
def test_paint_boxes_equals_reference_loop(shape, n, dtype, big, chunk, dev):
    mask, boxes, values = random_paint_case(shape, n, seed=n + shape[2], big=big)
    want = paint_ref.paint_boxes_ref(mask, boxes, values, dtype)
    ctx = _ctx()
    if dev:
        tdt = torch.uint8 if dtype == np.uint8 else torch.uint16
        outs = [torch.full(shape, 7, dtype=tdt, device="cuda") for _ in range(3)]
        ctx.paint_boxes(torch.from_numpy(mask).cuda(), shape, boxes, values, outs, chunk_voxels=chunk)
        got = [o.cpu().numpy() for o in outs]
    else:
        got = [np.full(shape, 7, dtype=dtype) for _ in range(3)]
        ctx.paint_boxes(mask, shape, boxes, values, got, chunk_voxels=chunk)
    for c in range(3):
        assert np.array_equal(got[c], want[c]), c
    # one channel
    (one,) = [np.empty(shape, dtype=dtype)]
    ctx.paint_boxes(mask, shape, boxes, values[:, 1], [one], chunk_voxels=chunk)
    assert np.array_equal(one, want[1])
    # the scratch was left clean: a second call gives the same result
    again = [np.empty(shape, dtype=dtype) for _ in range(3)]
    ctx.paint_boxes(mask, shape, boxes, values, again, chunk_voxels=chunk)
    assert all(np.array_equal(a, w) for a, w in zip(again, want))


def test_paint_rejects_bad_arguments():
    from delivr_cfos_b200._lib import DlvError
    ctx = _ctx()
    mask = np.ones((2, 2, 2), dtype=np.uint8)
    out = np.empty((2, 2, 2), dtype=np.uint8)
    with pytest.raises(DlvError):
        ctx.paint_boxes(mask, mask.shape, [[-1, 2, 0, 2, 0, 2]], [5], [out])
    with pytest.raises(ValueError):
        ctx.paint_boxes(mask, mask.shape, [[0, 2, 0, 2, 0, 2]], [5], [np.empty((2, 2, 2), dtype=np.float32)])


def _stage(tmp_path, g, brain="brainP"):
    d = {k: str(tmp_path / k) + "/" for k in ("pred", "post", "csv", "cache", "out", "maskdet")}
    for v in d.values():
        os.makedirs(v)
    os.makedirs(os.path.join(d["pred"], brain, "binary_segmentations"))
    mm = np.lib.format.open_memmap(os.path.join(d["pred"], brain, "binary_segmentations", "binaries.npy"), mode="w+",
                                   dtype=np.uint8, shape=g["shape"])
    mm[...] = g["mask"]
    mm.flush()
    settings = {"postprocessing": {"output_location": d["post"]},
                "visualization": {"input_prediction_location": d["pred"], "input_csv_location": d["csv"],
                                  "output_location": d["out"], "cache_location": d["cache"], "region_id_rgb": True,
                                  "region_id_grayvalues": True, "no_atlas_depthmap": False},
                "mask_detection": {"output_location": d["maskdet"],
                                   "downsample_steps": {"original_um_x": 1.62, "original_um_y": 1.62, "original_um_z": 6.0,
                                                        "downsample_um_x": 5.0, "downsample_um_y": 5.0, "downsample_um_z": 12.0}},
                "FLAGS": {"LOAD_ALL_RAM": True}}
    return d, settings


def _read_planes(fmt, shape, dtype):
    import cv2
    out = np.zeros(shape, dtype=dtype)
    for z in range(shape[0]):
        p = cv2.imread(fmt.format(z=str(z).zfill(4)), cv2.IMREAD_UNCHANGED)
        assert p is not None and p.dtype == dtype
        out[z] = p
    return out


@pytest.mark.parametrize("cached_stats", [True, False])
def test_blob_highlighter_reproduces_reference_golden(tmp_path, cached_stats):
    """The drop-in blob_highlighter on the golden's inputs writes the reference's files with the reference's pixels.
    cached_stats=False: the statistics come from dlv_ccl instead of the pickle the reference's count_blobs wrote."""
    from delivr_cfos_b200.blob_highlighter import blob_highlighter
    g = load_paint_golden()
    brain = "brainP"
    d, settings = _stage(tmp_path, g, brain)
    open(os.path.join(d["csv"], f"cells_{brain}.csv"), "wb").write(g["csv_bytes"])
    if cached_stats:
        with open(os.path.join(d["post"], f"{brain}-stats.pickle"), "wb") as f:
            pickle.dump({"voxel_counts": g["voxel_counts"], "bounding_boxes": g["bounding_boxes"].copy(), "centroids": g["centroids"]}, f)
    blob_highlighter(settings, (brain, ""), g["stack_shape"])
    rgb_dir = os.path.join(d["out"], brain + "_rgb_tiffs")
    for c, key in enumerate(("red", "green", "blue")):
        got = _read_planes(os.path.join(rgb_dir, brain + f"rgb_C0{c}_z" + "{z}.tif"), g["shape"], np.uint8)
        assert np.array_equal(got, g[key]), key
    reg = _read_planes(os.path.join(d["out"], brain, brain + "_region_id_tiffs", "region_id_{z}.tif"), g["shape"], np.uint16)
    assert np.array_equal(reg, g["region"])
    files = sorted(os.path.relpath(os.path.join(r, f), d["out"]) for r, _, fs in os.walk(d["out"]) for f in fs)
    assert files == [str(f) for f in g["files"]]
    assert not os.path.exists(os.path.join(d["cache"], brain))          # cache removed (blob_highlighter.py:167-170)


def test_depth_map_blobs_equals_restated_loop(tmp_path):
    import cv2
    from scipy.ndimage import distance_transform_edt
    from delivr_cfos_b200.blob_depthmap import blob_depths, depth_map_blobs
    from oracle import ccl_ref
    g = load_paint_golden()
    brain = "brainP"
    d, settings = _stage(tmp_path, g, brain)
    settings["visualization"]["no_atlas_depthmap"] = True
    ds = settings["mask_detection"]["downsample_steps"]
    dshape = tuple(int(np.ceil(s * ds[f"original_um_{a}"] / ds[f"downsample_um_{a}"])) + 1 for s, a in zip(g["shape"], "zyx"))
    zz, yy, xx = np.ogrid[:dshape[0], :dshape[1], :dshape[2]]
    small = ((((zz - dshape[0] / 2) / (dshape[0] / 2)) ** 2 + ((yy - dshape[1] / 2) / (dshape[1] / 2)) ** 2 +
              ((xx - dshape[2] / 2) / (dshape[2] / 2)) ** 2) < 1.0).astype(np.uint16) * 900
    os.makedirs(os.path.join(d["maskdet"], brain))
    assert cv2.imwritemulti(os.path.join(d["maskdet"], brain, "downsampled_masked_stack.tif"), [p for p in small])
    depth_map_blobs(settings, brain, g["stack_shape"])
    got = _read_planes(os.path.join(d["out"], brain, brain + "_depthmap_tiffs", "depthmap_{z}.tif"), g["shape"], np.uint16)
    # restated loop (blob_depthmap.py:174-207 with a 3-D volume and N = number of components)
    labels, n = ccl_ref.connected_components26(g["mask"])
    stats = ccl_ref.statistics(labels, n)
    dist = distance_transform_edt(np.pad(small, 1), sampling=(ds["downsample_um_z"], ds["downsample_um_y"], ds["downsample_um_x"]))
    dist = dist[1:-1, 1:-1, 1:-1].astype(np.uint16)
    depths = blob_depths(stats, dist, settings)
    st = {"bounding_boxes": np.array(stats["bounding_boxes"]).astype(np.int64)}
    (want,) = paint_ref.highlight_ref(g["mask"], st, np.arange(n), depths[:n], g["stack_shape"], dtype=np.uint16)
    assert np.array_equal(got, want)
    assert int((got > 0).sum()) > 0


@pytest.mark.parametrize("shape,sampling,exact", [((12, 17, 23), (12.0, 5.0, 5.0), True), ((30, 40, 35), (25.0, 25.0, 25.0), True),
                                                  ((9, 20, 21), (6.0, 1.62, 1.62), False), ((1, 5, 7), (1.0, 1.0, 1.0), True),
                                                  ((6, 6, 6), (2.0, 3.0, 4.0), True)])
def test_edt_equals_ndimage(shape, sampling, exact):
    """dlv_edt vs the reference's call (blob_depthmap.py:171-178): pad with zeros, distance_transform_edt, crop."""
    from scipy.ndimage import distance_transform_edt
    rng = np.random.default_rng(shape[2])
    stack = (rng.random(shape) < 0.93).astype(np.uint16) * 700
    if shape[0] > 6:
        stack[:2] = 0
        stack[shape[0] // 2, 3:9, 2:11] = 0
    want = distance_transform_edt(np.pad(stack, 1), sampling=sampling)[1:-1, 1:-1, 1:-1]
    got = _ctx().edt(np.ascontiguousarray(stack != 0).view(np.uint8), sampling)
    if exact:                       # sampling values whose products are exact in fp64: bit-identical distances
        assert np.array_equal(got, want)
    else:
        assert np.allclose(got, want, rtol=1e-14, atol=0)
    assert np.array_equal(got.astype(np.uint16), want.astype(np.uint16))
    dev = torch.empty(shape, dtype=torch.float64, device="cuda")
    _ctx().edt(torch.from_numpy(np.ascontiguousarray(stack != 0).view(np.uint8)).cuda(), sampling, out=dev)
    assert np.array_equal(dev.cpu().numpy(), got)
