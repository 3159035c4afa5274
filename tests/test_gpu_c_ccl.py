"""dlv_ccl vs the plain-C oracle (oracle/ccl_ref.c): labels, N, counts, sums, bboxes bit-exact."""
import numpy as np
import pytest
import torch

from gpu_common import ctx_with
from oracle import ccl_ref, pipeline_ref as P

pytestmark = pytest.mark.gpu


def _compare(mask):
    from delivr_cfos_b200 import Context
    ctx = _compare.ctx = getattr(_compare, "ctx", None) or Context(0)
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    lab = np.empty(mask.shape, dtype=np.uint32)
    t = ctx.ccl(mask, mask.shape, labels_out=lab)
    rl, rn = ccl_ref.connected_components26(mask)
    rs = ccl_ref.statistics(rl, rn)
    assert t["n"] == rn
    assert np.array_equal(lab, rl)
    assert np.array_equal(t["voxel_counts"], rs["voxel_counts"])
    assert np.array_equal(t["sums"], rs["sums"])
    assert np.array_equal(t["bounding_boxes"], rs["bounding_boxes"])
    assert np.array_equal(t["centroids"], rs["centroids"], equal_nan=True)
    # device-pointer path gives the same labels
    md = torch.from_numpy(mask).cuda()
    ld = torch.empty(mask.shape, dtype=torch.int32, device="cuda")
    t2 = ctx.ccl(md, mask.shape, labels_out=ld)
    assert t2["n"] == rn and np.array_equal(ld.cpu().numpy().view(np.uint32), rl)


@pytest.mark.parametrize("shape,kind,p", [
    ((40, 64, 128), "blobs", 0), ((33, 70, 90), "blobs", 0), ((20, 50, 61), "bernoulli", 0.08),
    ((16, 40, 200), "bernoulli", 0.5), ((8, 33, 257), "bernoulli", 0.3), ((64, 256, 256), "blobs", 0),
    ((5, 7, 3), "bernoulli", 0.4), ((1, 1, 1), "bernoulli", 1.0), ((3, 5, 1000), "bernoulli", 0.9),
    # rows of whole 32-voxel words take the word-per-lane init pass (W = 1 or Y = 1: the general one)
    ((10, 20, 64), "bernoulli", 0.5), ((6, 1, 64), "bernoulli", 0.7), ((7, 9, 32), "bernoulli", 0.9), ((5, 12, 96), "bernoulli", 0.05),
    ((3, 4, 2048), "bernoulli", 0.97),
])
def test_ccl_random(shape, kind, p):
    _compare(P.synth_mask(shape, 1003, kind=kind, p=p))


def test_ccl_adversarial():
    m = np.zeros((6, 6, 70), np.uint8)
    _compare(m)                                   # empty
    _compare(np.ones((4, 5, 67), np.uint8))       # full volume: one component
    m[2, 3, 40] = 1
    _compare(m)                                   # single voxel
    d = np.zeros((40, 40, 40), np.uint8)
    i = np.arange(40)
    d[i, i, i] = 1                                # pure diagonal: only 26-links
    d[i, i, 39 - i] = 1
    _compare(d)
    c = np.zeros((9, 9, 96), np.uint8)            # comb crossing 32-voxel word borders
    c[4, 4, :] = 1
    c[::2, 4, 31] = 1
    c[4, ::2, 64] = 1
    c[0, 0, 95] = 1
    _compare(c)
    s = np.zeros((12, 12, 12), np.uint8)          # spiral-ish: late merges
    s[::2, :, 0] = 1; s[:, ::2, 11] = 1; s[5, 5, :] = 1
    _compare(s)
    u = np.zeros((3, 40, 130), np.uint8)          # U shapes whose arms meet only at the bottom row
    u[1, :, ::4] = 1; u[1, 39, :] = 1
    _compare(u)


def test_ccl_csv_drop_last_component():
    """count_blobs.py:104 iterates range(1, N): the CSV has N-1 rows."""
    m = P.synth_mask((24, 40, 64), 7)
    lab, n, st = P.blob_table(m)
    from delivr_cfos_b200 import Context
    from delivr_cfos_b200.count_blobs import csv_text
    ctx = Context(0)
    t = ctx.ccl(m, m.shape)
    assert csv_text(t, t["n"]) == P.csv_text(st, n)
    assert csv_text(t, t["n"]).count("\n") - 1 == n - 1


def test_ccl_full_size_cfg3_properties():
    """BASELINE.json configs[2] at full size (1000x2048x2048 mask, 4.2e9 voxels - the oracle would take minutes):
    size-independent properties.  labels > 0 == mask; sum of counts == foreground; every label 1..N used; the labelling
    is idempotent (labelling `labels > 0` again reproduces labels and table); first voxels appear in raster order."""
    from delivr_cfos_b200 import Context
    from delivr_cfos_b200.synth import synth_mask_cuda
    ctx = Context(0)
    shape = (1000, 2048, 2048)
    mask = synth_mask_cuda(shape, 1003)
    labels = torch.empty(shape, dtype=torch.int32, device="cuda")
    t = ctx.ccl(mask, shape, labels_out=labels)
    n = t["n"]
    fg = int(mask.sum(dtype=torch.int64))
    assert n > 1_000_000 and 0.01 < fg / mask.numel() < 0.04
    assert int(t["voxel_counts"][1:].sum()) == fg and int(t["voxel_counts"][0]) == mask.numel() - fg
    assert int(t["voxel_counts"][1:].min()) >= 1
    for z0 in range(0, shape[0], 100):                 # labels > 0 exactly where the mask is set, all within 1..N
        l = labels[z0:z0 + 100]
        assert torch.equal(l > 0, mask[z0:z0 + 100] > 0)
        assert int(l.max()) <= n
    # first voxel of label k precedes the first voxel of label k+1 in raster order: bbox zmin is non-decreasing
    zmin = t["bounding_boxes"][1:, 0]
    assert (np.diff(zmin) >= 0).all()
    # coordinate sums against a direct reduction of one axis
    zs = sum(int((mask[z].sum(dtype=torch.int64)) * z) for z in range(0, shape[0]))
    assert int(t["sums"][1:, 0].sum()) == zs
    # idempotence
    mask2 = (labels > 0).to(torch.uint8)
    del mask
    labels2 = torch.empty(shape, dtype=torch.int32, device="cuda")
    t2 = ctx.ccl(mask2, shape, labels_out=labels2)
    assert t2["n"] == n and torch.equal(labels, labels2)
    for k in ("voxel_counts", "sums", "bounding_boxes"):
        assert np.array_equal(t[k], t2[k]), k


@pytest.mark.parametrize("shape,planes,host", [((23, 40, 50), 5, True), ((23, 40, 50), 5, False), ((16, 33, 65), 1, False),
                                               ((40, 64, 64), 16, True)])
def test_ccl_any_size_equals_single_call(shape, planes, host):
    """Volumes beyond one 32-bit label space (whole brain, count_blobs.py:61) are labelled in z sub-slabs and merged:
    forced here with a tiny label space; labels, N and the table must equal the single-call result bit for bit."""
    from delivr_cfos_b200 import Context
    from delivr_cfos_b200.slabs import ccl_any_size
    ctx = Context(0)
    mask = P.synth_mask(shape, 1003, kind="bernoulli", p=0.12)
    # long structures crossing many cuts, including one that only connects through a later sub-slab (U shape)
    mask[:, 3, 3] = 1
    mask[2:shape[0] - 1, 10, 10] = 1
    mask[2:shape[0] - 1, 10, 14] = 1
    mask[shape[0] - 2, 10, 10:15] = 1
    lab1 = np.empty(shape, dtype=np.uint32)
    t1 = ctx.ccl(mask, shape, labels_out=lab1)
    maxv = planes * shape[1] * shape[2]
    if host:
        labn = np.empty(shape, dtype=np.uint32)
        tn = ccl_any_size(ctx, mask, shape, labels_out=labn, max_voxels=maxv)
    else:
        ld = torch.empty(shape, dtype=torch.int32, device="cuda")
        tn = ccl_any_size(ctx, torch.from_numpy(mask).cuda(), shape, labels_out=ld, max_voxels=maxv)
        labn = ld.cpu().numpy().view(np.uint32)
    assert tn["n"] == t1["n"] and np.array_equal(labn, lab1)
    for k in ("voxel_counts", "sums", "bounding_boxes"):
        assert np.array_equal(tn[k], t1[k]), k
    assert np.array_equal(tn["centroids"], t1["centroids"], equal_nan=True)
    t0 = ccl_any_size(ctx, mask, shape, max_voxels=maxv)          # table only
    assert t0["n"] == t1["n"] and np.array_equal(t0["voxel_counts"], t1["voxel_counts"])
