"""Painter oracle and host logic against the golden of the unmodified reference blob_highlighter (no GPU)."""
import numpy as np
import pytest

from oracle import paint_ref
from paint_common import load_paint_golden, random_paint_case


def test_oracle_restatement_reproduces_reference_golden():
    g = load_paint_golden()
    stats = {"bounding_boxes": g["bounding_boxes"].copy()}
    ids = g["cells"]["connected_component_id"].to_numpy()
    r, gr, b = paint_ref.highlight_ref(g["mask"], stats, ids, g["cells"][["red", "green", "blue"]].to_numpy(), g["stack_shape"])
    assert np.array_equal(r, g["red"]) and np.array_equal(gr, g["green"]) and np.array_equal(b, g["blue"])
    # second pass of the reference: boxes were already grown once by pad_bb (in place)
    (reg,) = paint_ref.highlight_ref(g["mask"], stats, ids, g["cells"]["graph_order"].to_numpy(), g["stack_shape"], dtype=np.uint16)
    assert np.array_equal(reg, g["region"])
    assert not np.array_equal(stats["bounding_boxes"], g["bounding_boxes"])
    # the golden really exercises "last box wins": some foreground voxels do not carry their own blob's colour
    assert int((g["red"] > 0).sum()) < int(g["mask"].sum())


def test_product_box_sequence_plus_last_box_wins_equals_golden():
    """delivr_cfos_b200.blob_highlighter.padded_boxes (host logic of the product) + the literal box loop."""
    from delivr_cfos_b200.blob_highlighter import padded_boxes
    g = load_paint_golden()
    stats = {"bounding_boxes": g["bounding_boxes"].copy()}
    ids = g["cells"]["connected_component_id"].to_numpy()
    boxes = padded_boxes(stats, ids, g["stack_shape"])
    r, gr, b = paint_ref.paint_boxes_ref(g["mask"], boxes, g["cells"][["red", "green", "blue"]].to_numpy(), np.uint8)
    assert np.array_equal(r, g["red"]) and np.array_equal(gr, g["green"]) and np.array_equal(b, g["blue"])
    boxes2 = padded_boxes(stats, ids, g["stack_shape"])
    (reg,) = paint_ref.paint_boxes_ref(g["mask"], boxes2, g["cells"]["graph_order"].to_numpy(), np.uint16)
    assert np.array_equal(reg, g["region"])


@pytest.mark.parametrize("dups", [False, True])
def test_padded_boxes_matches_sequential_pad_bb(dups):
    from delivr_cfos_b200.blob_highlighter import padded_boxes
    rng = np.random.default_rng(5)
    stack_shape = (1, 1, 20, 30, 25)
    lo = np.stack([rng.integers(0, s, 40) for s in stack_shape[2:]], 1)
    hi = np.minimum(lo + rng.integers(0, 4, lo.shape), np.array(stack_shape[2:]) - 1)
    bb = np.empty((40, 6), dtype=np.int64)
    bb[:, 0::2], bb[:, 1::2] = lo, hi
    ids = rng.permutation(40)[:25]
    if dups:
        ids = np.concatenate([ids, ids[:5], ids[:2]])
    a, b = {"bounding_boxes": bb.copy()}, {"bounding_boxes": bb.copy()}
    got = padded_boxes(a, ids, stack_shape)
    want = np.stack([paint_ref.pad_bb(b["bounding_boxes"][i], stack_shape).copy() for i in ids])
    assert np.array_equal(got, want) and np.array_equal(a["bounding_boxes"], b["bounding_boxes"])


def test_last_box_wins_formulation():
    """out[v] = mask[v] * value[last box containing v]: the formulation the CUDA kernels implement."""
    mask, boxes, values = random_paint_case((9, 14, 11), 30, 3)
    want = paint_ref.paint_boxes_ref(mask, boxes, values, np.uint8)
    owner = np.zeros(mask.shape, dtype=np.int64)
    for k, bb in enumerate(boxes):
        owner[bb[0]:bb[1], bb[2]:bb[3], bb[4]:bb[5]] = k + 1
    for c in range(3):
        v = np.where(owner > 0, mask.astype(np.int64) * values[np.maximum(owner - 1, 0), c], 0).astype(np.uint8)
        assert np.array_equal(v, want[c])
