"""CPU restatement of the reference's blob painter (TEST INFRASTRUCTURE - never imported by the product).

Follows blob_highlighter.py:17-22 (pad_bb), :107-124 (R/G/B colouring loop), :143-151 (region-id loop) and
blob_depthmap.py:198-207 (depth loop) statement by statement, including two properties of the reference that a
cleaned-up implementation would lose:

* ``pad_bb`` mutates the row of ``stats['bounding_boxes']`` it is given (numpy row views), so a box that is
  visited twice - a second pass (region-id after RGB), or a duplicated ``connected_component_id`` - grows again;
* every assignment rewrites the whole box, so where boxes overlap the LAST one in order wins.

Pinned by tests/golden/p1_highlight.npz, produced by the unmodified ``blob_highlighter`` (oracle/make_golden_paint.py).
"""
import numpy as np


def pad_bb(bb, stack_shape):
    """blob_highlighter.py:17-22 - in place on the row it is given."""
    if bb[1] < stack_shape[2]:
        bb[1] += 1
    if bb[3] < stack_shape[3]:
        bb[3] += 1
    if bb[5] < stack_shape[4]:
        bb[5] += 1
    return bb


def paint_boxes_ref(mask, boxes, values, dtype):
    """The literal loop: for k in order, out_c[box_k] = mask[box_k] * values[k][c] (numpy cast on assignment)."""
    values = np.asarray(values, dtype=np.int64)
    values = values.reshape(len(boxes), values.shape[1] if values.ndim == 2 else 1)
    outs = [np.zeros(mask.shape, dtype=dtype) for _ in range(values.shape[1])]
    for k, bb in enumerate(boxes):
        sl = (slice(int(bb[0]), int(bb[1])), slice(int(bb[2]), int(bb[3])), slice(int(bb[4]), int(bb[5])))
        for c, o in enumerate(outs):
            with np.errstate(over="ignore"):
                o[sl] = (mask[sl].astype(np.int64) * values[k, c]).astype(dtype)
    return outs


def highlight_ref(bin_img, stats, cell_ids, colours, stack_shape, dtype=np.uint8):
    """blob_highlighter.py:107-124 / :143-151 with ``colours`` [n, nch] the per-cell values (red, green, blue) or
    (graph_order,).  ``stats['bounding_boxes']`` is mutated exactly like in the reference."""
    colours = np.asarray(colours, dtype=np.int64).reshape(len(cell_ids), -1)
    outs = [np.zeros(bin_img.shape, dtype=dtype) for _ in range(colours.shape[1])]
    for k, cc_id in enumerate(cell_ids):
        bb = stats["bounding_boxes"][cc_id]
        bb = pad_bb(bb, stack_shape)
        sl = (slice(bb[0], bb[1]), slice(bb[2], bb[3]), slice(bb[4], bb[5]))
        for c, o in enumerate(outs):
            with np.errstate(over="ignore"):
                o[sl] = (bin_img[sl].astype(np.int64) * colours[k, c]).astype(dtype)
    return outs
