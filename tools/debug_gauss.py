"""Diagnostics for the Gaussian-blend parity test: where the largest differences sit and what the neighbours look like."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from gpu_common import ctx_with
from oracle import pipeline_ref as P

ctx, sd, onet = ctx_with("random")
roi, shape = (32, 48, 32), (60, 100, 70)
vol = P.synth_volume(shape, 61, roi=roi)
sub = vol[:shape[0], :shape[1], :shape[2]]
sub[sub == 0] = 500
vol[:6] = 0
w = P.gaussian_importance_map(roi)
pred = lambda t: onet(t.cuda()).cpu()
ref = P.infer_average_weighted(vol, roi, 0.5, pred, w)
refc = P.infer_average_weighted(vol, roi, 0.5, pred, np.ones(roi, np.float32))
b = np.empty(shape, dtype=np.uint8)
mine = np.empty(vol.shape, dtype=np.float32)
ctx.segment(vol, vol.shape, shape, roi, b, overlap=0.5, blend_mode=1, avg_logits_out=mine)
const = np.empty(vol.shape, dtype=np.float32)
ctx.segment(vol, vol.shape, shape, roi, b, overlap=0.5, blend_mode=0, avg_logits_out=const)
mask = P.ccl_ref.erode6((vol[:shape[0], :shape[1], :shape[2]] > 0).astype(np.uint8), 30) > 0
sl = tuple(slice(0, s) for s in shape)
d = np.abs(mine[sl] - ref[sl]); dc = np.abs(const[sl] - refc[sl])
print("gauss: max on mask", d[mask].max(), "max anywhere", d.max(), "| const: max on mask", dc[mask].max(), "anywhere", dc.max())
print("quantiles on mask (gauss):", np.quantile(d[mask], [0.5, 0.99, 0.999, 0.9999]))
idx = np.argsort((d * mask).ravel())[-8:]
for i in idx:
    z, y, x = np.unravel_index(i, shape)
    print((z, y, x), "mine", mine[z, y, x], "ref", ref[z, y, x], "const mine", const[z, y, x], "const ref", refc[z, y, x])
