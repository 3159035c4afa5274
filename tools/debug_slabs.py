import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import unet_ref
from delivr_cfos_b200 import Context, slabs
from delivr_cfos_b200.synth import synth_volume_cuda
shape, roi, world, tta = (100, 80, 70), (32, 32, 32), 3, False
ctx = Context(0); ctx.load_weights(unet_ref.random_state_dict(4))
vol = synth_volume_cuda(shape, 77, roi=roi, blobs_per_mvox=2500.0)
v32 = vol.to(torch.int32); v32[:shape[0], :shape[1], :shape[2]].clamp_(min=1); v32[:3] = 0; v32[:, :, :6] = 0
vol = v32.to(torch.uint16); shape_pad = tuple(vol.shape)
def single():
    b1 = torch.empty(shape, dtype=torch.uint8, device="cuda"); a1 = torch.empty(shape_pad, dtype=torch.float32, device="cuda")
    ctx.segment(vol, shape_pad, shape, roi, b1, tta=tta, erosion_block_planes=17, avg_logits_out=a1)
    return b1, a1
b1, a1 = single(); b2, a2 = single()
print("single twice: bin diff", int((b1 != b2).sum()), "avg maxdiff", float((a1 - a2).abs().max()))
plan = slabs.SlabPlan(shape, roi, 0.5, world)
print("layers", plan.layers, [plan.rank(r) for r in range(world)])
workers = [slabs.CudaSlabWorker(ctx, plan, r, lambda a, b: vol[a:b].contiguous(), tta=tta, erosion_block_planes=17) for r in range(world)]
active = [w.accumulate() for w in workers]
for r in range(world):
    info = plan.rank(r); q = plan._next_nonempty(r)
    if info["send"] is not None and q is not None:
        g0, g1 = info["send"]; workers[q].add_planes(g0, g1, workers[r].acc_planes(g0, g1))
ag = np.concatenate(active)
# compare averaged logits on owned planes
for r, w in enumerate(workers):
    z0, z1 = w.info["slab"]; o0, o1 = w.info["own"]
    acc = w.acc.clone()
    ctx.seg_average(acc, z1 - z0, z0, plan.shape_pad, plan.roi, plan.overlap, ag, passes=1)
    avg = acc.view(torch.float32)[o0 - z0:o1 - z0]
    d = (avg - a1[o0:o1]).abs()
    print("rank", r, "own", (o0, o1), "avg maxdiff", float(d.max()), "planes with diff", torch.nonzero(d.amax(dim=(1, 2)) > 0).flatten().tolist()[:20])
for w in workers: w.finalise(ag)
bN = torch.cat([w.binaries for w in workers])
diff = (bN != b1)
print("bin diff", int(diff.sum()), "planes", torch.nonzero(diff.sum(dim=(1, 2))).flatten().tolist())
