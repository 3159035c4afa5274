"""z-slab sharding: N virtual slabs on one GPU must reproduce the single-slab result bit for bit
(binaries, labels, table), because the blend accumulates in fixed point and the table merge is integer-exact."""
import numpy as np
import pytest
import torch

from oracle import pipeline_ref as P, unet_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,roi,world,tta,blend", [((100, 80, 70), (32, 32, 32), 3, False, 0), ((70, 64, 64), (32, 48, 32), 2, False, 0),
                                                       ((150, 60, 50), (32, 32, 32), 4, True, 0), ((40, 64, 64), (32, 32, 32), 5, False, 0),
                                                       ((100, 80, 70), (32, 32, 32), 3, False, 1), ((120, 60, 50), (32, 48, 32), 4, True, 1)])
def test_virtual_slabs_equal_single_gpu(shape, roi, world, tta, blend):
    from delivr_cfos_b200 import Context
    from delivr_cfos_b200 import slabs
    from delivr_cfos_b200.synth import synth_volume_cuda
    ctx = Context(0)
    ctx.load_weights(unet_ref.random_state_dict(4))
    vol = synth_volume_cuda(shape, 77, roi=roi, blobs_per_mvox=2500.0)
    v32 = vol.to(torch.int32)
    v32[:shape[0], :shape[1], :shape[2]].clamp_(min=1)
    v32[:3] = 0
    v32[:, :, :6] = 0
    vol = v32.to(torch.uint16)
    shape_pad = tuple(vol.shape)
    # single slab
    b1 = torch.empty(shape, dtype=torch.uint8, device="cuda")
    ctx.segment(vol, shape_pad, shape, roi, b1, tta=tta, erosion_block_planes=17, blend_mode=blend)
    l1 = torch.empty(shape, dtype=torch.int32, device="cuda")
    t1 = ctx.ccl(b1, shape, labels_out=l1)
    assert t1["n"] > 3 and int(b1.sum()) > 0
    # N virtual slabs
    plan = slabs.SlabPlan(shape, roi, 0.5, world)
    assert plan.shape_pad == shape_pad
    workers = [slabs.CudaSlabWorker(ctx, plan, r, lambda a, b: vol[a:b].contiguous(), tta=tta, erosion_block_planes=17, blend_mode=blend)
               for r in range(world)]
    table = slabs.run_virtual(workers, plan)
    bN = torch.cat([w.binaries for w in workers])
    lN = torch.cat([w.labels for w in workers])
    assert torch.equal(bN, b1)
    assert torch.equal(lN, l1)
    assert table["n"] == t1["n"]
    for k in ("voxel_counts", "sums", "bounding_boxes"):
        assert np.array_equal(table[k], t1[k]), k
    assert np.array_equal(table["centroids"], t1["centroids"], equal_nan=True)


def test_tta_weighted_passes_equal_13_explicit_passes():
    """inference.py:265-279 runs 13 passes (5 plain, 4 flip z, 4 flip y once the sub-resolution noise is dropped); the
    library evaluates 3 and blends them 5/4/4 times.  The int32 blend sums must be bit-identical to 13 explicit passes."""
    from delivr_cfos_b200 import Context
    from delivr_cfos_b200.synth import synth_volume_cuda
    ctx = Context(0)
    ctx.load_weights(unet_ref.random_state_dict(5))
    shape, roi = (64, 64, 64), (32, 32, 32)
    vol = synth_volume_cuda(shape, 78, roi=roi, blobs_per_mvox=2500.0)
    plan_windows = np.array([(z, y, x) for z in (0, 16, 32) for y in (0, 16, 32) for x in (0, 16, 32)], dtype=np.int32)
    def sched(flips):
        return np.concatenate([np.concatenate([plan_windows, np.full((len(plan_windows), 1), f, np.int32)], axis=1) for f in flips])
    acc13 = torch.zeros(tuple(vol.shape), dtype=torch.int32, device="cuda")
    acc3 = torch.zeros_like(acc13)
    torch.cuda.synchronize()
    ctx.seg_accumulate(vol, sched([0, 0, 2, 3, 0, 2, 3, 0, 2, 3, 0, 2, 3]), roi, acc13)
    ctx.seg_accumulate(vol, sched([0 | (4 << 8), 2 | (3 << 8), 3 | (3 << 8)]), roi, acc3)
    assert int(acc13.abs().sum()) > 0
    assert torch.equal(acc13, acc3)
