#!/usr/bin/env python
"""Real-NCCL parity check of the z-slab driver (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/nccl_parity.py

Every rank builds the same seeded volume; the distributed run (halo exchange of the int32 logit sums, boundary label
plane, all-gathers over NCCL) must reproduce rank 0's single-GPU dlv_segment + dlv_ccl result bit for bit: binaries,
labels, N and the whole table."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from delivr_cfos_b200 import Context, slabs
    from delivr_cfos_b200.synth import synth_volume_cuda
    from oracle import unet_ref          # seeded random-init weights of the architecture (a test tool, like tests/)
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = Context(local)
    ctx.load_weights(unet_ref.random_state_dict(4))
    ok = True
    for shape, roi, tta in [((100, 80, 70), (32, 32, 32), False), ((150, 96, 64), (32, 48, 32), True)]:
        vol = synth_volume_cuda(shape, 77, roi=roi, blobs_per_mvox=2500.0, device=dev)
        v32 = vol.to(torch.int32)
        v32[:shape[0], :shape[1], :shape[2]].clamp_(min=1)
        v32[:3] = 0
        v32[:, :, :6] = 0
        vol = v32.to(torch.uint16)
        torch.cuda.synchronize()
        plan = slabs.SlabPlan(shape, roi, 0.5, world)
        w = slabs.CudaSlabWorker(ctx, plan, rank, lambda a, b: vol[a:b].contiguous(), tta=tta, erosion_block_planes=17)
        table = slabs.run_distributed(w, plan, slabs.TorchComm())
        torch.cuda.synchronize()
        # gather the owned planes on rank 0
        sizes = [plan.rank(r)["own_real"][1] - plan.rank(r)["own_real"][0] for r in range(world)]
        bs = [torch.empty((max(s, 0), shape[1], shape[2]), dtype=torch.uint8, device=dev) for s in sizes]
        ls = [torch.empty((max(s, 0), shape[1], shape[2]), dtype=torch.int32, device=dev) for s in sizes]
        for r in range(world):
            if sizes[r] > 0:
                dist.broadcast(w.binaries if r == rank else bs[r], src=r)
                dist.broadcast(w.labels if r == rank else ls[r], src=r)
        if rank == 0:
            bs[0], ls[0] = w.binaries, w.labels
            bN, lN = torch.cat(bs), torch.cat(ls)
            b1 = torch.empty(shape, dtype=torch.uint8, device=dev)
            ctx.segment(vol, tuple(vol.shape), shape, roi, b1, tta=tta, erosion_block_planes=17)
            l1 = torch.empty(shape, dtype=torch.int32, device=dev)
            t1 = ctx.ccl(b1, shape, labels_out=l1)
            good = (torch.equal(bN, b1) and torch.equal(lN, l1) and table["n"] == t1["n"] and t1["n"] > 3 and
                    all(np.array_equal(table[k], t1[k]) for k in ("voxel_counts", "sums", "bounding_boxes")) and
                    np.array_equal(table["centroids"], t1["centroids"], equal_nan=True))
            print(f"nccl parity world={world} shape={shape} tta={tta}: {'OK' if good else 'MISMATCH'} "
                  f"(components {t1['n']}, foreground {int(b1.sum())})", flush=True)
            ok = ok and good
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
