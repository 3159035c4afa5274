"""One rank of a multi-GPU ``run_inference`` started by the single-process caller (inference._spawn_ranks):
``python -m delivr_cfos_b200.inference.rank_main payload.json`` with RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* set,
i.e. exactly what ``torchrun`` provides - the rank re-enters run_inference, which takes its one-rank-per-GPU branch."""
import json
import sys


def main():
    with open(sys.argv[1]) as f:
        kw = json.load(f)
    from delivr_cfos_b200.inference.inference import run_inference
    kw["stack_shape"] = tuple(kw["stack_shape"])
    kw["crop_size"] = tuple(kw["crop_size"])
    run_inference(**kw)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
