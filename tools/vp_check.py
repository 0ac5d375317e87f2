"""Catalog-sharded engine vs single-GPU engine on the same batch (long-tail-gan_b200/vp_check.py).
    python tools/vp_check.py                       (one shard, one GPU)
    torchrun --nproc-per-node N tools/vp_check.py  (N item shards)"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    vpc = importlib.import_module("long-tail-gan_b200.vp_check")
    args = [a for a in sys.argv[1:] if not a.endswith(".json")]
    graphs = "graphs" in args     # capture phase A / D / G (NCCL collectives inside) as CUDA graphs; the second step is a replay
    args = [a for a in args if a != "graphs"]
    I = int(args[0]) if args else 2400
    res = vpc.run_check(I, 96, rank, world, use_graphs=graphs)
    if rank == 0:
        print(json.dumps(res, indent=1))
        for a in sys.argv[1:]:
            if a.endswith(".json"):
                json.dump(res, open(a, "w"), indent=1)
        assert res["ok"], "catalog-sharded step differs from the single-GPU step"
        print("vp_check ok")
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
