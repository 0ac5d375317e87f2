"""N-GPU correctness check of the data-parallel step against the single-GPU engine on the same global batch
(long-tail-gan_b200/dp_check.py).   torchrun --nproc-per-node 2 tools/dp_check.py [out.json]"""
import importlib
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dpc = importlib.import_module("long-tail-gan_b200.dp_check")
    tabs = dpc.small_problem(world)
    res = dpc.run_check(tabs, tabs["n_items"], 50, rank, world)
    if rank == 0:
        print(json.dumps(res, indent=1))
        if len(sys.argv) > 1:
            json.dump(res, open(sys.argv[1], "w"), indent=1)
        assert res["ok"], "data-parallel step differs from the single-GPU step"
        print("dp_check ok")
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
