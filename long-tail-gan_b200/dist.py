"""User-sharded data parallelism: one process per GPU, contiguous user ranges per rank, global batch =
concatenation of the rank slices. Pure host logic -- the collectives themselves are torch.distributed calls issued by
engine.GanEngine.run_d_step / run_g_step (NCCL on GPUs; the CPU tests drive the same helpers over gloo)."""
import os


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n_batches, rank, world):
    """Contiguous block of global batches owned by `rank`: (first, count). Every rank gets the same count (the last
    n_batches % world batches are dropped so that collectives stay aligned)."""
    per = n_batches // world
    return rank * per, per


def global_step_layout(batch_size, world):
    """One global step = `world` consecutive local batches; B_global is what the mean losses are normalised by
    (MultiVAE.py:110-112,161: reduce_mean over the batch)."""
    return dict(B_local=batch_size, B_global=batch_size * world)


def allreduce_sum_(tensors, group=None):
    """In-place SUM all-reduce of a list of tensors (gradients are already divided by B_global, so SUM is the mean)."""
    import torch.distributed as dist
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tensors
