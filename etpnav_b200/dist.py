"""Data-parallel plumbing for the planner step: the batch axis is the only sharded axis, weights
are replicated, and the one collective is a SUM all-reduce of the flat gradient slice once per optimizer step
(the reference: DDP bucketed all-reduce, ss_trainer_ETP.py:211-212).  The mean over ranks is applied as
``grad_scale = 1/world`` inside the fused AdamW kernel, so no extra pass over the buffer is needed."""
import torch


def rank_seed(base: int, rank: int) -> int:
    """Each rank draws its own synthetic shard (mirrors per-rank environments, ss_trainer_ETP.py:159-163)."""
    return base + rank


def shard_batch(global_batch: int, world: int, rank: int):
    """Contiguous [start, end) rows of the global batch owned by ``rank`` (even split, remainder to low ranks)."""
    per, rem = divmod(global_batch, world)
    start = rank * per + min(rank, rem)
    return start, start + per + (1 if rank < rem else 0)


def allreduce_flat_(flat_grad: torch.Tensor, world: int) -> float:
    """In-place SUM all-reduce of a flat gradient slice; returns the scale that turns the sum into DDP's mean."""
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return 1.0 / world
