"""Multi-GPU sharding of the hot path: frames are independent units, so the batch is
sharded BY FRAME across ranks (one process per GPU) with no data-path collective. torch.distributed
is used only for the barrier and for the max-over-ranks timing / sum-over-ranks counters that the
benchmark reports. Works with the `nccl` backend on GPUs and `gloo` on CPU (tests)."""
import os


def rank_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def frames_for_rank(rank, world, frames_per_rank):
    """Global frame ids [rank*B, (rank+1)*B): weak scaling, every rank owns B distinct frames."""
    if not (0 <= rank < world) or frames_per_rank < 0:
        raise ValueError("bad rank / world / frames_per_rank")
    return list(range(rank * frames_per_rank, (rank + 1) * frames_per_rank))


def frame_owner(frame_id, world, frames_per_rank):
    return (frame_id // frames_per_rank) % world


def max_over_ranks(value, device=None):
    """max over ranks of a python float (elapsed ms); identity when not distributed."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(frames_this_rank, elapsed_s_this_rank, device=None):
    """Whole-job frames/s = (sum of frames over ranks) / (max elapsed over ranks)."""
    total = sum_over_ranks(frames_this_rank, device)
    tmax = max_over_ranks(elapsed_s_this_rank, device)
    return total / tmax
