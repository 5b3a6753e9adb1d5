"""Multi-GPU plumbing: the env batch shards across ranks with no data-path collective
(every reference env owns its own b2World, reference multi_car_racing.py:138); NCCL/gloo only
gathers throughput counters at report time."""
import os


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_envs(total_envs, rank, world):
    """Contiguous block of env indices owned by `rank`: (first, count).  All agents of an env stay
    on one rank (cars of one env interact)."""
    base, rem = divmod(int(total_envs), int(world))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def reduce_throughput(frames, elapsed_ms, device=None):
    """Whole-job numbers from per-rank counters: (sum of frames, max of elapsed).  Uses the
    initialised torch.distributed group (nccl on GPUs, gloo in the CPU tests); identity when the
    job is a single process."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(frames), float(elapsed_ms)
    t_sum = torch.tensor([float(frames)], dtype=torch.float64, device=device)
    t_max = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t_sum, op=dist.ReduceOp.SUM)
    dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    return float(t_sum.item()), float(t_max.item())
