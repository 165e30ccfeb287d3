"""Multi-GPU plumbing: envs shard embarrassingly (contiguous global index blocks, one process per GPU);
the only collective on the path is the end-of-rollout all-reduce of the episode-statistics vector."""
import os

import torch

from . import _capi


def shard_range(total_envs: int, rank: int, world_size: int):
    """Contiguous block of global env indices owned by `rank`: [first, first + count)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(total_envs), int(world_size))
    first = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return first, count


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def all_reduce_stats(stats: torch.Tensor) -> dict:
    """Sum the per-device statistics vector over all ranks (NCCL on device tensors, gloo on CPU tensors).
    Without an initialised process group the local values are returned."""
    s = stats.clone()
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(s, op=torch.distributed.ReduceOp.SUM)
    vals = s.cpu().tolist()
    return {k: int(v) for k, v in zip(_capi.STAT_NAMES, vals)}


def make_sharded_env(total_envs: int, **env_kwargs):
    """This rank's shard of a `total_envs`-battle job launched one process per GPU (torchrun: RANK / LOCAL_RANK /
    WORLD_SIZE): contiguous global env indices, device cuda:LOCAL_RANK, bot seeds derived from the GLOBAL index so
    that results do not depend on the number of GPUs (tests/test_gpu_full_size.py checks exactly that)."""
    from .env import FootsiesEnv
    rank, local_rank, world = env_rank_world()
    first, count = shard_range(total_envs, rank, world)
    torch.cuda.set_device(local_rank)
    return FootsiesEnv(num_envs=count, device=torch.device("cuda", local_rank), first_env_index=first, **env_kwargs)
