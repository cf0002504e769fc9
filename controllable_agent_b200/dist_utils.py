"""Host-side helpers of the data-parallel layout (DESIGN.md "Multi-GPU"): device-agnostic, so that the world_size > 1 logic is
testable with the gloo backend on CPU.  The compute itself has no CPU route."""
from __future__ import annotations

import typing as tp

import torch

# metrics that are per-rank means (or replicated values): averaged over ranks; every other entry of the metrics block is a
# partial sum over the rank's rows of a global-batch quantity: summed over ranks
MEAN_KEYS = frozenset({"target_M", "M1", "F1", "B", "B_norm", "z_norm", "orth_linf", "orth_l2"})


def shard_layout(global_batch: int, world: int, rank: int) -> tp.Tuple[int, int]:
    """(rows of this rank, global index of its first row): contiguous equal row blocks in rank order — the order
    all_gather_into_tensor concatenates the per-rank exchange blocks in."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} / world {world}")
    if global_batch % world:
        raise ValueError(f"batch_size {global_batch} must be divisible by the world size {world}")
    local = global_batch // world
    return local, rank * local


def shard_episodes(n_episodes: int, world: int, rank: int) -> tp.Tuple[int, int]:
    """[first, last) episode indices of this rank's replay shard: equal shards (remainder episodes are dropped so that uniform
    sampling inside every shard is uniform sampling over the union)."""
    per = n_episodes // world
    return rank * per, (rank + 1) * per


def reduce_metrics(m: tp.Mapping[str, float], keys: tp.Sequence[str], world: int, device: tp.Union[str, torch.device]) -> tp.Dict[str, float]:
    """Per-rank metric blocks -> global values with ONE all-reduce (fp64)."""
    import torch.distributed as dist
    t = torch.tensor([m[k] / world if k in MEAN_KEYS else m[k] for k in keys], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t)
    return {k: float(v) for k, v in zip(keys, t.tolist())}
