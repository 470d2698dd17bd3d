"""Multi-GPU host logic: batch-sharded encoding + the one exchange step of the path.

Sentences are independent, so the path shards by batch with no collective during the encode
(SURVEY §8e): rank r encodes the contiguous block ``shard_bounds(N, world, r)`` and a single
``all_gather`` assembles the ``[N, 1024]`` embedding matrix on every rank (NCCL over NVLink on GPUs;
the same code runs under ``gloo`` on CPU tensors for the world-size-2 tests).
"""

from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
from torch import Tensor


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first ``n % world`` ranks get one extra item."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_rows(local: Tensor, n_total: int, group=None) -> Tensor:
    """All-gather row blocks of unequal size (as produced by ``shard_bounds``) into ``[n_total, D]``."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    base, extra = divmod(n_total, world)
    cap = base + (1 if extra else 0)
    d = local.shape[1]
    padded = torch.zeros((cap, d), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * cap, d), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    if extra == 0:
        return out
    parts = []
    for r in range(world):
        s, e = shard_bounds(n_total, world, r)
        parts.append(out[r * cap : r * cap + (e - s)])
    return torch.cat(parts, dim=0)


def encode_sharded(predict: Callable[[Sequence[str]], Tensor], sentences: Sequence[str], group=None) -> Tensor:
    """Every rank passes the SAME sentence list; each encodes its shard with ``predict`` (e.g.
    ``lambda s: pipeline.predict(s, "eng_Latn", batch_size=...)``) and all get the ``[N, D]`` matrix in
    input order."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    s, e = shard_bounds(len(sentences), world, rank)
    local = predict(list(sentences[s:e])) if e > s else None
    # collectives run on the rank's CUDA device under NCCL (also for a rank whose shard is empty), on CPU under gloo
    coll_dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    width = torch.tensor([local.shape[1] if local is not None else 0], dtype=torch.int64, device=coll_dev)
    dist.all_reduce(width, op=dist.ReduceOp.MAX, group=group)
    if local is None:  # empty shard: the feature width comes from the other ranks
        local = torch.zeros((0, int(width.item())), dtype=torch.float32, device=coll_dev)
    return gather_rows(local, len(sentences), group)
