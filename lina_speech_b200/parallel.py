"""Batch-sharded generation across the GPUs of one node.

Sequences are independent (per-sequence state, no cross-sequence op anywhere in generate_batch), so rank r owns
sequences [r*B/W, (r+1)*B/W) with its own Cache and replicated weights.  The ONE exchange of the data path is
an all-gather of the sampled token ids per step so that every rank holds the global token / stop bookkeeping
(model/modeling_lina.py:168-173 early exit, rank-0 post-processing and codec decode).  NCCL over NVLink on the
GPU box; the same code runs on gloo/CPU tensors, which is how tests/test_parallel.py covers it.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice of the batch owned by ``rank`` (first ``global_batch % world`` ranks get one more)."""
    base, extra = divmod(global_batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_tokens(q_local: torch.Tensor, group=None) -> torch.Tensor:
    """[Q, b_local, 1] int64 ids of this rank -> [Q, b_global, 1], rank-major (equal b_local on every rank)."""
    world = dist.get_world_size(group)
    if world == 1:
        return q_local
    Q, b, n = q_local.shape
    out = torch.empty(world * Q, b, n, dtype=q_local.dtype, device=q_local.device)     # concatenated along dim 0
    dist.all_gather_into_tensor(out, q_local.contiguous(), group=group)
    return out.view(world, Q, b, n).permute(1, 0, 2, 3).reshape(Q, world * b, n)
