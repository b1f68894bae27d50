"""Batch sharding across ranks: MPC instances are independent, so rank r owns a contiguous block
and the only exchange is an all-gather of the command block cmd[B, nu].  Pure host logic, backend agnostic (NCCL on the
GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_bounds(total: int, world: int, rank: int):
    """Contiguous blocks of ceil(total/world) instances; trailing ranks may own fewer (or zero)."""
    per = -(-total // world)
    lo = min(rank * per, total)
    hi = min(lo + per, total)
    return lo, hi


def gather_commands(local_cmd, total: int, group=None):
    """All-gather of the per-rank command blocks into cmd[total, nu] on every rank (torch tensors; device follows the
    input).  Ragged tails are padded to the common block size and trimmed after the collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    per = -(-total // world)
    nu = local_cmd.shape[1]
    pad = torch.zeros((per, nu), dtype=local_cmd.dtype, device=local_cmd.device)
    pad[: local_cmd.shape[0]] = local_cmd
    out = torch.empty((world * per, nu), dtype=local_cmd.dtype, device=local_cmd.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return out[:total]
