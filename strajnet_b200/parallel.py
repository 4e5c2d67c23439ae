"""Data-parallel inference over the GPUs of one box (BASELINE config 4): samples are independent
(no cross-sample op anywhere in STrajNet.call), so the batch is split contiguously by
rank, weights are replicated, and the ONLY collective is one all-gather of the output grids.

One process per GPU (torchrun); `torch.distributed` (NCCL over NVLink/NVSwitch, gloo in CPU tests)
is used purely as plumbing.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous sample range [lo, hi) of `rank`; the global batch must divide evenly (all-gather needs equal shards)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by {world} ranks")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def shard_inputs(inputs: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    B = next(iter(inputs.values())).shape[0]
    lo, hi = shard_bounds(B, rank, world)
    return {k: v[lo:hi] for k, v in inputs.items()}


def gather_outputs(local_out: torch.Tensor, group=None) -> torch.Tensor:
    """The single collective of the path: all-gather [B/world,256,256,32] -> [B,256,256,32] on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_out
    local_out = local_out.contiguous()
    full = torch.empty((world * local_out.shape[0],) + tuple(local_out.shape[1:]), dtype=local_out.dtype,
                       device=local_out.device)
    dist.all_gather_into_tensor(full, local_out, group=group)
    return full


class DataParallelSTrajNet:
    """`model(global_inputs)` on every rank -> the full [B,256,256,32] output on every rank."""

    def __init__(self, model, group=None):
        self.model, self.group = model, group

    def __call__(self, ogm, map_img, training=True, obs=None, occ=None, mapt=None, flow=None):
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        sh = shard_inputs(dict(ogm=ogm, map_img=map_img, obs=obs, occ=occ, flow=flow), rank, world)
        y = self.model(sh["ogm"], sh["map_img"], training=training, obs=sh["obs"], occ=sh["occ"], flow=sh["flow"])
        return gather_outputs(y, self.group)
