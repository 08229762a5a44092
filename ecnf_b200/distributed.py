"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The hot path shards by independent units: trajectories / minibatch rows are split across ranks with no
data-path collective.  Only two exchanges exist, both tiny:
  * gradient all-reduce (sum) of the flat fp32 gradient + loss scalar before the fused Adam step
    (the reference has no collective at all: ecnf/utils/loop.py:104-106 is a vestigial pmap);
  * all-gather of the 5 ESS sufficient statistics per rank (setup_training.py:175-182) -- or of log_w itself.
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of global sample indices owned by `rank` (remainder spread over the first ranks).
    Noise is keyed by the global index, so any world size reproduces the single-GPU draw."""
    base, rem = divmod(global_batch, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def merge_ess_stats_list(stats: torch.Tensor) -> torch.Tensor:
    """[R, 5] per-rank {max w, sum e^(w-max), sum e^(2(w-max)), max(-w), sum e^(-w-max(-w))} -> merged [5]."""
    s = stats.to(torch.float64)
    mx = s[:, 0].max()
    nmx = s[:, 3].max()
    s1 = (s[:, 1] * torch.exp(s[:, 0] - mx)).sum()
    s2 = (s[:, 2] * torch.exp(2 * (s[:, 0] - mx))).sum()
    s3 = (s[:, 4] * torch.exp(s[:, 3] - nmx)).sum()
    return torch.stack([mx, s1, s2, nmx, s3]).to(stats.dtype)


def merge_ess_stats(local: torch.Tensor) -> torch.Tensor:
    """All-gather the per-rank statistics and merge them (identical result on every rank)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    buf = torch.empty(world, local.numel(), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, local.contiguous()) if local.is_cuda else dist.all_gather(
        list(buf.unbind(0)), local.contiguous())
    return merge_ess_stats_list(buf)


def gather_log_weights(local: torch.Tensor) -> torch.Tensor:
    """All-gather per-rank log-weights (equal shard sizes) into one [world * n] vector, rank-major."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if local.is_cuda:
        out = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous())
    return torch.cat(parts)


def make_grad_allreduce(world: int) -> Callable[[torch.Tensor, torch.Tensor], torch.Tensor]:
    """Hook for flow_matching_update_fn: every rank computed sum-over-its-rows / (global_B * D), so a plain SUM
    all-reduce of the flat gradient and of the loss gives the global-batch mean on every rank."""
    def hook(flat_grad: torch.Tensor, loss: torch.Tensor) -> torch.Tensor:
        if world > 1:
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
            loss = loss.clone()
            dist.all_reduce(loss, op=dist.ReduceOp.SUM)
        return loss
    return hook
