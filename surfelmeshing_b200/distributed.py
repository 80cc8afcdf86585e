"""Multi-GPU plumbing: replicas only.

A single scene's surfel cloud does not shard (association scatters every surfel into one
shared image-space raster, neighbour links are arbitrary cross references), so
the pipeline scales as independent RGB-D streams, one per GPU, with NO collective on the data
path. torch.distributed is used only for the start barrier and for reducing the timing
(max over ranks) and the work done (sum over ranks) of a benchmark run.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass
class RankInfo:
    rank: int
    local_rank: int
    world_size: int

    @property
    def is_distributed(self) -> bool:
        return self.world_size > 1


def rank_info_from_env() -> RankInfo:
    return RankInfo(int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
                    int(os.environ.get("WORLD_SIZE", "1")))


def init_process_group(info: RankInfo, backend: str = "nccl"):
    if info.is_distributed and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29512")
        dist.init_process_group(backend=backend, rank=info.rank, world_size=info.world_size)


def stream_ids_for_rank(info: RankInfo, streams_total: int):
    """Stream s is processed by rank s % world_size (one stream per GPU in config 4)."""
    return [s for s in range(streams_total) if s % info.world_size == info.rank]


def barrier(info: RankInfo, device=None):
    if info.is_distributed:
        if device is not None and torch.device(device).type == "cuda":
            dist.barrier(device_ids=[torch.device(device).index or 0])
        else:
            dist.barrier()


def gather_values(info: RankInfo, value: float, device="cpu"):
    """The value of every rank, in rank order (diagnostics: how far the replicas spread)."""
    if not info.is_distributed:
        return [float(value)]
    t = torch.zeros(info.world_size, dtype=torch.float64, device=device)
    t[info.rank] = value
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.tolist()]


def aggregate(info: RankInfo, elapsed_ms: float, units: float, device="cpu"):
    """Returns (max elapsed over ranks, total units over ranks): whole-job throughput is
    total units / max time."""
    if not info.is_distributed:
        return float(elapsed_ms), float(units)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    u = torch.tensor([units], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())
