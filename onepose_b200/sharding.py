"""Object-sharded multi-GPU plumbing.

The path shards by (object, sequence): frames of different objects never interact, so every rank runs
the matcher on its own objects with NO collective on the data path.  What remains of the reference's
distributed code (src/utils/comm.py:141-215, a pickle gather over gloo used once for validation metrics)
is ONE fixed-size all_gather of per-rank records at the end.

  partition_lpt(costs, world)   static longest-processing-time-first assignment of work items to ranks
  gather_records(rec, group)    all_gather of a float64 record per rank -> [world, len(rec)] on every rank

Works with any torch.distributed backend (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.distributed as dist


def partition_lpt(costs: Sequence[float], world: int) -> list[list[int]]:
    """Greedy LPT: items sorted by decreasing cost, each to the currently least-loaded rank.
    Deterministic (ties broken by index / lowest rank) so every rank computes the same plan locally."""
    if world <= 0:
        raise ValueError("world must be positive")
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    loads = [0.0] * world
    plan: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        plan[r].append(i)
        loads[r] += float(costs[i])
    return plan


def object_cost(n_frames: int, n2d: int, n3d: int) -> float:
    """Cost model of one (object, sequence): frames x points (the GNN is linear in N2D + N3D)."""
    return float(n_frames) * float(n2d + n3d)


def gather_records(record: torch.Tensor, group=None) -> torch.Tensor:
    """record: 1-D float64 tensor (same length on every rank, on the backend's device) -> [world, len]."""
    if not dist.is_available() or not dist.is_initialized():
        return record[None].clone()
    world = dist.get_world_size(group)
    parts = [torch.empty_like(record) for _ in range(world)]
    dist.all_gather(parts, record.contiguous(), group=group)
    return torch.stack(parts, 0)


def hetero_job(seed: int = 5, n_objects: int = 80):
    """Stand-in for the reference's evaluation set (configs/experiment/test_GATsSPG.yaml:27-106: 80 (object, sequence) pairs looped
    by inference.py:185-198) when the data is not mounted -- SURVEY 8d: per object M ~ U[800, 2500] 3D points, N ~ U[300, 2000]
    query points, U[50, 400] frames.  Deterministic in `seed`; every rank computes the same list.
    Returns a list of dicts {id, M, N, frames, cost}."""
    import numpy as np
    rs = np.random.RandomState(seed)
    jobs = []
    for i in range(n_objects):
        M, N, F = int(rs.randint(800, 2501)), int(rs.randint(300, 2001)), int(rs.randint(50, 401))
        jobs.append({"id": i, "M": M, "N": N, "frames": F, "cost": object_cost(F, N, M)})
    return jobs
