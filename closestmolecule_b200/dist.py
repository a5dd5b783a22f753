"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests).

The reference's own data-parallel axis is the sweep over independent problem instances
(concentrations c = 1..9 in advection_diffusion_mixed.py:273-334; 4 instances per least-squares
evaluation in correct_boundary_finite_element.py:76-81).  Instances are sharded round-robin over the
ranks, solved without any data-path collective, and their per-instance results (the `results` rows of
advection_diffusion_mixed.py:268) are gathered in instance order on every rank.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init(backend=None):
    """Initialise from the torchrun environment; returns (rank, world_size, local_rank)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def shard_instances(items, rank, world):
    """Round-robin shard: rank r gets items r, r+world, ... (with their global indices)."""
    return [(i, it) for i, it in enumerate(items) if i % world == rank]


def gather_rows(local_rows, n_total, width, device="cpu"):
    """local_rows: [(global_index, row[width])].  Returns the [n_total, width] float64 table on
    every rank (rows of instances no rank solved stay NaN)."""
    table = torch.full((n_total, width), float("nan"), dtype=torch.float64, device=device)
    mask = torch.zeros(n_total, dtype=torch.float64, device=device)
    for i, row in local_rows:
        table[i] = torch.as_tensor(row, dtype=torch.float64, device=device)
        mask[i] = 1.0
    if dist.is_initialized() and dist.get_world_size() > 1:
        filled = torch.nan_to_num(table, nan=0.0)
        dist.all_reduce(filled)
        dist.all_reduce(mask)
        table = torch.where(mask[:, None] > 0, filled, torch.full_like(filled, float("nan")))
    return table


def max_over_ranks(values, device="cpu"):
    """Elementwise MAX over ranks of a list of python floats (timings)."""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def sweep(instances, solve_one, width, device="cpu"):
    """Run solve_one(instance) -> row[width] for this rank's share and gather the table."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    rows = [(i, solve_one(it)) for i, it in shard_instances(instances, rank, world)]
    return gather_rows(rows, len(instances), width, device)
