"""Sharding of a Cartesian grid over the GPUs of one box.

Cells are independent on this path (the reference even declines to look at a neighbour's facet,
impl_unfitted_domain.cpp:399-401), so rank r of W owns the contiguous flat-id range made of the layers
[k0, k1) of the last direction (direction 0 is the fastest index, cart_grid_tp / tensor_index_tp.cpp:130-144).
The only exchange is one all-gather of each rank's totals, from which every rank derives the global offset of
its cut cells and of its interior / boundary points in the assembled containers.
"""
from __future__ import annotations

import numpy as np


def slab_range(n_layers: int, rank: int, world: int) -> tuple[int, int]:
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return n_layers * rank // world, n_layers * (rank + 1) // world


def slab_cell_range(n_cells_dir, rank: int, world: int) -> tuple[int, int]:
    """Flat cell id range [lo, hi) of the rank's slab."""
    n = [int(v) for v in n_cells_dir]
    k0, k1 = slab_range(n[-1], rank, world)
    per_layer = int(np.prod(n[:-1], dtype=np.int64))
    return k0 * per_layer, k1 * per_layer


def exchange_totals(local_totals, group=None, device=None):
    """All-gathers one int64 vector per rank; returns (all_totals[W, k], offsets[k] of this rank).

    Uses torch.distributed (NCCL on GPUs over NVLink, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.asarray(local_totals, np.int64), device=device)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        allt = t.cpu().numpy()[None, :]
        return allt, np.zeros_like(allt[0])
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    allt = torch.stack(out).cpu().numpy()
    return allt, allt[:rank].sum(axis=0)
