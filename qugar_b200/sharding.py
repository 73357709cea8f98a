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


def balanced_slabs(layer_weights, world: int) -> list[tuple[int, int]]:
    """Contiguous layer ranges [k0,k1) with (nearly) equal total weight -- use the cut cells per layer of the last direction.

    Equal-cell slabs are fine for a periodic TPMS but poor for, e.g., a sphere whose cut cells sit in the middle layers
    (SURVEY 8e).  Rank r takes [cuts[r], cuts[r+1]) where cuts[r] is the first layer at which the running weight reaches
    r/world of the total (nearest layer boundary); every rank gets at least one layer."""
    w = np.asarray(layer_weights, np.float64)
    n = len(w)
    if world < 1 or n < world:
        raise ValueError("need at least one layer per rank")
    total = float(w.sum())
    if total <= 0.0:
        return [slab_range(n, r, world) for r in range(world)]
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(cum, target, side="left"))
        if k > 0 and target - cum[k - 1] < cum[min(k, n)] - target:
            k -= 1                            # the layer boundary nearest to the target weight
        k = max(k, cuts[-1] + 1)              # at least one layer for the previous rank
        k = min(k, n - (world - r))           # and for every rank still to come
        cuts.append(k)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def exchange_layer_counts(local_cut_cells, n_cells_dir, group=None, device=None):
    """Cut cells per layer of the last direction, summed over the ranks (each rank passes the cut cells of ITS slab)."""
    import torch
    import torch.distributed as dist
    n = [int(v) for v in n_cells_dir]
    per_layer = int(np.prod(n[:-1], dtype=np.int64))
    counts = np.bincount(np.asarray(local_cut_cells, np.int64) // per_layer, minlength=n[-1]).astype(np.int64)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        t = torch.as_tensor(counts, device=device)
        dist.all_reduce(t, group=group)
        counts = t.cpu().numpy()
    return counts
