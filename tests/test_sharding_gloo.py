"""world_size-2 (gloo, CPU) test of the sharded path's host logic: slab ranges + the totals all-gather."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

import oracle_lib as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    from qugar_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    breaks = [np.linspace(0, 1, n + 1)] * 3
    d = O.OracleDomain(3, O.K_SCHOEN, [1.3, 0.9, 1.1], breaks)
    lo, hi = sharding.slab_cell_range([n, n, n], rank, world)
    cut = d.cut_cells
    mine = cut[(cut >= lo) & (cut < hi)]
    r = d.quad_cells(mine, 3)
    b = d.quad_unf_bdry(mine, 3, False, False)
    allt, off = sharding.exchange_totals([len(mine), len(r.weights), len(b.weights)])
    q.put((rank, allt.tolist(), off.tolist(), len(cut)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_offsets_match_single_domain():
    n, world = 10, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    breaks = [np.linspace(0, 1, n + 1)] * 3
    d = O.OracleDomain(3, O.K_SCHOEN, [1.3, 0.9, 1.1], breaks)
    cut = d.cut_cells
    full_int = d.quad_cells(cut, 3)
    full_bnd = d.quad_unf_bdry(cut, 3, False, False)
    allt = np.array(res[0][1])
    assert np.array_equal(allt, np.array(res[1][1]))
    assert allt[:, 0].sum() == len(cut)
    assert allt[:, 1].sum() == len(full_int.weights)
    assert allt[:, 2].sum() == len(full_bnd.weights)
    assert res[0][2] == [0, 0, 0]
    assert res[1][2] == allt[0].tolist()
    # rank 1's point offset is where its first cell's points start in the single-domain rule
    first1 = int(np.searchsorted(cut, 10 * 10 * 5))
    assert res[1][2][1] == int(full_int.n_per[:first1].sum())


def test_slab_ranges_partition():
    from qugar_b200 import sharding
    for n in (7, 256, 512):
        for w in (1, 2, 3, 4, 8):
            r = [sharding.slab_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))


def test_balanced_slabs():
    from qugar_b200 import sharding
    # a sphere: cut cells per layer peak near the poles' rings, none outside the sphere
    n = 64
    z = (np.arange(n) + 0.5) / n
    w = np.where(np.abs(z - 0.5) < 0.4, 1.0 / np.sqrt(np.maximum(1e-3, 0.16 - (z - 0.5) ** 2)), 0.0)
    for world in (1, 2, 3, 8):
        slabs = sharding.balanced_slabs(w, world)
        assert slabs[0][0] == 0 and slabs[-1][1] == n and all(a[1] == b[0] for a, b in zip(slabs, slabs[1:]))
        assert all(k1 > k0 for k0, k1 in slabs)
        loads = np.array([w[k0:k1].sum() for k0, k1 in slabs])
        equal = np.array([w[sharding.slab_range(n, r, world)[0]:sharding.slab_range(n, r, world)[1]].sum() for r in range(world)])
        if world == 8:
            assert loads.max() < 0.8 * equal.max()         # equal-cell slabs leave the outer ranks idle here
        assert loads.max() <= w.sum() / world + w.max() + 1e-12
    assert sharding.balanced_slabs(np.zeros(8), 4) == [sharding.slab_range(8, r, 4) for r in range(4)]
    import pytest
    with pytest.raises(ValueError):
        sharding.balanced_slabs(np.ones(3), 4)


def _worker_balance(rank, world, port, n, q):
    import torch.distributed as dist
    from qugar_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    breaks = [np.linspace(0, 1, n + 1)] * 3
    d = O.OracleDomain(3, O.K_SPHERE, [0.3, .5, .5, .35], breaks)
    lo, hi = sharding.slab_cell_range([n, n, n], rank, world)
    cut = d.cut_cells
    counts = sharding.exchange_layer_counts(cut[(cut >= lo) & (cut < hi)], [n, n, n])
    q.put((rank, counts.tolist(), sharding.balanced_slabs(counts, world)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_layer_counts_and_rebalance():
    n, world = 12, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_balance, args=(r, world, port, n, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = O.OracleDomain(3, O.K_SPHERE, [0.3, .5, .5, .35], [np.linspace(0, 1, n + 1)] * 3)
    ref = np.bincount(d.cut_cells // (n * n), minlength=n)
    assert res[0][1] == ref.tolist() == res[1][1]
    assert res[0][2] == res[1][2]
    (a0, a1), (b0, b1) = res[0][2]
    assert a0 == 0 and a1 == b0 and b1 == n
    # the sphere sits low in z (center 0.35): the balanced cut is below the middle layer
    assert a1 < n // 2 and abs(ref[:a1].sum() - ref[a1:].sum()) <= ref.max()
