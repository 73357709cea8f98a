"""BASELINE.json configs[4]: 3-D Schwarz-diamond TPMS on a 512^3 grid, 6 pts/dir, sharded by contiguous cell ranges (z-slabs)
over the GPUs of one box, NCCL all-gather of per-rank totals -> global offsets.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/run_c5_sharded.py

Every rank classifies and integrates its slab (device-resident rules), reduces its partial integrals on the GPU with torch
(DLPack-free: the device pointers of the rule are wrapped with torch.from_blob-like ctypes access via cudaMemcpy to torch
tensors), and the job checks size-independent properties of the assembled result: |Omega| = 1/2 (symmetry of the Schwarz
diamond), the per-rank counts add up, offsets are the exclusive scan of the counts."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qugar_b200 import cpp, sharding  # noqa: E402

GRID, NPTS, PERIODS = 512, 6, [2.0, 2.0, 2.0]


def device_array(ptr, n, dtype, device):
    """Copy n elements at device pointer ptr into a torch tensor on the same device (no host round trip)."""
    t = torch.empty(n, dtype=dtype, device=device)
    if n:
        cpp._check(cpp._load().qg_memcpy_device(C.c_void_p(t.data_ptr()), C.c_void_p(ptr), C.c_int64(n * t.element_size()),
                                                device.index))
    return t


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    lib = cpp._load()
    vp = C.c_void_p
    lib.qg_domain_create_slab.argtypes = [vp, vp, C.c_int, C.c_int64, C.c_int64, C.POINTER(vp)]
    lib.qg_domain_cut_cells_device.argtypes = [vp, C.POINTER(vp)]
    full = np.linspace(0.0, 1.0, GRID + 1)
    grid = cpp.create_cart_grid([full] * 3)
    func = cpp.create_SchwarzDiamond(PERIODS)
    k0, k1 = sharding.slab_range(GRID, rank, world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dom, dc, r1, r2 = vp(), vp(), vp(), vp()
    cpp._check(lib.qg_domain_create_slab(func._h, grid._h, local, k0, k1, C.byref(dom)))
    cnt = (C.c_int64 * 4)()
    lib.qg_domain_counts(dom, cnt)
    cpp._check(lib.qg_domain_cut_cells_device(dom, C.byref(dc)))
    cpp._check(lib.qg_quad_cells_device(dom, dc, NPTS, C.byref(r1)))
    cpp._check(lib.qg_quad_unf_bdry_device(dom, dc, NPTS, C.byref(r2)))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ne, n1, n2 = C.c_int64(), C.c_int64(), C.c_int64()
    pw, pn = vp(), vp()
    lib.qg_rule_device_view(r1, C.byref(ne), C.byref(n1), None, None, C.byref(pw), None)
    w_int = device_array(pw.value, n1.value, torch.float64, dev)
    lib.qg_rule_device_view(r2, C.byref(ne), C.byref(n2), None, None, C.byref(pw), C.byref(pn))
    w_bnd = device_array(pw.value, n2.value, torch.float64, dev)
    h = 1.0 / GRID
    # cnt = [full, empty, cut, records] of the slab
    part = torch.tensor([cnt[0] * h ** 3 + float(w_int.sum()) * h ** 3, float(w_bnd.sum()) * h ** 2, dt], dtype=torch.float64, device=dev)
    totals = [int(cnt[2]), int(n1.value), int(n2.value), int(cnt[0]), int(cnt[1])]
    allt, off = sharding.exchange_totals(totals, device=dev)
    if world > 1:
        red = part.clone()
        dist.all_reduce(red[:2])
        tmax = part[2:].clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        vol, area, tmax = float(red[0]), float(red[1]), float(tmax[0])
    else:
        vol, area, tmax = float(part[0]), float(part[1]), dt
    lib.qg_rule_free(r1)
    lib.qg_rule_free(r2)
    lib.qg_dcells_free(dc)
    lib.qg_domain_free(dom)
    ok = abs(vol - 0.5) < 1e-12 and int(allt[:, 3:5].sum() + allt[:, 0].sum()) == GRID ** 3
    ok = ok and np.array_equal(off, allt[:rank].sum(axis=0))
    if rank == 0:
        print(json.dumps({"config": "C5 Schwarz-diamond (2,2,2), 512^3, 6 pts/dir", "n_gpus": world, "cut_cells": int(allt[:, 0].sum()),
                          "interior_nodes": int(allt[:, 1].sum()), "boundary_nodes": int(allt[:, 2].sum()), "volume": vol,
                          "volume_error": abs(vol - 0.5), "unfitted_boundary_area": area, "wall_s_max_over_ranks": tmax,
                          "cut_cells_per_s": int(allt[:, 0].sum()) / tmax, "per_rank_cut_cells": allt[:, 0].tolist(),
                          "rank_offsets_ok": bool(ok)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
