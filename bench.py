#!/usr/bin/env python
"""bench.py -- cut cells/s and quadrature points/s on the BASELINE.json workloads.

Default workload (--config C3, the configuration BASELINE.json's metric is quoted on): 3-D Schoen gyroid, periods
(2,2,2), 256^3 cells on [0,1]^3, 8 Gauss points per direction.  The other BASELINE configs are selectable with
--config {C1,C2,C3,C3b,C4,C5}; their lines are committed under profiles/ and tabulated in BASELINE.md.
A step = one pass of the reference's hot path over the grid:
    create_unfitted_impl_domain      (kd-tree sign pruning + cell / facet classification)
  + create_quadrature(cut cells, n)  (interior rule)
  + create_unfitted_bound_quadrature(cut cells, n, False, False)   (unfitted-boundary rule + normals)
value  : cut cells / s with everything device-resident (inputs are the grid breaks and the function
         parameters; the rules stay in HBM), timed with CUDA events on the engine's stream.
e2e    : the same step through the reference-facing API (qugar_b200.cpp -> C ABI) with HOST buffers:
         breaks/cell ids go up, cell statuses, facet records and both rules come back to pinned host memory.
Multi-GPU: --scaling weak (default; what the driver's 1/2/4/8 run measures): the grid grows to n x n x (n N) cells on
[0,1]^2 x [0,N] and rank r classifies and integrates its own n^3 z-slab.  --scaling strong: the NAMED grid is split into
N z-slabs of (nearly) equal cut-cell count.  Either way a rank owns a contiguous flat-id range; the only collective is
one all-gather of 4 int64 per rank (cut cells, interior points, boundary points, device time) for the global offsets.
--impl reference: the reference's CPU path (oracle port; the real QUGaR needs Algoim, absent offline) on all host
cores -- one PROCESS per core, each classifying and integrating its own sub-slab -- over a
bounded z-slab sample of the same grid.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# oracle / engine function kinds (include/qugar_b200.h; tests/test_abi.py checks that both enumerations agree)
K_SPHERE, K_BEZIER, K_SCHOEN, K_SCHWARZ_D = 0, 2, 10, 14
_BEZIER_C4 = np.concatenate([[4.0, 4.0, 4.0], 2.0 * np.random.default_rng(20260101).random(64) - 1.0])

# BASELINE.json configs.  `make`: the level set through the reference-facing factories.
CONFIGS = {
    "C1": dict(dim=2, n=64, npts=4, kind=K_SPHERE, params=[0.4, 0.5, 0.5],
               make=lambda cpp: cpp.create_disk(0.4, np.array([0.5, 0.5]), False),
               text="2D disk r=0.4 c=(.5,.5) (general path), 64^2 cells on [0,1]^2, 4 pts/dir"),
    "C2": dict(dim=3, n=128, npts=5, kind=K_SPHERE, params=[0.4, 0.5, 0.5, 0.5],
               make=lambda cpp: cpp.create_sphere(0.4, np.array([0.5, 0.5, 0.5]), False),
               text="3D sphere r=0.4 c=(.5,.5,.5) (general path), 128^3 cells on [0,1]^3, 5 pts/dir"),
    "C3": dict(dim=3, n=256, npts=8, kind=K_SCHOEN, params=[2.0, 2.0, 2.0], make=lambda cpp: cpp.create_Schoen([2.0, 2.0, 2.0]),
               text="3D Schoen gyroid periods (2,2,2), 256^3 cells on [0,1]^3, 8 pts/dir"),
    "C3b": dict(dim=3, n=256, npts=8, kind=K_SCHOEN, params=[1.9, 2.1, 2.3], make=lambda cpp: cpp.create_Schoen([1.9, 2.1, 2.3]),
                text="3D Schoen gyroid periods (1.9,2.1,2.3) (not grid aligned), 256^3 cells on [0,1]^3, 8 pts/dir"),
    "C4": dict(dim=3, n=256, npts=8, kind=K_BEZIER, params=_BEZIER_C4,
               make=lambda cpp: cpp.create_bezier([4, 4, 4], _BEZIER_C4[3:]),
               text="3D degree-3 BezierTP (64 coefficients 2U(0,1)-1, default_rng(20260101)), 256^3 cells, 8 pts/dir; integrated "
                    "with the GENERAL algorithm (the reference uses ImplicitPolyQuadrature: node sets differ, DESIGN.md 6)"),
    "C5": dict(dim=3, n=512, npts=6, kind=K_SCHWARZ_D, params=[2.0, 2.0, 2.0], make=lambda cpp: cpp.create_SchwarzDiamond([2.0, 2.0, 2.0]),
               text="3D Schwarz diamond periods (2,2,2), 512^3 cells on [0,1]^3, 6 pts/dir"),
}
STEP_TEXT = ": domain classification + interior rule + unfitted-boundary rule"


def _tracked_json(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return None


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every few milliseconds from a thread (the timed
    region is a few hundred milliseconds; nvidia-smi would return one or two samples), nvidia-smi as the fallback."""

    def __init__(self, gpu_index):
        self.rows = []          # (sm_mhz, sm_max_mhz, set of reason names)
        self.stop = False
        self.gpu = gpu_index
        self.th = None
        self.source = "nvml"

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = self.gpu
        if vis:
            ent = vis.split(",")[self.gpu].strip()
            if ent.isdigit():
                idx = int(ent)
            else:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(ent.encode() if hasattr(ent, "encode") else ent)
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def _run_nvml(self):
        nv, h = self._nvml_handle()
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown",
                                                getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown",
                                                getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4))}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        while not self.stop:
            sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            mask = int(get_reasons(h))
            self.rows.append((sm, mx, {n for n, bit in names.items() if mask & bit}))
            time.sleep(0.004)

    def _run_smi(self):
        self.source = "nvidia-smi"
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                c = [v.strip() for v in out.split(",")]
                if len(c) >= 6:
                    self.rows.append((float(c[0]), float(c[1]), {n for i, n in enumerate(names) if c[2 + i].lower().startswith("active")}))
            except Exception:
                pass
            time.sleep(0.05)

    def _run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        time.sleep(0.02)      # let the first sample land before the timed region starts
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted(set().union(*[r[2] for r in self.rows]))
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(r[1] for r in self.rows), "reasons": reasons,
                "samples": len(self.rows), "source": self.source}


def _cpu_worker(job):
    """One PROCESS of the CPU arm: classifies and integrates its own sub-slab with the single-threaded oracle."""
    dim, kind, params, breaks, npts = job
    import oracle_lib as O
    lib = O.load(False)
    lib.orc_time_rules.restype = C.c_longlong
    dom = O.OracleDomain(dim, kind, params, breaks, nthreads=1)
    cut = dom.cut_cells
    nbp = C.c_longlong()
    ni = lib.orc_time_rules(C.c_void_p(dom.h), cut.ctypes.data_as(C.c_void_p), C.c_int64(len(cut)), npts, 1, 1, 1, C.byref(nbp))
    return len(cut), int(ni) + int(nbp.value)


_POOL = None


def _pool(cores):
    global _POOL
    if _POOL is None:
        _POOL = mp.get_context("spawn").Pool(cores)
        _POOL.map(_cpu_warm, range(cores))      # start the interpreters and load the oracle outside any timed region
    return _POOL


def _cpu_warm(_):
    import oracle_lib as O
    O.load(False)
    return 0


def _cpu_step(cfg, slab, cores):
    """One step of the CPU path (classify + interior rule + boundary rule) on a slab of `slab` cell layers of the last
    direction taken from the middle of the grid (same cell sizes and positions as the GPU arm's cells), split into
    contiguous sub-slabs over `cores` processes.  Returns (seconds, cut cells, nodes)."""
    n, dim = cfg["n"], cfg["dim"]
    full = np.linspace(0.0, 1.0, n + 1)
    k0 = n // 2 - slab // 2
    nproc = max(1, min(cores, slab))
    edges = [k0 + slab * i // nproc for i in range(nproc + 1)]
    jobs = [(dim, cfg["kind"], list(map(float, cfg["params"])), [full] * (dim - 1) + [full[a:b + 1]], cfg["npts"])
            for a, b in zip(edges[:-1], edges[1:])]
    pool = _pool(cores)
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker, jobs, chunksize=1)
    dt = time.perf_counter() - t0
    return dt, sum(r[0] for r in res), sum(r[1] for r in res)


def _cpu_sample(cfg, args, target_s):
    """Slab thickness whose CPU step takes about target_s seconds (calibrated on a thin slab), capped at the grid."""
    cores = os.cpu_count() or 1
    n = cfg["n"]
    if args.ref_slab > 0:
        return cores, min(n, args.ref_slab)
    probe = min(n, max(4, cores // 4))
    dt, _, _ = _cpu_step(cfg, probe, cores)
    dt, _, _ = _cpu_step(cfg, probe, cores)
    slab = int(max(probe, min(n, probe * target_s / max(dt, 1e-3))))
    return cores, (slab - slab % 4 if slab < n else n)


def run_reference(args):
    """CPU arm: the reference's algorithm (oracle port; the real QUGaR needs Algoim, absent offline) on all host
    cores, one process per core; every step is a bounded z-slab sample of the same grid (about 10 s of CPU work)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    # every step a bounded sample: ~10 s of CPU work, less when many steps are requested (whole run within ~2 minutes)
    cores, slab = _cpu_sample(cfg, args, min(10.0, max(1.0, 120.0 / max(1, args.steps + args.warmup))))
    times = []
    ncut = npts = 0
    for it in range(args.warmup + args.steps):
        dt, ncut, npts = _cpu_step(cfg, slab, cores)
        if it >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    val = ncut / t
    n = cfg["n"]
    k0 = n // 2 - slab // 2
    shape = "x".join([str(n)] * (cfg["dim"] - 1) + [str(slab)])
    line = {
        "impl": "reference", "metric": "cut_cells_per_s", "value": val, "unit": "cut cells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["text"] + STEP_TEXT, "name": args.config,
                   "sample": f"slab of {slab} cell layers of the last direction (layers {k0}..{k0 + slab - 1} of {n})"},
        "cpu_baseline": {"value": val, "unit": "cut cells/s", "cores": cores, "kind": "port",
                         "sample": f"{ncut} cut cells of a {shape} slab per step, {t:.1f} s per step on {cores} processes "
                                   "(one sub-slab each); oracle port (Algoim unavailable offline)"},
        "points_per_s": npts / t,
        "e2e": {"value": val, "unit": "cut cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


class Engine:
    """Thin ctypes access to the device-resident entry points of libqugar_b200.so."""

    def __init__(self):
        from qugar_b200 import cpp
        self.cpp = cpp
        self.lib = cpp._load()
        vp = C.c_void_p
        L = self.lib
        L.qg_domain_create_slab.argtypes = [vp, vp, C.c_int, C.c_int64, C.c_int64, C.POINTER(vp)]
        L.qg_domain_cut_cells_device.argtypes = [vp, C.POINTER(vp)]
        L.qg_dcells_count.argtypes = [vp, C.POINTER(C.c_int64)]
        L.qg_timer_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.qg_timer_start.argtypes = [vp]
        L.qg_timer_stop.argtypes = [vp, C.POINTER(C.c_double)]
        L.qg_launch_count.restype = C.c_longlong
        L.qg_rule_d2h_bytes.restype = C.c_longlong
        L.qg_rule_d2h_bytes.argtypes = [vp]

    def device_step(self, func, grid, device, k0, k1, NPTS):
        L, cpp = self.lib, self.cpp
        dom, dc, r1, r2 = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        cpp._check(L.qg_domain_create_slab(func._h, grid._h, device, k0, k1, C.byref(dom)))
        cpp._check(L.qg_domain_cut_cells_device(dom, C.byref(dc)))
        cpp._check(L.qg_quad_cells_device(dom, dc, NPTS, C.byref(r1)))
        cpp._check(L.qg_quad_unf_bdry_device(dom, dc, NPTS, C.byref(r2)))
        ncut = C.c_int64()
        L.qg_dcells_count(dc, C.byref(ncut))
        ne, n1, n2 = C.c_int64(), C.c_int64(), C.c_int64()
        L.qg_rule_device_view(r1, C.byref(ne), C.byref(n1), None, None, None, None)
        L.qg_rule_device_view(r2, C.byref(ne), C.byref(n2), None, None, None, None)
        ms1, ms2 = (C.c_double * 8)(), (C.c_double * 8)()
        L.qg_rule_pass_ms(r1, ms1)
        L.qg_rule_pass_ms(r2, ms2)
        cn = (C.c_longlong * 4)()
        L.qg_rule_pass_counts(r1, cn)
        self.lines_interior = int(cn[2])
        L.qg_rule_pass_counts(r2, cn)
        self.lines_boundary = int(cn[2])
        L.qg_rule_free(r1)
        L.qg_rule_free(r2)
        L.qg_dcells_free(dc)
        L.qg_domain_free(dom)
        return ncut.value, n1.value, n2.value, list(ms1), list(ms2)


def _bind_to_gpu_numa_node(local):
    """Pin this rank (and therefore its pinned host buffers, placed by first touch) to the NUMA node its GPU hangs off:
    with 8 ranks each bringing 25 GB per step to the host, cross-socket traffic would otherwise cap the e2e figure."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
        if bus is None:
            out = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                 capture_output=True, text=True, timeout=10).stdout.strip()
            bus = out[-12:] if out else None      # 00000000:1B:00.0 -> 0000:1b:00.0
        if not bus:
            return None
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def _private_stdout():
    """Keeps stdout for the one JSON line: libraries that write to file descriptor 1 on their own (NCCL prints its version
    there on the first communicator) are sent to stderr instead."""
    sys.stdout.flush()
    keep = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return keep


def _slabs_strong(eng, cpp, func, grid, device, n_last, world):
    """Strong scaling: layers [k0,k1) per rank with (nearly) equal cut-cell count (qugar_b200/sharding.py).  Every rank
    classifies the whole grid once, outside the timed region, to get the cut cells per layer."""
    from qugar_b200 import sharding
    if world == 1:
        return [(0, n_last)]
    dom = cpp.create_unfitted_impl_domain(func, grid, device=device)
    cut = dom.get_cut_cells()
    per_layer = int(np.prod(grid.num_cells_dir.astype(np.int64)[:-1]))
    counts = np.bincount(cut // per_layer, minlength=n_last)
    del dom
    return sharding.balanced_slabs(counts, world)


def run_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    json_out = _private_stdout()
    import torch
    numa = _bind_to_gpu_numa_node(local) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine()
    cpp = eng.cpp
    if cpp.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; the CUDA extension is the only implementation of this path")
    device = local % cpp.device_count()
    cfg = CONFIGS[args.config]
    GRID, NPTS, dim = cfg["n"], cfg["npts"], cfg["dim"]
    strong = args.scaling == "strong"
    full = np.linspace(0.0, 1.0, GRID + 1)
    func = cfg["make"](cpp)
    if strong:
        # the NAMED grid, split into `world` slabs of the last direction with equal cut-cell counts
        breaks = [full] * dim
        grid = cpp.create_cart_grid(breaks)
        k0, k1 = _slabs_strong(eng, cpp, func, grid, device, GRID, world)[rank]
        global_grid = "x".join([str(GRID)] * dim) + " cells on [0,1]^%d" % dim
    else:
        # Weak scaling: rank r owns the r-th n^dim block of an n x .. x (n N) grid on [0,1]^(dim-1) x [0,N] (same cell
        # size, same periods per unit length), i.e. the contiguous flat-id range of its slab of the last direction.
        zfull = np.linspace(0.0, float(world), GRID * world + 1)
        breaks = [full] * (dim - 1) + [zfull]
        grid = cpp.create_cart_grid(breaks)
        k0, k1 = GRID * rank, GRID * (rank + 1)
        global_grid = "x".join([str(GRID)] * (dim - 1) + [str(GRID * world)]) + " cells on [0,1]^%d x [0,%d]" % (dim - 1, world)

    timer = C.c_void_p()
    cpp._check(eng.lib.qg_timer_create(device, C.byref(timer)))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps
    for _ in range(args.warmup):
        eng.device_step(func, grid, device, k0, k1, NPTS)
    barrier()
    pass_acc = np.zeros((2, 8))
    with ClockSampler(device) as clocks:
        launches0 = eng.lib.qg_launch_count()
        cpp._check(eng.lib.qg_timer_start(timer))
        t_wall0 = time.perf_counter()
        for _ in range(args.steps):
            ncut, npts_int, npts_bnd, ms1, ms2 = eng.device_step(func, grid, device, k0, k1, NPTS)
            pass_acc[0] += ms1
            pass_acc[1] += ms2
        ms = C.c_double()
        cpp._check(eng.lib.qg_timer_stop(timer, C.byref(ms)))
        wall = time.perf_counter() - t_wall0
        launches = eng.lib.qg_launch_count() - launches0
    barrier()
    dev_ms = ms.value / args.steps
    pass_acc /= args.steps
    clk = clocks.summary()

    # the one collective of the sharded path: per-rank totals -> global offsets
    totals = torch.tensor([ncut, npts_int, npts_bnd, int(dev_ms * 1e6)], dtype=torch.int64, device=f"cuda:{local}")
    if dist is not None:
        gathered = [torch.zeros_like(totals) for _ in range(world)]
        dist.all_gather(gathered, totals)
        allv = torch.stack(gathered).cpu().numpy()
    else:
        allv = totals.cpu().numpy()[None, :]
    ncut_all, nint_all, nbnd_all = int(allv[:, 0].sum()), int(allv[:, 1].sum()), int(allv[:, 2].sum())
    max_ms = float(allv[:, 3].max()) / 1e6

    # ---- end to end through the reference-facing API, host buffers (rank-local slab)
    e2e_ms = []
    h2d = d2h = 0
    for it in range(args.e2e_steps + 1 if args.e2e_steps > 0 else 0):
        barrier()
        t0 = time.perf_counter()
        g2 = cpp.create_cart_grid(breaks)
        f2 = cfg["make"](cpp)
        hdom = C.c_void_p()
        cpp._check(eng.lib.qg_domain_create_slab(f2._h, g2._h, device, k0, k1, C.byref(hdom)))
        cnt = (C.c_int64 * 4)()
        eng.lib.qg_domain_counts(hdom, cnt)
        cut = np.empty(cnt[2], np.int64)
        cpp._check(eng.lib.qg_domain_cells(hdom, 0, cut.ctypes.data_as(C.c_void_p)))   # status + records -> host
        r1, r2 = C.c_void_p(), C.c_void_p()
        cpp._check(eng.lib.qg_quad_cells(hdom, cut.ctypes.data_as(C.c_void_p), len(cut), NPTS, C.byref(r1)))
        cpp._check(eng.lib.qg_quad_unf_bdry(hdom, cut.ctypes.data_as(C.c_void_p), len(cut), NPTS, 0, 0, C.byref(r2)))
        # touch the host results like a caller would (first and last point / weight of the interior rule)
        ne, np_, pd = C.c_int64(), C.c_int64(), C.c_int()
        pw, pp = C.c_void_p(), C.c_void_p()
        eng.lib.qg_rule_view(r1, C.byref(ne), C.byref(np_), C.byref(pd), None, C.byref(pp), C.byref(pw), None)
        w = (C.c_double * np_.value).from_address(pw.value)
        x = (C.c_double * (dim * np_.value)).from_address(pp.value)
        chk = w[0] + w[np_.value - 1] + x[0] + x[dim * np_.value - 1] if np_.value else 0.0
        dt = time.perf_counter() - t0
        h2d = sum(len(b) for b in breaks) * 8 + 2 * len(cut) * 8
        # cut-cell ids + what the two rule downloads actually moved (the 3-D interior rule travels encoded)
        d2h = len(cut) * 8 + eng.lib.qg_rule_d2h_bytes(r1) + eng.lib.qg_rule_d2h_bytes(r2)
        eng.lib.qg_rule_free(r1)
        eng.lib.qg_rule_free(r2)
        eng.lib.qg_domain_free(hdom)
        if it > 0:
            e2e_ms.append(dt * 1e3)
        assert chk == chk
    # ---- end to end with the device-side consumer: the rules stay in HBM, the custom-coefficient array
    # of the reference's FEniCSx layer (custom_coefficients.py:278-512) is packed on the GPU and ONLY it crosses PCIe.
    # Integral domain = the slab's full + cut cells, no coefficient functions (n_old_cols = 0), float64 and float32 forms.
    consumer = {}
    if args.e2e_steps > 0 and world == 1:
        from qugar_b200 import coeffs as qcoeffs
        for dt_name, dt in (("f64", np.float64), ("f32", np.float32)):
            eng.lib.qg_pinned_release(None)      # keep the pinned footprint to one phase's buffers (outside every timed region)
            tms = []
            nbytes = 0
            for it in range(min(args.e2e_steps, 3) + 1):
                barrier()
                t0 = time.perf_counter()
                g2 = cpp.create_cart_grid(breaks)
                dom = cpp.create_unfitted_impl_domain(cfg["make"](cpp), g2, device=device, slab=(k0, k1))
                cut = dom.get_cut_cells()
                fullc = dom.get_full_cells()
                ents = np.sort(np.concatenate([cut, fullc]))
                custom_ids = np.searchsorted(ents, cut)
                r1 = cpp.create_quadrature_device(dom, cut, NPTS)
                r2 = cpp.create_unfitted_bound_quadrature_device(dom, cut, NPTS)
                packed = qcoeffs.pack_custom_coeffs(np.zeros((len(ents), 0), dt), custom_ids, np.zeros(0, np.int64), [r1, r2],
                                                    to_host=True, device=device)
                chk = float(packed.host[-1, 0]) + float(packed.host[len(ents), 0])
                dt_s = time.perf_counter() - t0
                nbytes = packed.d2h_bytes + len(cut) * 8 + len(fullc) * 8
                del packed, r1, r2, dom
                if it > 0:
                    tms.append(dt_s * 1e3)
                assert chk == chk
            consumer[dt_name] = {"ms_per_step": float(np.mean(tms)), "value": ncut / (float(np.mean(tms)) * 1e-3), "unit": "cut cells/s",
                                 "d2h_bytes_per_step": int(nbytes), "samples": len(tms)}
    eng.lib.qg_pinned_release(None)
    e2e_local = float(np.mean(e2e_ms)) if e2e_ms else float("nan")
    e2e_t = torch.tensor([e2e_local], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_max_ms = float(e2e_t.item())

    if rank == 0:
        hbm, which = _peaks()
        names = ["ctl", "k0", "k1", "k2", "k3", "scans", "other", "total"]
        step_ms = dev_ms
        # ---- rooflines.  The DOMINANT kernel of the step is picked from the step's own events (pass_ms): the root-finding
        # pass K2 (fp64-issue bound) or the node writer K3 (HBM-write bound), interior + boundary rule together.
        k2_ms, k3_ms = pass_acc[0][3] + pass_acc[1][3], pass_acc[0][4] + pass_acc[1][4]
        # K3 interior: 8*(dim+1) B per node, boundary: 8*(2*dim+1) B per node, both launches
        k3_bytes = npts_int * 8.0 * (dim + 1) + npts_bnd * 8.0 * (2 * dim + 1)
        km = _tracked_json("r2_kernel_metrics.json") or {}
        k3m = km.get("k3_cell3", {})
        roof_k3 = {"bound": "hbm", "kernel": "k3 node writers (k3_cell3_kernel interior + k3_kernel boundary)",
                   "achieved": k3_bytes / (k3_ms * 1e-3) / 1e9 if k3_ms > 0 else None, "peak": hbm, "peak_source": which, "unit": "GB/s",
                   "frac": k3_bytes / (k3_ms * 1e-3) / 1e9 / hbm if k3_ms > 0 else None,
                   "traffic": (k3m["dram_bytes_per_node"] * npts_int) if "dram_bytes_per_node" in k3m else None,
                   "traffic_source": k3m.get("source"), "launch_ms": k3_ms, "share_of_step": k3_ms / step_ms}
        peaks = _tracked_json("peaks.json") or {}
        fp64_peak = float(peaks.get("fp64_instr_per_s", 17.39e12))
        k2m = km.get("k2", {}).get(str(cfg["kind"]))
        lines = eng.lines_interior + eng.lines_boundary
        if k2m:
            rate = (eng.lines_interior * k2m["fp64_instr_per_line_interior"] + eng.lines_boundary * k2m["fp64_instr_per_line_boundary"]) / (k2_ms * 1e-3)
            roof_k2 = {"bound": "fp64", "kernel": "k2_kernel (root finding along the height direction, interior + boundary rule)",
                       "achieved": rate / 1e12, "peak": fp64_peak / 1e12, "unit": "T fp64 instr/s", "frac": rate / fp64_peak,
                       "peak_source": peaks.get("source", "tools/ubench/peaks.cu (profiles/r1_ubench_peaks.txt)"),
                       "traffic": k2m.get("dram_bytes_per_line", 0.0) * lines or None, "launch_ms": k2_ms, "lines": lines,
                       "fp64_instr_per_line": [k2m["fp64_instr_per_line_interior"], k2m["fp64_instr_per_line_boundary"]],
                       "instr_source": k2m.get("source"), "share_of_step": k2_ms / step_ms}
        else:
            roof_k2 = {"bound": "fp64", "kernel": "k2_kernel", "achieved": None, "peak": fp64_peak / 1e12, "unit": "T fp64 instr/s",
                       "frac": None, "launch_ms": k2_ms, "lines": lines, "share_of_step": k2_ms / step_ms,
                       "note": "no tracked ncu instruction count for this level-set kind (profiles/r2_kernel_metrics.json)"}
        dominant_is_k2 = k2_ms >= k3_ms and roof_k2["frac"] is not None
        roof = dict(roof_k2 if dominant_is_k2 else roof_k3)
        roof["dominant_by"] = "pass_ms of this run: k2 %.2f ms, k3 %.2f ms of a %.2f ms step" % (k2_ms, k3_ms, step_ms)
        # whole step against the HBM floor (every byte the step must write: statuses, ids, both rules)
        step_bytes = k3_bytes + ncut * 20.0 + float(np.prod(grid.num_cells_dir.astype(np.int64))) / world
        line = {
            "metric": "cut_cells_per_s", "value": ncut_all / (max_ms * 1e-3), "unit": "cut cells/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": max_ms, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["text"] + STEP_TEXT, "name": args.config, "global_grid": global_grid, "cut_cells": ncut_all,
                       "interior_points": nint_all, "boundary_points": nbnd_all,
                       "parallelism": f"one slab of the last direction per GPU x{world} ({args.scaling} scaling), no data-path collective "
                                      "(one all-gather of 4 int64 per rank for global offsets)",
                       "l2": "outputs stream through HBM and exceed the 126 MB L2 (C2-C5); no reuse between steps"},
            "points_per_s": (nint_all + nbnd_all) / (max_ms * 1e-3),
            "rates": {"classification_cut_cells_per_s": ncut / ((dev_ms - pass_acc[0][7] - pass_acc[1][7]) * 1e-3),
                      "interior_rule_cut_cells_per_s": ncut / (pass_acc[0][7] * 1e-3), "interior_points_per_s": npts_int / (pass_acc[0][7] * 1e-3),
                      "boundary_rule_cut_cells_per_s": ncut / (pass_acc[1][7] * 1e-3), "boundary_points_per_s": npts_bnd / (pass_acc[1][7] * 1e-3),
                      "note": "rank 0, device-resident"},
            "gpu_launches": int(launches),
            "pass_ms": {"interior": dict(zip(names, [round(v, 3) for v in pass_acc[0]])),
                        "boundary": dict(zip(names, [round(v, 3) for v in pass_acc[1]])),
                        "domain": round(dev_ms - pass_acc[0][7] - pass_acc[1][7], 3),
                        "host_wall_ms_per_step": wall * 1e3 / args.steps, "rank0_device_ms_per_step": dev_ms},
            "e2e": {"value": ncut_all / (e2e_max_ms * 1e-3) if e2e_ms else None, "unit": "cut cells/s",
                    "ms_per_step": e2e_max_ms if e2e_ms else None, "samples": len(e2e_ms),
                    "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world,
                    "numa_bound": numa is not None},
            "e2e_device_consumer": consumer or None,
            "roofline": roof,
            "roofline_k2_fp64": roof_k2,
            "roofline_k3_hbm": roof_k3,
            "step_hbm_frac": step_bytes / (dev_ms * 1e-3) / 1e9 / hbm,
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line), file=json_out, flush=True)
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline(args):
    cfg = CONFIGS[args.config]
    cores, slab = _cpu_sample(cfg, args, 15.0)
    dt, ncut, _ = _cpu_step(cfg, slab, cores)
    shape = "x".join([str(cfg["n"])] * (cfg["dim"] - 1) + [str(slab)])
    return {"value": ncut / dt, "unit": "cut cells/s", "cores": cores, "kind": "port",
            "sample": f"{ncut} cut cells of a {shape} slab (same step: classify + interior + boundary), "
                      f"{dt:.1f} s on {cores} processes; oracle port (Algoim unavailable offline)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=5, help="timed end-to-end samples (one more is run first, untimed)")
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--ref-slab", type=int, default=0, help="z-layers of the CPU sample (0 = calibrate to ~10-15 s)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
