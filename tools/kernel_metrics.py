"""Distils an `ncu --set full` report of tools/prof_step.py into profiles/r2_kernel_metrics.json, the tracked file
bench.py derives its roofline inputs from (fp64 instructions per stage-2 line, DRAM bytes per node / line).

    python tools/kernel_metrics.py report.ncu-rep lines_interior lines_boundary nodes_interior kind > profiles/r2_kernel_metrics.json

The report must hold the launches of ONE step, in order: k2 (n=1 cells), k2 (n=1 facets), k2 interior, k2 boundary and
the k3 launches (regex "k[23]_"); the interior / boundary launches are the two largest k2 launches and the k3_cell3 one."""
import csv
import json
import subprocess
import sys

rep, lines_int, lines_bnd, nodes_int, kind = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
recs = [dict(zip(hdr, r)) for r in rows[2:]]


def num(d, k):
    return float(d[k].replace(",", "")) if d.get(k) not in (None, "") else 0.0


FP64_OPS = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")


def fp64(d):
    """fp64 issue slots of one launch, as thread slots (warp instructions x 32: an fp64 warp instruction occupies the pipe
    whatever its active mask).  The opcode mix comes from the source page (SASS patching counters, which this ncu reports
    summed over two passes), scaled to the hardware counter smsp__inst_executed.sum."""
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", d["ID"],
                          "--launch-count", "1"], capture_output=True, text=True, errors="replace").stdout
    rs = list(csv.reader(src.splitlines()))
    hi = next(i for i, r in enumerate(rs) if r and r[0] == "Address")
    h = rs[hi]
    iI, iS = h.index("Instructions Executed"), h.index("Source")
    tot = f64 = 0
    for r in rs[hi + 1:]:
        if len(r) <= max(iI, iS) or not r[0].startswith("0x"):
            continue
        t = r[iS].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        n = int(r[iI])
        tot += n
        if op in FP64_OPS:
            f64 += n
    return 32.0 * num(d, "smsp__inst_executed.sum") * f64 / max(tot, 1)


def dram(d):
    # ncu reports these in the unit of the report (bytes when exported raw)
    return num(d, "dram__bytes_read.sum") + num(d, "dram__bytes_write.sum")


k2 = sorted([d for d in recs if "k2_kernel" in d["Kernel Name"]], key=lambda d: -num(d, "gpu__time_duration.sum"))[:2]
k2_int, k2_bnd = (k2 if lines_int >= lines_bnd else k2[::-1])
k3c = [d for d in recs if "k3_cell3" in d["Kernel Name"]][0]
units = dict(zip(hdr, rows[1]))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units["dram__bytes_read.sum"]]
res = {
    "source": f"ncu --set full --clock-control none of tools/prof_step.py (one step), {rep.split('/')[-1]}; distilled by tools/kernel_metrics.py",
    "k2": {kind: {
        "fp64_instr_per_line_interior": fp64(k2_int) / lines_int,
        "fp64_instr_per_line_boundary": fp64(k2_bnd) / lines_bnd,
        "dram_bytes_per_line": scale * (dram(k2_int) + dram(k2_bnd)) / (lines_int + lines_bnd),
        "source": "fp64 warp instructions (DFMA, DMUL, DADD, DSETP: share of the source page's executed instructions x "
                  "smsp__inst_executed.sum) x 32 lanes of the interior / boundary k2_kernel launch, divided by its stage-2 lines",
        "launch_ms": [num(k2_int, "gpu__time_duration.sum"), num(k2_bnd, "gpu__time_duration.sum")],
        "fp64_pipe_pct": [num(k2_int, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                          num(k2_bnd, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")],
    }},
    "k3_cell3": {"dram_bytes_per_node": scale * dram(k3c) / nodes_int,
                 "source": "dram__bytes_read.sum + dram__bytes_write.sum of k3_cell3_kernel divided by the interior nodes of the step"},
}
print(json.dumps(res, indent=1))
