#!/usr/bin/env python
"""Sweep every [M]-tagged (from-memory) choice of the oracle's Algoim restatement and score each variant against
the reference's golden point totals and centroids (cpp/test/test_impl_general_quad_0.cpp:153-306).

    python tools/calibrate_oracle.py [--quick] [--out profiles/r2_oracle_calibration.tsv]

Score = sum over the five totals of |oracle - golden| / max(golden, 1)  (2-D: 213 / 60; 3-D: 795837 / 177421 / 4626)
plus 1e5 x the largest centroid distance.  Classification counts (4/4/8, 896/896/2304) must hold or the variant is
rejected.  Test infrastructure only.
"""
import argparse, itertools, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "calib"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import score as S
import oracle_lib as O

GRID = dict(
    trig_bound=(0, 1),                 # sin/cos remainder: |f''| <= 1, or the tight sup |f''| on the range
    trig_C=(1.0, 1.25, 1.5, 1.75, 2.0, 2.5),  # multiplier on the 1/2 D^2 remainder
    rootfind_maxlevel=(2, 4, 6, 8),    # subdivision cap of the non-monotone rootFind
    rf_cap_mode=(0, 1, 2),             # at that cap: bracketed Newton anyway / push the midpoint / nothing
    rf_use_sign=(0, 1),                # interval sign test excludes segments before the monotonicity test
    psi_default_sign=(0, -1),          # sign decoded from a default-constructed PsiCode (read at the depth cap)
)
QUICK = dict(trig_bound=(0,), trig_C=(1.0, 1.5, 2.0), rootfind_maxlevel=(4,), rf_cap_mode=(0, 2), rf_use_sign=(1,),
             psi_default_sign=(0, -1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--libm", type=int, default=1)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    grid = QUICK if a.quick else GRID
    lib = O.load(bool(a.libm))
    lib.orc_set_knob.argtypes = [O.C.c_char_p, O.C.c_double]
    names = list(grid)
    rows = []
    t0 = time.time()
    for vals in itertools.product(*[grid[n] for n in names]):
        for n, v in zip(names, vals):
            assert lib.orc_set_knob(n.encode(), float(v)) == 0
        r = S.score(libm=bool(a.libm))
        ok = r["c2"] == (4, 4, 8) and r["c3"] == (896, 896, 2304)
        tot = r["t2"] + r["t3"]
        gold = S.GOLD2 + S.GOLD3
        sc = sum(abs(t - g) / max(g, 1) for t, g in zip(tot, gold))
        cd = max(np.abs(r["cen_int"] * 1e-7 - (S.GOLD3_CEN_INT - .5)).max(), np.abs(r["cen_bnd"] * 1e-6 - (S.GOLD3_CEN_BND - .5)).max())
        sc += 1e5 * cd
        rows.append((sc if ok else 1e9, vals, tot, ok, cd))
    rows.sort(key=lambda x: x[0])
    lines = ["score\t" + "\t".join(names) + "\tint2\tbnd2\tfac2\tint3\tbnd3\tfac3\tclassif_ok\tmax_centroid_dist"]
    for sc, vals, tot, ok, cd in rows:
        lines.append("%.5f\t" % sc + "\t".join(str(v) for v in vals) + "\t" + "\t".join(str(t) for t in tot) + "\t%d\t%.2e" % (ok, cd))
    txt = "\n".join(lines) + "\n"
    if a.out:
        with open(a.out, "w") as f:
            f.write("# golden: 213 60 0 795837 177421 4626 ; libm=%d ; %d variants ; %.0f s\n" % (a.libm, len(rows), time.time() - t0))
            f.write(txt)
    print("\n".join(lines[:25]))


if __name__ == "__main__":
    main()
