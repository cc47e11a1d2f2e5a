#!/usr/bin/env python3
"""Engine vs oracle on many random inputs (tests/random_inputs.py).  usage: tools/random_parity_hunt.py FIRST LAST [n_photons]
Prints one line per seed (worst relative difference of the channel rows and of the per-line histories, in units of the
array maximum) and a summary; exit code 1 when a seed exceeds the parity tolerance of the tests (2e-6)."""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

import numpy as np  # noqa: E402
import xmimsim_b200 as x  # noqa: E402
from helpers import Pair  # noqa: E402
from random_inputs import random_input  # noqa: E402

RTOL = 2e-6


def main():
    first, last = int(sys.argv[1]), int(sys.argv[2])
    n_photons = int(sys.argv[3]) if len(sys.argv) > 3 else 1500
    worst, bad = 0.0, []
    for seed in range(first, last):
        t0 = time.time()
        inp, opts = random_input(seed, n_photons=n_photons)
        o = x.main_options(**opts)
        P = Pair(inp)
        sa = P.grid(hits_per_single=300, n=96)
        ch, br, vr = P.sim.main_msim(o, sa)
        ch_o, vr_o, cnt = P.oracle(o, sa, 0)
        P.close()
        e_ch = float(np.abs(ch - ch_o).max() / max(np.abs(ch_o).max(), 1e-300))
        e_vr = float(np.abs(vr - vr_o).max() / max(np.abs(vr_o).max(), 1e-300))
        ok = np.isfinite(ch).all() and e_ch <= RTOL and e_vr <= RTOL
        worst = max(worst, e_ch, e_vr)
        if not ok:
            bad.append(seed)
        print("seed %4d  layers %d  n_int %d  lines %d  cont %2d  histories %6d  err channels %.2e  history %.2e  %s  %.1fs"
              % (seed, len(inp.layers), inp.n_interactions_trajectory, len(inp.discrete), len(inp.continuous), P.n_total,
                 e_ch, e_vr, "ok" if ok else "MISMATCH", time.time() - t0), flush=True)
    print("seeds %d..%d: worst relative difference %.2e, mismatches: %s" % (first, last - 1, worst, bad or "none"))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
