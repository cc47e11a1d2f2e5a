"""Mimics the test sequence in front of test_synthetic_ten_layers_eight_interactions (four srm1412 runs with option
variants, then the 10-layer run), engine and oracle, many times in one process; on a mismatch prints which side moved
(against the values of the matching repetitions) and repeats with a fresh Pair."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

import numpy as np  # noqa: E402
import xmimsim_b200 as x  # noqa: E402
from helpers import Pair  # noqa: E402
from inputs import synthetic_layers, example  # noqa: E402


def both(inp, opt, grid_n=128):
    P = Pair(inp)
    sa = P.grid(hits_per_single=400, n=grid_n)
    ch, br, vr = P.sim.main_msim(opt, sa)
    ch_o, vr_o, cnt = P.oracle(opt, sa, 0)
    g = np.ctypeslib.as_array(sa.solid_angles, shape=(grid_n * grid_n,)).copy()
    P.close()
    return ch, ch_o, g, cnt


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    good = None
    for rep in range(reps):
        for opts in (dict(use_M_lines=0), dict(use_cascade_auger=0), dict(use_cascade_radiative=0), dict(use_cascade_auger=0, use_cascade_radiative=0)):
            a = example("srm1412"); a.n_photons_line = 800
            ch, ch_o, g, cnt = both(a, x.main_options(**opts))
            err = float(np.abs(ch - ch_o).max() / np.abs(ch_o).max())
            if err > 2e-6:
                print(json.dumps({"rep": rep, "input": "srm1412 %r" % opts, "err": err}), flush=True)
        ch, ch_o, g, cnt = both(synthetic_layers(n_photons=30000, n_int=8), x.main_options())
        err = float(np.abs(ch - ch_o).max() / np.abs(ch_o).max())
        if err <= 2e-6 and good is None:
            good = (ch.copy(), ch_o.copy(), g.copy())
            print(json.dumps({"rep": rep, "good": True, "gpu_rows": [float(v) for v in ch.sum(axis=1)], "grid_sum": float(g.sum())}), flush=True)
        if err > 2e-6:
            out = {"rep": rep, "input": "synthetic10", "err": err, "gpu_rows": [float(v) for v in ch.sum(axis=1)],
                   "orc_rows": [float(v) for v in ch_o.sum(axis=1)], "grid_sum": float(g.sum()), "cnt": [int(c) for c in cnt]}
            if good is not None:
                out["gpu_equals_good"] = bool(np.array_equal(ch, good[0]))
                out["orc_vs_good"] = float(np.abs(ch_o - good[1]).max() / np.abs(good[1]).max())
                out["grid_equals_good"] = bool(np.array_equal(g, good[2]))
            ch2, ch_o2, g2, cnt2 = both(synthetic_layers(n_photons=30000, n_int=8), x.main_options())
            out["fresh_pair_err"] = float(np.abs(ch2 - ch_o2).max() / np.abs(ch_o2).max())
            out["fresh_gpu_equals_bad_gpu"] = bool(np.array_equal(ch2, ch))
            out["fresh_orc_vs_bad_orc"] = float(np.abs(ch_o2 - ch_o).max() / np.abs(ch_o2).max())
            print(json.dumps(out), flush=True)
    print("done", reps)


if __name__ == "__main__":
    main()
