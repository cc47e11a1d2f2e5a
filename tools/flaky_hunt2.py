"""Repeats test_history_gpu.py::run_both (engine and oracle on the same input) in one process and reports mismatches."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

import numpy as np  # noqa: E402
import xmimsim_b200 as x  # noqa: E402
from helpers import Pair  # noqa: E402
from inputs import synthetic_layers, example, caso4  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    makers = [("synthetic10", lambda: synthetic_layers(n_photons=30000, n_int=8), 128),
              ("srm1155", lambda: _ex("srm1155"), None), ("caso4", lambda: _c4(), 128)]
    ref = {}
    for rep in range(reps):
        for name, mk, gn in makers:
            inp = mk()
            P = Pair(inp)
            sa = P.grid(hits_per_single=400, n=gn)
            ch, br, vr = P.sim.main_msim(x.main_options(), sa)
            ch_o, vr_o, cnt = P.oracle(x.main_options(), sa, 0)
            P.close()
            err = float(np.abs(ch - ch_o).max() / np.abs(ch_o).max())
            if name not in ref:
                ref[name] = (ch.copy(), ch_o.copy())
                print(json.dumps({"name": name, "first_err": err, "gpu_max": float(ch.max()), "orc_max": float(ch_o.max()), "cnt": [int(c) for c in cnt]}), flush=True)
            if err > 2e-6 or not np.array_equal(ch, ref[name][0]):
                i = np.unravel_index(np.argmax(np.abs(ch - ch_o)), ch.shape)
                print(json.dumps({"name": name, "rep": rep, "err": err, "gpu_max": float(ch.max()), "orc_max": float(ch_o.max()),
                                  "at": [int(k) for k in i], "gpu_at": float(ch[i]), "orc_at": float(ch_o[i]),
                                  "gpu_equals_first": bool(np.array_equal(ch, ref[name][0])),
                                  "orc_vs_first": float(np.abs(ch_o - ref[name][1]).max() / np.abs(ref[name][1]).max()),
                                  "cnt": [int(c) for c in cnt]}), flush=True)
    print("done", reps)


def _ex(n):
    a = example(n); a.n_photons_line = 1500
    return a


def _c4():
    a = caso4(); a.n_photons_line = 20000
    return a


if __name__ == "__main__":
    main()
