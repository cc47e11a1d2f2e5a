"""Kernel-resident time of the history kernel on the bench workload at a reduced size (srm1412, n photons/line):
for comparing experiment builds (XMIMSIM_B200_LIB=...).  Prints one JSON line."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402
import xmimsim_b200 as x  # noqa: E402
from inputs import example, synthetic_layers  # noqa: E402


def main():
    n_line = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    which = sys.argv[2] if len(sys.argv) > 2 else "srm1412"
    if which == "synthetic10":
        inp = synthetic_layers(n_photons=n_line, n_int=8)
    else:
        inp = example(which)
        inp.n_photons_line = n_line
        if os.environ.get("XMB_BENCH_NCH"):   # experiment: fewer channels (a smaller staging area in shared memory)
            inp.nchannels = int(os.environ["XMB_BENCH_NCH"])
    sim = x.Simulation(inp, quality=0)
    g, r, t = sim.solid_angle_calculation(hits_per_single=5000, seed=1)
    sa = sim.make_solid_angle(g.copy(), r.copy(), t.copy())
    opt = x.main_options()
    ms = []
    for _ in range(4):
        ex = sim.main_msim_device(opt, sa)
        ms.append(round(ex.kernel_ms, 2))
    print(json.dumps({"lib": os.environ.get("XMIMSIM_B200_LIB", "default"), "workload": which, "histories": int(ex.n_histories),
                      "kernel_ms": ms, "histories_per_s": ex.n_histories / (min(ms[1:]) * 1e-3)}))
    sim.close()


if __name__ == "__main__":
    main()
