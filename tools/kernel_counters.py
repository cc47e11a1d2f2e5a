"""One warm + one measured launch of the history kernel on a bounded sample of a named workload, meant to run under
`ncu --metrics ...` (bench.py starts it as a sub-process OUTSIDE its timed region to measure warp instructions and DRAM
bytes per history in the same run; tools/profile_kernel.sh uses it for the `--set full` captures).  Prints one JSON line.

    python tools/kernel_counters.py srm1412 400000        # workload, photons per line (or total histories for configs3/4)
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import xmimsim_b200 as x  # noqa: E402
from xmimsim_b200 import workloads  # noqa: E402


def build(name, n):
    if name == "configs3":
        return workloads.synthetic_layers(n_photons=n, n_int=8)
    if name == "configs4":
        per = max(1, n // workloads.n_source_segments(workloads.ebel_tube(1)))
        return workloads.ebel_tube(n_photons_interval=per)
    inp = workloads.example(name)
    inp.n_photons_line = n
    return inp


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "srm1412"
    n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 400000
    launches = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    grid_n = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
    inp = build(name, n)
    sim = x.Simulation(inp, quality=0)
    if grid_n >= 1024:
        g, r, t = sim.solid_angle_calculation(hits_per_single=2000, seed=1)
        g, r, t = g.copy(), r.copy(), t.copy()
    else:
        r_full, t_full = sim.solid_angle_inputs()
        r = np.linspace(r_full[0], r_full[-1], grid_n); t = np.linspace(t_full[0], t_full[-1], grid_n)
        g, _ = sim.solid_angle_grid(r, t, hits_per_single=2000, seed=1)
    sa = sim.make_solid_angle(g, r, t)
    opt = x.main_options()
    ms = []
    for _ in range(launches):
        ex = sim.main_msim_device(opt, sa)
        ms.append(ex.kernel_ms)
    print(json.dumps({"workload": name, "histories": int(ex.n_histories), "interactions": int(ex.n_interactions),
                      "kernel_ms": ms, "launches": launches}))
    sim.close()


if __name__ == "__main__":
    main()
