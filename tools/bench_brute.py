"""Times the brute-force mode (use_variance_reduction = 0) on the bench workload (srm1412) on the GPU and a bounded
sample on the CPU oracle."""
import ctypes as C
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

import numpy as np  # noqa: E402
import xmimsim_b200 as x  # noqa: E402
from inputs import example  # noqa: E402


def main():
    n_line = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
    inp = example("srm1412")
    inp.n_photons_line = n_line
    sim = x.Simulation(inp, quality=0)
    opt = x.main_options(use_variance_reduction=0)
    out = {"workload": "srm1412, %d lines x %d photons/line, brute force, M-lines + both cascades" % (len(inp.discrete), n_line)}
    for rep in range(3):
        ex = sim.main_msim_device(opt, None)
        out["gpu_kernel_ms_run%d" % rep] = round(ex.kernel_ms, 2)
    n = int(ex.n_histories)
    out["histories"] = n
    out["gpu_histories_per_s"] = n / (ex.kernel_ms * 1e-3)
    out["counters"] = sim.brute_counters()
    if "--no-cpu" not in sys.argv:
        import orc
        from helpers import DEFAULT_SEED
        ci = x.CInput(inp)
        od = orc.init_input(C.pointer(ci.input))
        cores = os.cpu_count() or 1
        n_cpu = min(n, 20000000)
        t0 = time.time()
        orc.main_msim_brute_range(C.pointer(ci.input), od, sim.L.xmb_get_tables(sim.hdf5F), opt, DEFAULT_SEED, 0, n_cpu,
                                  inp.n_interactions_trajectory, inp.nchannels, cores)
        dt = time.time() - t0
        out["cpu_oracle_histories_per_s"] = n_cpu / dt
        out["cpu_cores"] = cores
        out["speedup_vs_cpu_oracle"] = out["gpu_histories_per_s"] / out["cpu_oracle_histories_per_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
