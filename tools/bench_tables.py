"""Times the table generator at the reference's full integration resolution (quality 1) on a full-range energy
window (a Gaussian-broadened source line forces the 0.1-200 keV grid: 400 energies) for the 17 elements of srm1412:
host generator (OpenMP, provider called at every integration step) against the GPU generator."""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import xmimsim_b200 as x  # noqa: E402
from inputs import example  # noqa: E402


def main():
    inp = example("srm1412")
    inp.discrete[0].distribution_type = 1
    inp.discrete[0].scale_parameter = 0.05           # broadened line: tables must cover up to 200 keV
    out = {"workload": "17 elements x 400 energies x 2 processes x 1e5 theta steps + 17 x 1e7 pz steps", "cores": os.cpu_count()}
    t0 = time.time(); dev = x.Simulation(inp, quality=1, gpu_tables=True); out["gpu_path_wall_s"] = round(time.time() - t0, 3)
    out["gpu_kernels_ms"] = round(dev.L.xmb_tables_gpu_last_ms(), 2)
    t0 = time.time(); dev2 = x.Simulation(inp, quality=1, gpu_tables=True); out["gpu_path_wall_s_second"] = round(time.time() - t0, 3)
    assert dev.tables.n_icdf_E == 400
    if "--no-cpu" not in sys.argv:
        t0 = time.time(); host = x.Simulation(inp, quality=1); out["host_path_wall_s"] = round(time.time() - t0, 3)
        out["speedup_wall"] = round(out["host_path_wall_s"] / out["gpu_path_wall_s_second"], 1)
        host.close()
    dev.close(); dev2.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
