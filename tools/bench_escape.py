"""Times the escape-ratio Monte Carlo (xmi_escape_ratios_calculation defaults: 1990 energies x 500000 photons in
the Si crystal of examples/srm1155.xmsi) on the GPU and a bounded sample of it on the CPU oracle."""
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
    inp = example("srm1155")
    sim = x.Simulation(inp, quality=0)
    ero = sim.escape_ratios_options()
    t0 = time.time()
    ein, eh = sim.escape_ratios_handles(ero, quality=0)
    t_tables = time.time() - t0
    out = {"workload": "srm1155 Si crystal, %d energies x %d photons" % (ero.n_input_energies, ero.n_photons),
           "tables_s": round(t_tables, 2)}
    for rep in range(2):
        t0 = time.time()
        er = sim.escape_ratios_run(ein, eh, ero, seed=0)
        wall = time.time() - t0
        ms = sim.L.xmb_escape_ratios_last_ms()
        if rep == 0:
            Z, fluo, e_in, compt, e_out = sim.escape_ratios_arrays(er)
        sim.escape_ratios_free(er)
        out["gpu_kernel_ms_run%d" % rep] = round(ms, 2)
        out["gpu_wall_s_run%d" % rep] = round(wall, 3)
    n = ero.n_input_energies * ero.n_photons
    out["gpu_photons_per_s"] = n / (out["gpu_kernel_ms_run1"] * 1e-3)
    out["si_K_escape_at_5keV_8keV_20keV"] = [float(fluo[i, :29, 0].sum()) for i in (40, 70, 190)]
    if "--no-cpu" not in sys.argv:
        import orc
        ero_s = sim.escape_ratios_options(n_input_energies=64, n_photons=100000, input_energy_delta=3.0)
        ein_s, eh_s = sim.escape_ratios_handles(ero_s, quality=0)
        cin = sim.L.xmb_input_F2C(ein_s)
        od = orc.init_input(cin)
        T = sim.L.xmb_get_tables(eh_s)
        cores = os.cpu_count() or 1
        t0 = time.time()
        orc.escape_ratios(cin, od, T, 5, 64, T.contents.nZ, 100000, 1999, 0.1, 0.1, cores)
        dt = time.time() - t0
        out["cpu_oracle_photons_per_s"] = 64 * 100000 / dt
        out["cpu_cores"] = cores
        out["cpu_sample"] = "64 energies (1..190 keV) x 100000 photons"
        out["speedup_vs_cpu_oracle"] = out["gpu_photons_per_s"] / out["cpu_oracle_photons_per_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
