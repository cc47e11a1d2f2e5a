"""CPU-only: runs the ORACLE repeatedly on the inputs of tests/test_history_gpu.py (fresh tables and a fresh context per
repetition, as the tests do) and reports repetitions whose spectra differ from the first one by more than the parity
tolerance -- summation order over threads moves the last bits only, anything larger is a defect of the oracle."""
import ctypes as C
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

import numpy as np  # noqa: E402
import orc  # noqa: E402
import xmimsim_b200 as x  # noqa: E402
from inputs import example, synthetic_layers  # noqa: E402


def run(inp, opt, n_threads, grid_n=64):
    sim = x.Simulation(inp, quality=0)
    ci = x.CInput(inp)
    od = orc.init_input(C.pointer(ci.input))
    n_total = orc.lib().orc_total_histories(C.cast(C.pointer(ci.input), C.c_void_p))
    r_full, t_full = sim.solid_angle_inputs()
    r = np.linspace(r_full[0], r_full[-1], grid_n)
    t = np.linspace(t_full[0], t_full[-1], grid_n)
    rr = np.random.default_rng(5)
    sa = sim.make_solid_angle(rr.uniform(1e-4, 2e-4, (grid_n, grid_n)), r.copy(), t.copy())
    ch, vr, cnt = orc.main_msim_range(C.pointer(ci.input), od, sim.L.xmb_get_tables(sim.hdf5F), opt, sa, 0x584D494D53494D, 0,
                                      n_total, inp.n_interactions_trajectory, inp.nchannels, n_threads)
    ch = ch.copy(); vr = vr.copy()
    sim.close()
    return ch, vr, [int(c) for c in cnt]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    n_threads = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    cases = {"synthetic10": lambda: synthetic_layers(n_photons=30000, n_int=8)}
    a = lambda: example("srm1412")
    def srm():
        i = a(); i.n_photons_line = 800
        return i
    cases["srm1412"] = srm
    for name, mk in cases.items():
        first = None
        bad = 0
        for rep in range(reps):
            ch, vr, cnt = run(mk(), x.main_options(), n_threads)
            if first is None:
                first = (ch, vr, cnt)
                continue
            e_ch = float(np.abs(ch - first[0]).max() / np.abs(first[0]).max())
            e_vr = float(np.abs(vr - first[1]).max() / np.abs(first[1]).max())
            if e_ch > 1e-9 or e_vr > 1e-9 or cnt != first[2]:
                bad += 1
                print(json.dumps({"case": name, "rep": rep, "err_ch": e_ch, "err_vr": e_vr, "cnt": cnt, "cnt0": first[2]}), flush=True)
        print(json.dumps({"case": name, "reps": reps, "bad": bad}), flush=True)


if __name__ == "__main__":
    main()
