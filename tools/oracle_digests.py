"""CPU-only: SHA-256 digests of the oracle's history output (one thread: fixed summation order) on ten inputs -- the
four shipped examples, the 10-layer synthetic sample and five option variants.  Argument: an alternative oracle library
(tools/sanitize_host.sh builds sanitizer / auto-var-init variants) or "default"."""
import ctypes as C
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

import numpy as np  # noqa: E402
import orc  # noqa: E402

if len(sys.argv) > 1 and sys.argv[1] != "default":
    orc.LIB_PATH = sys.argv[1]
import xmimsim_b200 as x  # noqa: E402
from inputs import example, synthetic_layers  # noqa: E402


def run(inp, opt, grid_n=64):
    sim = x.Simulation(inp, quality=0)
    ci = x.CInput(inp)
    od = orc.init_input(C.pointer(ci.input))
    n_total = orc.lib().orc_total_histories(C.cast(C.pointer(ci.input), C.c_void_p))
    r_full, t_full = sim.solid_angle_inputs()
    r = np.linspace(r_full[0], r_full[-1], grid_n)
    t = np.linspace(t_full[0], t_full[-1], grid_n)
    sa = sim.make_solid_angle(np.random.default_rng(5).uniform(1e-4, 2e-4, (grid_n, grid_n)), r.copy(), t.copy())
    if opt.use_advanced_compton:
        sim.L.xmb_tables_enable_advanced_compton(sim.hdf5F)
    ch, vr, cnt = orc.main_msim_range(C.pointer(ci.input), od, sim.L.xmb_get_tables(sim.hdf5F), opt, sa, 0x584D494D53494D, 0,
                                      n_total, inp.n_interactions_trajectory, inp.nchannels, 1)
    h = hashlib.sha256(ch.tobytes() + vr.tobytes()).hexdigest()[:16]
    sim.close()
    return h, [int(c) for c in cnt]


def main():
    out = {"synthetic10": run(synthetic_layers(n_photons=8000, n_int=8), x.main_options())}
    for nm in ("srm1155", "srm1412", "srm1132", "In"):
        a = example(nm); a.n_photons_line = 200
        out[nm] = run(a, x.main_options())
    for k, o in enumerate((dict(use_M_lines=0), dict(use_cascade_auger=0), dict(use_cascade_radiative=0),
                           dict(use_cascade_auger=0, use_cascade_radiative=0), dict(use_advanced_compton=1))):
        a = example("srm1412"); a.n_photons_line = 200
        out["srm1412 option variant %d" % k] = run(a, x.main_options(**o))
    print(json.dumps(out, sort_keys=True))


if __name__ == "__main__":
    main()
