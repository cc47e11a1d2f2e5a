"""Runs the history engine repeatedly on small inputs (partial batches: the drain path of the scheduler) and reports
how many repetitions differ from the first one -- the integer sums must never differ."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402
import xmimsim_b200 as x  # noqa: E402
from inputs import example, synthetic_layers  # noqa: E402


def check(name, inp, reps, **optkw):
    sim = x.Simulation(inp, quality=0)
    r_full, t_full = sim.solid_angle_inputs()
    r, t = r_full[::8], t_full[::8]
    rr = np.random.default_rng(5)
    sa = sim.make_solid_angle(rr.uniform(1e-4, 2e-4, (t.size, r.size)), r.copy(), t.copy())
    opt = x.main_options(**optkw)
    shas = {}
    for _ in range(reps):
        limbs, ex = sim.main_msim_raw(opt, sa)
        h = hashlib.sha256(limbs.tobytes()).hexdigest()[:12]
        shas[h] = shas.get(h, 0) + 1
    sim.close()
    return {"input": name, "reps": reps, "digests": shas}


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    out = [check("synthetic10 30000 photons, 8 interactions", synthetic_layers(n_photons=30000, n_int=8), reps),
           check("synthetic10 300000 photons", synthetic_layers(n_photons=300000, n_int=8), reps)]
    a = example("srm1412"); a.n_photons_line = 3000
    out.append(check("srm1412 3000 photons/line", a, reps))
    print(json.dumps({"lib": os.environ.get("XMIMSIM_B200_LIB", "default"), "results": out}))


if __name__ == "__main__":
    main()
