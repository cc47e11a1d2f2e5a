"""One small engine run of the 10-layer / 8-interaction sample (partial batches, per-layer queues, drain path) and of
srm1412 (energy-class batches), meant to run under compute-sanitizer (tools/sanitize_gpu.sh).  No torch import."""
import hashlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402
import xmimsim_b200 as x  # noqa: E402
from inputs import example, synthetic_layers  # noqa: E402


def one(name, inp):
    sim = x.Simulation(inp, quality=0)
    r_full, t_full = sim.solid_angle_inputs()
    r, t = r_full[::16], t_full[::16]
    sa = sim.make_solid_angle(np.random.default_rng(5).uniform(1e-4, 2e-4, (t.size, r.size)), r.copy(), t.copy())
    limbs, ex = sim.main_msim_raw(x.main_options(), sa)
    print(name, hashlib.sha256(limbs.tobytes()).hexdigest()[:12], flush=True)
    sim.close()


which = sys.argv[1] if len(sys.argv) > 1 else "syn"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
if which == "syn":
    one("synthetic10", synthetic_layers(n_photons=n, n_int=8))
else:
    a = example("srm1412"); a.n_photons_line = max(1, n // 25)
    one("srm1412", a)
