"""One detector-response pass (escape peaks + pile-up + Poisson) on a BASELINE configs[4]-shaped spectrum, for ncu
(tools/profile_kernels.sh det).  Prints the device time of the response."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

import xmimsim_b200 as x  # noqa: E402
from xmimsim_b200 import workloads  # noqa: E402

inp = workloads.ebel_like(n_intervals=1000, n_photons_interval=2000, n_photons_line=2000)
sim = x.Simulation(inp, quality=0)
g, r, t = sim.solid_angle_calculation(hits_per_single=500, seed=1)
sa = sim.make_solid_angle(g.copy(), r.copy(), t.copy())
opt = x.main_options(use_sum_peaks=1, use_escape_peaks=1, use_poisson=1)
ch, br, vr = sim.main_msim(opt, sa)
er = sim.escape_ratios_calculation(options=opt)
conv = sim.detector_convolute_all(ch, br, vr, opt, er.contents)
print(json.dumps({"detector_ms": sim.L.xmb_detector_last_ms(), "launches": int(sim.L.xmb_detector_last_launches()), "sum": float(conv[-1].sum())}))
sim.escape_ratios_free(er)
sim.close()
