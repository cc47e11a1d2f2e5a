import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests"); sys.path.insert(0, ROOT + "/oracle")
import numpy as np, xmimsim_b200 as x
from helpers import Pair
from inputs import example
def run(nlines, n, n_int, gridn):
    inp = example("srm1155"); inp.n_photons_line = n; inp.n_interactions_trajectory = n_int
    inp.discrete = inp.discrete[:nlines]
    P = Pair(inp); sa = P.grid(hits_per_single=400, n=gridn)
    o = x.main_options()
    ch, br, vr = P.sim.main_msim(o, sa)
    ch_o, vr_o, cnt = P.oracle(o, sa, 0, n_threads=16)
    wmax = max(d.horizontal_intensity + d.vertical_intensity for d in inp.discrete)
    unit = wmax / n * inp.live_time * 2.0 ** -56
    print("lines", nlines, "N/line", n, "n_int", n_int, "grid", gridn, "unit %.3e" % unit)
    for (z, l) in [(26, 3), (24, 3), (16, 5), (82, 126), (82, 119), (26, 384), (26, 385)]:
        g, r = vr[z - 1, l - 1, 0], vr_o[z - 1, l - 1, 0]
        print("   Z%d slot %d gpu %.10e orc %.10e diff/unit %.1f rel %.2e" % (z, l, g, r, (g - r) / unit, (g - r) / r))
    P.close()
run(1, 1500, 1, 128)
run(1, 20000, 1, 128)
run(26, 1500, 1, 128)
run(1, 1500, 4, 128)
run(1, 1500, 1, None)
